// bmu_tc_kernel.cuh -- the tensor-core BMU kernel template (see bmu_tc.cu for the overview).  Included
// by the bmu_tc_inst_*.cu translation units, each of which instantiates one family of variants.
#pragma once
#include <float.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pixie {

using namespace ptx;

// Image layout (bytes): block b (32 columns) at b * Ntot * 128, row r at r * 128, 16-byte chunk c at
// ((c ^ (r & 7)) << 4) -- exactly what TMA SWIZZLE_128B would have written, so one linear
// cp.async.bulk brings it into place.  Column j < C of row k holds -2 * W[k, j].  Behind the
// blocks sits the bias block (plan.off_bias): one 8-column K-step per row in the no-swizzle
// core-matrix layout (bias_offset), columns 0..2 = ||w_k||^2 split into three tf32-exact terms
// (multiplied by the all-ones A tile of the bias K-step); rows >= K hold zeros and a huge bias so
// they never win.
__device__ __forceinline__ uint32_t img_offset(int Ntot, int row, int col)
{
    const int b = col >> 5, cc = col & 31;
    return (uint32_t)b * (uint32_t)Ntot * 128u + (uint32_t)row * 128u +
           (uint32_t)((((cc >> 2) ^ (row & 7)) << 4) + ((cc & 3) << 2));
}

// K-major no-swizzle operand, 8 columns (one tf32 K-step) per row: 8-row x 16-byte core matrices,
// the two K-adjacent ones 128 bytes apart (LBO), 8-row groups 256 bytes apart (SBO).
__host__ __device__ __forceinline__ uint32_t bias_offset(int row, int col)
{
    return (uint32_t)(row >> 3) * 256u + (uint32_t)(col >> 2) * 128u + (uint32_t)(row & 7) * 16u +
           (uint32_t)(col & 3) * 4u;
}

// ------------------------------------------------------------------------------------------------
// device helpers shared by the epilogue stages
// ------------------------------------------------------------------------------------------------
// Both the X stages (written by TMA SWIZZLE_128B) and the codebook image keep logical 16-byte
// chunk c of row r of a 32-column block at physical chunk (c ^ (r & 7)).
__device__ __forceinline__ const float4 *x_chunk_ptr(const uint8_t *xs, int row, int blk, int chunk)
{
    return reinterpret_cast<const float4 *>(xs + (size_t)blk * 16384u + (size_t)row * 128u +
                                            (size_t)(((chunk ^ (row & 7)) & 7) << 4));
}
__device__ __forceinline__ const float4 *w_chunk_ptr(const uint8_t *ws, int Ntot, int node, int blk,
                                                     int chunk)
{
    return reinterpret_cast<const float4 *>(ws + (size_t)blk * (size_t)Ntot * 128u +
                                            (size_t)node * 128u +
                                            (size_t)(((chunk ^ (node & 7)) & 7) << 4));
}

// acc += (x + 0.5 w')^2 over one 16-byte chunk, packed fp32 (w' = -2 w, so x + 0.5 w' = x - w)
__device__ __forceinline__ void dist2_chunk(const float4 x, const float4 w, uint64_t half2,
                                            uint64_t &acc0, uint64_t &acc1)
{
    const uint64_t d0 = fma2(pack2(w.x, w.y), half2, pack2(x.x, x.y));
    const uint64_t d1 = fma2(pack2(w.z, w.w), half2, pack2(x.z, x.w));
    acc0 = fma2(d0, d0, acc0);
    acc1 = fma2(d1, d1, acc1);
}

// stage 2: fp32 squared distance between tile row `row` and codebook node `node`.
// Full 32-channel blocks are walked in PHYSICAL chunk order (the sum does not care about order):
// physical X chunk pc holds logical chunk pc ^ (row & 7), which sits in the codebook row at
// physical chunk pc ^ (row & 7) ^ (node & 7) -- one XOR per chunk, no per-operand swizzle math.
// Lane l starts at physical chunk `rot` = l & 7 (callers pass it): the lanes of a warp read
// DIFFERENT rows, and chunk pc of eight different rows is one 16-byte bank group -- an 8-way
// conflict -- while chunks pc ^ rot spread a quarter-warp over all eight groups.
// The last, partial block is walked logically so the bias columns of the image are never read.
__device__ __forceinline__ float pair_dist2_f32(const uint8_t *xs, const uint8_t *ws, int Ntot,
                                                int nchunks16, int row, int node, uint32_t rot)
{
    const uint64_t half2 = pack2(0.5f, 0.5f);
    uint64_t acc0 = 0ull, acc1 = 0ull;
    const int nfull = nchunks16 >> 3;
    const uint32_t t = (uint32_t)((row ^ node) & 7);
    // rows are 128-byte aligned: OR-ing the start chunk in and XOR-ing pc walks pc ^ rot
    uint32_t xrow = smem_u32(xs) + (uint32_t)row * 128u + (rot << 4);
    uint32_t wrow = smem_u32(ws) + (uint32_t)node * 128u + ((rot ^ t) << 4);
    for (int b = 0; b < nfull; ++b) {
#pragma unroll
        for (uint32_t pc = 0; pc < 8; ++pc) {
            const uint4 xu = lds128(xrow ^ (pc << 4));
            const uint4 wu = lds128(wrow ^ (pc << 4));
            const float4 x = make_float4(__uint_as_float(xu.x), __uint_as_float(xu.y),
                                         __uint_as_float(xu.z), __uint_as_float(xu.w));
            const float4 w = make_float4(__uint_as_float(wu.x), __uint_as_float(wu.y),
                                         __uint_as_float(wu.z), __uint_as_float(wu.w));
            dist2_chunk(x, w, half2, acc0, acc1);
        }
        xrow += 16384u;
        wrow += (uint32_t)Ntot * 128u;
    }
    const int rem = nchunks16 & 7;
    for (int lc = 0; lc < rem; ++lc) {
        const float4 x = *x_chunk_ptr(xs, row, nfull, lc);
        const float4 w = *w_chunk_ptr(ws, Ntot, node, nfull, lc);
        dist2_chunk(x, w, half2, acc0, acc1);
    }
    float a, b, c, d;
    unpack2(acc0, a, b);
    unpack2(acc1, c, d);
    return (a + b) + (c + d);
}

// stage 2, two-candidate form: fp32 squared distances of tile row `row` to nodes `na` and `nb` in
// one walk over the row (the common case: a flagged row has exactly two candidates).  The thread
// reads ITS OWN row, chunk c at physical chunk c ^ (row & 7), so a quarter-warp (8 consecutive
// rows) covers all eight bank groups; codebook chunks follow the same logical order.
__device__ __forceinline__ void duel_dist2_f32(const uint8_t *xs, const uint8_t *ws, int Ntot,
                                               int nchunks16, int row, int na, int nb, float &da,
                                               float &db)
{
    const uint64_t half2 = pack2(0.5f, 0.5f);
    uint64_t a0 = 0ull, a1 = 0ull, b0 = 0ull, b1 = 0ull;
    // rows are 128-byte aligned: OR the row's swizzle key in, XOR the logical chunk index
    uint32_t xrow = smem_u32(xs) + (uint32_t)row * 128u + ((uint32_t)(row & 7) << 4);
    uint32_t wa = smem_u32(ws) + (uint32_t)na * 128u + ((uint32_t)(na & 7) << 4);
    uint32_t wb = smem_u32(ws) + (uint32_t)nb * 128u + ((uint32_t)(nb & 7) << 4);
    for (int left = nchunks16; left > 0; left -= 8) {
        if (left >= 8) {
#pragma unroll
            for (uint32_t c = 0; c < 8; ++c) {
                const uint4 xu = lds128(xrow ^ (c << 4));
                const uint4 au = lds128(wa ^ (c << 4));
                const uint4 bu = lds128(wb ^ (c << 4));
                const float4 x = make_float4(__uint_as_float(xu.x), __uint_as_float(xu.y),
                                             __uint_as_float(xu.z), __uint_as_float(xu.w));
                dist2_chunk(x, make_float4(__uint_as_float(au.x), __uint_as_float(au.y),
                                           __uint_as_float(au.z), __uint_as_float(au.w)),
                            half2, a0, a1);
                dist2_chunk(x, make_float4(__uint_as_float(bu.x), __uint_as_float(bu.y),
                                           __uint_as_float(bu.z), __uint_as_float(bu.w)),
                            half2, b0, b1);
            }
        } else {
            for (uint32_t c = 0; c < (uint32_t)left; ++c) {
                const uint4 xu = lds128(xrow ^ (c << 4));
                const uint4 au = lds128(wa ^ (c << 4));
                const uint4 bu = lds128(wb ^ (c << 4));
                const float4 x = make_float4(__uint_as_float(xu.x), __uint_as_float(xu.y),
                                             __uint_as_float(xu.z), __uint_as_float(xu.w));
                dist2_chunk(x, make_float4(__uint_as_float(au.x), __uint_as_float(au.y),
                                           __uint_as_float(au.z), __uint_as_float(au.w)),
                            half2, a0, a1);
                dist2_chunk(x, make_float4(__uint_as_float(bu.x), __uint_as_float(bu.y),
                                           __uint_as_float(bu.z), __uint_as_float(bu.w)),
                            half2, b0, b1);
            }
        }
        xrow += 16384u;
        wa += (uint32_t)Ntot * 128u;
        wb += (uint32_t)Ntot * 128u;
    }
    float p, q, r, t;
    unpack2(a0, p, q);
    unpack2(a1, r, t);
    da = (p + q) + (r + t);
    unpack2(b0, p, q);
    unpack2(b1, r, t);
    db = (p + q) + (r + t);
}

// stage 3: the reference's fp64 operation sequence for one (row, node) pair
// (oracle/pixie_oracle.c nearest_node): tmp = x - w; acc = acc + tmp * tmp (separately rounded),
// in channel order; d = sqrt(acc).
static __device__ __noinline__ double pair_dist_f64(const uint8_t *xs, const uint8_t *ws, int Ntot, int C,
                                             int row, int node)
{
    double acc = 0.0;
    for (int j = 0; j < C; ++j) {
        const int blk = j >> 5, cc = j & 31;
        const float xf = reinterpret_cast<const float *>(x_chunk_ptr(xs, row, blk, cc >> 2))[cc & 3];
        const float wf = reinterpret_cast<const float *>(w_chunk_ptr(ws, Ntot, node, blk, cc >> 2))[cc & 3];
        const double tmp = __dsub_rn((double)xf, (double)(-0.5f * wf));
        acc = __dadd_rn(acc, __dmul_rn(tmp, tmp));
    }
    return __dsqrt_rn(acc);
}

// Grid-wide barrier for the persistent kernel (every CTA is resident: grid <= SM count and one CTA
// per SM fits).  `counter` only ever grows during a launch; `target` is the value it reaches when
// all CTAs have arrived at this barrier instance.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1u);
        unsigned spins = 0;
        while (*reinterpret_cast<volatile unsigned int *>(counter) < target) {
            __nanosleep(32);
            if (++spins > (1u << 25)) __trap();  // seconds: never on a healthy launch
        }
        __threadfence();
    }
    __syncthreads();
}

// tiles of one mini-batch step that this shard holds: local tiles first, first + B, ... whose
// GLOBAL index (local + tile_offset) is congruent to m mod B
struct StepTiles {
    int64_t first, stride, count;
};
__device__ __forceinline__ StepTiles step_tiles(const TcParams &p, int st)
{
    StepTiles r;
    if (p.nsteps <= 1 && !p.apply) {
        r.first = p.tile_first;
        r.stride = p.tile_stride;
        r.count = p.ntiles;
        return r;
    }
    const int64_t B = p.B;
    const int64_t m = (int64_t)(p.t0 + st) % B;
    r.first = ((m - p.tile_offset) % B + B) % B;
    r.stride = B;
    r.count = r.first < p.tiles_total ? (p.tiles_total - r.first + B - 1) / B : 0;
    return r;
}

// ------------------------------------------------------------------------------------------------
// whole-pass mode, end of a step: fold the CTAs' sums, apply the batch update (DESIGN.md section 4,
// same arithmetic as som_apply_kernel) and rewrite the codebook image for the next step.
// Called by every thread of every CTA; three grid barriers.
// ------------------------------------------------------------------------------------------------
template <int NG>
__device__ __noinline__ void step_update(const TcParams &p, int st, uint8_t *smem,
                                         unsigned int &gb_target)
{
    const TcPlan &pl = p.plan;
    const int K = pl.K, C = pl.C, len = K * (C + 1);
    const int tid = threadIdx.x, nthr = blockDim.x;
    float *acc = reinterpret_cast<float *>(smem + pl.off_acc);
    unsigned int *gsync = &p.ctl->grid_sync;
    const bool timing = blockIdx.x == 0 && tid == 0;
    uint64_t tm = timing ? global_timer_ns() : 0;
    auto lap = [&](int slot) {
        if (timing) {
            const uint64_t now = global_timer_ns();
            p.ctl->phase_ns[slot] += now - tm;
            tm = now;
        }
    };

    // 1. groups -> this CTA's partial (group order), accumulators cleared for the next step
    float *mine = p.partials + (size_t)blockIdx.x * len;
    for (int i = tid; i < len; i += nthr) {
        float v = acc[i];
        acc[i] = 0.f;
#pragma unroll
        for (int gg = 1; gg < NG; ++gg) {
            v += acc[gg * len + i];
            acc[gg * len + i] = 0.f;
        }
        mine[i] = v;
    }
    gb_target += gridDim.x;
    grid_barrier(gsync, gb_target);
    lap(1);

    // 2. every CTA folds its slice of the table over all partials, CTA order, fp64
    double *fold_dst = p.world > 1 ? p.peer_buf[p.rank] + (size_t)(st & 1) * len : p.SN;
    {
        const int nparts = gridDim.x;
        const int per = (len + nparts - 1) / nparts;
        const int e0 = blockIdx.x * per;
        const int e1 = min(len, e0 + per);
        const int oct = tid >> 3, q = tid & 7;
        const int rounds = (per + nthr / 8 - 1) / (nthr / 8);  // uniform trip count
        for (int it = 0; it < rounds; ++it) {
            const int e = e0 + it * (nthr / 8) + oct;
            float v[kFoldMax];
#pragma unroll
            for (int u = 0; u < kFoldMax; ++u) {
                const int pp = q + 8 * u;
                v[u] = (e < e1 && pp < nparts) ? __ldcg(p.partials + (size_t)pp * len + e) : 0.f;
            }
            double a = 0.0;
#pragma unroll
            for (int u = 0; u < kFoldMax; ++u) a += (double)v[u];
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            if (e < e1 && q == 0) fold_dst[e] = a;
        }
    }
    if (p.world > 1) {
        // ---- cross-GPU sum over NVLink peer memory, fused into the step: every rank has folded
        // its shard's table into its exchange buffer; after a flag handshake each CTA sums ITS
        // slice over the ranks in rank order (so every rank computes bit-identical totals).
        gb_target += gridDim.x;
        grid_barrier(gsync, gb_target);  // this rank's buffer is complete
        if (blockIdx.x == 0 && tid == 0) {
            __threadfence_system();
            const uint32_t val = p.flag_base + (uint32_t)st + 1u;
            const size_t flag_off = (size_t)2 * len * sizeof(double);
            for (int r = 0; r < p.world; ++r) {
                if (r == p.rank) continue;
                uint32_t *dst = reinterpret_cast<uint32_t *>(
                                    reinterpret_cast<char *>(p.peer_buf[r]) + flag_off) + p.rank;
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(val) : "memory");
            }
            const uint32_t *mine_flags = reinterpret_cast<const uint32_t *>(
                reinterpret_cast<const char *>(p.peer_buf[p.rank]) + flag_off);
            for (int r = 0; r < p.world; ++r) {
                if (r == p.rank) continue;
                uint32_t seen, spins = 0;
                do {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine_flags + r) : "memory");
                    if (seen - val < 0x80000000u) break;  // seen >= val (wrap-safe)
                    __nanosleep(64);
                    if (++spins > (1u << 26)) __trap();  // a peer died: fail instead of hanging
                } while (true);
            }
        }
        gb_target += gridDim.x;
        grid_barrier(gsync, gb_target);  // all ranks' buffers are complete and visible
        const int nparts = gridDim.x;
        const int per = (len + nparts - 1) / nparts;
        const int e0 = blockIdx.x * per;
        const int e1 = min(len, e0 + per);
        for (int e = e0 + tid; e < e1; e += nthr) {
            // all ranks' values first (up to 8 NVLink round trips in flight together instead of one
            // after the other), then the sum in rank order
            double v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                v[r] = 0.0;
                if (r < p.world) {
                    const double *src = p.peer_buf[r] + (size_t)(st & 1) * len + e;
                    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v[r]) : "l"(src) : "memory");
                }
            }
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < p.world) a += v[r];
            p.SN[e] = a;
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        // slot the update below publishes the next step's norms into
        p.ctl->pp_wmax_bits[(st + 1) & 1] = 0;
        p.ctl->pp_w_has_negative[(st + 1) & 1] = 0;
    }
    lap(2);
    gb_target += gridDim.x;
    grid_barrier(gsync, gb_target);
    lap(3);

    // 3. batch update of node k by CTA k (k < K); schedule of step t = t0 + st of T
    {
        const double frac = (double)(p.t0 + st) / (double)p.T;
        const double r = p.r0 - (p.r0 - p.r1) * frac;
        const double r_eff = r < 1.0 ? 0.5 : r;
        const double sigma = 0.5 * r_eff;
        const double inv2s2 = 1.0 / (2.0 * sigma * sigma);
        const double alpha = p.a0 - (p.a0 - p.a1) * frac;
        // the folded table, staged once per CTA in the (idle, already cleared) accumulator area
        double *s_sn = reinterpret_cast<double *>(acc);
        if ((int)blockIdx.x < K)
            for (int i = tid; i < len; i += nthr) s_sn[i] = __ldcg(p.SN + i);
        __syncthreads();
        double *s_h = reinterpret_cast<double *>(smem + pl.off_pairs);  // [K], pair lists are idle
        double *s_cnt = s_h + K;                                        // [K]
        double *s_red = s_cnt + K;                                      // [32] block reduction
        int *s_flag = reinterpret_cast<int *>(s_red + 32);
        const int ydim = p.ydim;
        for (int k = blockIdx.x; k < K; k += gridDim.x) {
            const int kx = k / ydim, ky = k % ydim;
            for (int b = tid; b < K; b += nthr) {
                const int dx = abs(kx - b / ydim), dy = abs(ky - b % ydim);
                const double d = (double)(dx > dy ? dx : dy);
                const double cnt = s_sn[(size_t)b * (C + 1) + C];
                s_cnt[b] = cnt;
                s_h[b] = cnt == 0.0 ? 0.0 : exp(-d * d * inv2s2);
            }
            if (tid == 0) *s_flag = 0;
            __syncthreads();
            double den = 0.0;
            for (int b = 0; b < K; ++b) den += s_h[b] * s_cnt[b];
            const double beta = den > 0.0 ? 1.0 - pow(1.0 - alpha, den) : 0.0;
            double nrm2 = 0.0;
            bool neg = false;
            char *img = reinterpret_cast<char *>(p.wimg_rw);
            for (int c = tid; c < C; c += nthr) {
                double w = p.W64[(size_t)k * C + c];
                if (den > 0.0) {
                    double num = 0.0;
                    for (int b = 0; b < K; ++b) num += s_h[b] * s_sn[(size_t)b * (C + 1) + c];
                    w += beta * (num / den - w);
                    p.W64[(size_t)k * C + c] = w;
                }
                const float wf = (float)w;
                p.W32[(size_t)k * C + c] = wf;
                *reinterpret_cast<float *>(img + img_offset(pl.Ntot, k, c)) = -2.0f * wf;
                nrm2 += (double)wf * (double)wf;
                if (__float_as_int(wf) < 0) neg = true;
            }
            // block reduction of ||w_k||^2 (only the first C threads contribute)
            for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
            if (neg) atomicOr(s_flag, 1);
            if ((tid & 31) == 0 && (tid >> 5) < 32) s_red[tid >> 5] = nrm2;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < (nthr + 31) / 32 && w < 32; ++w) tot += s_red[w];
                float bias = (float)tot;
                if (!(bias <= FLT_MAX)) bias = FLT_MAX;
                const float h = __uint_as_float(__float_as_uint(bias) & 0xFFFFE000u);
                const float r1 = bias - h;
                const float m = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
                const float l = r1 - m;
                float *bb = reinterpret_cast<float *>(img + pl.off_bias + bias_offset(k, 0));
                bb[0] = h;
                bb[1] = m;
                bb[2] = l;
                float nr = (float)sqrt(tot) * 1.0000005f;
                if (!(nr <= FLT_MAX)) nr = FLT_MAX;
                atomicMax(&p.ctl->pp_wmax_bits[(st + 1) & 1], __float_as_int(nr));
                if (*s_flag) atomicOr(&p.ctl->pp_w_has_negative[(st + 1) & 1], 1);
            }
            __syncthreads();
        }
    }
    // the staged table sat in the accumulator area: clear it again for the next step
    if ((int)blockIdx.x < K)
        for (int i = tid; i < (len * 2 + 1); i += nthr)
            if (i < NG * len) acc[i] = 0.f;
    lap(4);
    gb_target += gridDim.x;
    grid_barrier(gsync, gb_target);
    lap(5);
}

// ------------------------------------------------------------------------------------------------
// train mode: add one tile's rows to the epilogue group's private K x (C+1) table (sums + counts).
// Called by the 128 threads of group g once the tile's labels are known (L = label bin of this
// thread's row, K = "skip": padding and unassignable rows).  Deterministic, no atomics:
//   1. stable counting sort of the 128 rows by label: rank inside the warp from match.any, one
//      histogram byte per (warp, label), every warp scans the bins itself (32 lanes x nbl bins);
//   2. warp w walks sorted positions [32w, 32w+32) with lanes = channels: the rows of a node are
//      contiguous, so a node's sum is a register accumulation over independent shared-memory loads
//      (the walk that used to be a 128-long chain of dependent read-modify-writes on the table);
//      each finished segment is one read-modify-write on cells no other warp touches -- except the
//      warp's FIRST segment, whose node may continue from the previous warp's range: that one goes
//      to a side buffer;
//   3. after a group barrier the side buffers are added in warp order.
// Sum order inside a tile: ascending row inside a warp range, warp ranges in order.
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void tile_accumulate_sorted(uint8_t *smem, const uint8_t *xs,
                                                           uint32_t off_acc, uint32_t off_sort, int K,
                                                           int C, int nblkX, int g, int quad,
                                                           int lane, int L)
{
    const SortLayout sl = sort_layout(C, K);
    uint8_t *base = smem + off_sort;
    uint32_t *hist32 = reinterpret_cast<uint32_t *>(base);
    uint8_t *wbase = base + sl.off_wbase + (uint32_t)quad * sl.bins;
    uint8_t *order = base + sl.off_order;
    uint16_t *slab = reinterpret_cast<uint16_t *>(base + sl.off_slab);
    float *side = reinterpret_cast<float *>(base + sl.off_side);
    int *side_lab = reinterpret_cast<int *>(base + sl.off_sidelab);
    const int acc_ld = C + 1;
    float *tab = reinterpret_cast<float *>(smem + off_acc) + (size_t)g * K * acc_ld;
    const uint32_t bar = 1u + (uint32_t)g;

    // 1a. rows of this warp with my label; the lowest such lane publishes their count
    const unsigned peers = __match_any_sync(0xffffffffu, L);
    const int rk = __popc(peers & ((1u << lane) - 1u));
    if (rk == 0) base[4 * L + quad] = (uint8_t)__popc(peers);
    bar_sync(bar, 128);
    // 1b. every warp scans all bins: lane l owns bins [l * nbl, (l + 1) * nbl)
    {
        const uint32_t *hp = hist32 + (uint32_t)lane * sl.nbl;
        uint32_t run = 0;
        for (uint32_t i = 0; i < sl.nbl; ++i) run += __dp4a(hp[i], 0x01010101u, 0u);
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        uint32_t pos0 = inc - run;
        const uint32_t below = 0x01010101u & ((1u << (8 * quad)) - 1u);  // warps before this one
        uint8_t *wp = wbase + (uint32_t)lane * sl.nbl;
        for (uint32_t i = 0; i < sl.nbl; ++i) {
            const uint32_t w = hp[i];
            wp[i] = (uint8_t)(pos0 + __dp4a(w, below, 0u));
            pos0 += __dp4a(w, 0x01010101u, 0u);
        }
    }
    __syncwarp();
    {
        const int pos = (int)wbase[L] + rk;
        order[pos] = (uint8_t)(quad * 32 + lane);
        slab[pos] = (uint16_t)L;
    }
    bar_sync(bar, 128);
    if (rk == 0) base[4 * L + quad] = 0;  // histograms are zero again for the next tile

    // 2. segment walk
    const int my_row = order[quad * 32 + lane];
    const int my_L = slab[quad * 32 + lane];
    if (lane == 0) side_lab[quad] = -1;
    for (int cb = 0; cb < nblkX; ++cb) {
        const int ch = cb * 32 + lane;
        const uint8_t *xb = xs + (uint32_t)cb * 16384u + (uint32_t)(lane & 3) * 4u;
        const uint32_t chunk = (uint32_t)lane >> 2;
        float a = 0.f;
        int curL = __shfl_sync(0xffffffffu, my_L, 0), seg_len = 0;
        bool first = true;
        auto flush = [&]() {
            if (curL < K) {
                if (first) {
                    if (ch < C) side[quad * acc_ld + ch] = a;
                    if (cb == 0 && lane == 0) {
                        side[quad * acc_ld + C] = (float)seg_len;
                        side_lab[quad] = curL;
                    }
                } else {
                    if (ch < C) tab[curL * acc_ld + ch] += a;
                    if (cb == 0 && lane == 0) tab[curL * acc_ld + C] += (float)seg_len;
                }
            }
        };
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const uint32_t r = (uint32_t)__shfl_sync(0xffffffffu, my_row, i0 + u);
                v[u] = *reinterpret_cast<const float *>(xb + r * 128u + ((chunk ^ (r & 7u)) << 4));
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int Lu = __shfl_sync(0xffffffffu, my_L, i0 + u);
                if (Lu != curL) {  // warp-uniform
                    flush();
                    first = false;
                    curL = Lu;
                    a = 0.f;
                    seg_len = 0;
                }
                if (Lu < K) {
                    a += v[u];
                    ++seg_len;
                }
            }
        }
        if (seg_len > 0) flush();
    }
    bar_sync(bar, 128);
    // 3. side buffers, warp order; thread t of the group owns column t of the table
    const int col = quad * 32 + lane;
    if (col <= C) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int sL = side_lab[w];
            if (sL >= 0) tab[sL * acc_ld + col] += side[w * acc_ld + col];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int SL, int SPC, int NCH, int NG, bool ACC>
__global__ void __launch_bounds__(NG * 128 + 64, 1)
bmu_tc_kernel(const __grid_constant__ CUtensorMap tmX, const TcParams p)
{
    constexpr int NEPI = NG * 4;            // epilogue warps
    constexpr int NCHUNK = SL * SPC;        // codebook rows per accumulator chunk
    constexpr int NMMA = (NCHUNK + 15) / 16 * 16;  // UMMA N (columns past NCHUNK are never read)
    static_assert(NCH == 1 || NCHUNK % 8 == 0, "chunk base must stay on a swizzle-atom row");
    constexpr int NS = SPC * NCH;           // slices per tile
    constexpr int NW = (SL + 31) / 32;      // mask words per slice
    constexpr int NBUF = NCH == 1 ? NG : 2; // TMEM accumulator buffers
    static_assert(NCH == 1 || NG == 2, "two chunks per tile only with two epilogue groups");

    extern __shared__ uint8_t smem_raw[];
    const TcPlan &pl = p.plan;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // 1 KiB-aligned carve-up (SWIZZLE_128B atoms are 1024 bytes)
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;
    const uint32_t sbase = raw_u32 + pad;
    uint8_t *ws = smem;                  // codebook image
    uint8_t *ones = smem + pl.off_ones;  // 4 KiB of 1.0f: the A operand of the bias K-step
    uint8_t *xs0 = smem + pl.off_x;      // X stages
    const uint32_t bar0 = sbase + pl.off_bar;
    const uint32_t bar_full = bar0;                      // [kMaxStages]
    const uint32_t bar_empty = bar0 + 8u * kMaxStages;   // [kMaxStages]
    const uint32_t bar_tfull = bar0 + 16u * kMaxStages;  // [4]
    const uint32_t bar_tempty = bar_tfull + 32u;         // [4]
    const uint32_t bar_w = bar_tempty + 32u;             // codebook image landed
    volatile uint32_t *tmem_slot =
        reinterpret_cast<volatile uint32_t *>(smem + pl.off_bar + 16u * kMaxStages + 72u);

    const int nstage = pl.nstage;
    // plain assignment (ACC == false) is always a single step: let the compiler drop the loop
    const int nsteps = ACC ? (p.nsteps > 1 ? p.nsteps : 1) : 1;

    // ---------------------------------------------------------------- one-time setup
    if (warp == NEPI && lane == 0) {
        prefetch_tensormap(&tmX);
        for (int s = 0; s < nstage; ++s) {
            mbar_init(bar_full + 8u * s, 1);   // producer's arrive.expect_tx
            mbar_init(bar_empty + 8u * s, 4);  // one arrive per warp of the owning epilogue group
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(bar_tfull + 8u * b, 1);   // tcgen05.commit
            mbar_init(bar_tempty + 8u * b, 4);  // one arrive per epilogue warp of the consumer
        }
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    if (warp == NEPI + 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), (uint32_t)pl.tmem_cols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<float *>(ones)[i] = 1.0f;
    if constexpr (ACC) {
        float *a = reinterpret_cast<float *>(smem + pl.off_acc);
        for (int i = threadIdx.x; i < NG * pl.K * (pl.C + 1); i += blockDim.x) a[i] = 0.f;
        // the sort scratch keeps its histograms zero between tiles
        uint32_t *sc = reinterpret_cast<uint32_t *>(smem + pl.off_lab);
        for (uint32_t i = threadIdx.x; i < (uint32_t)NG * pl.sort_stride / 4u; i += blockDim.x) sc[i] = 0u;
    }
    fence_proxy_async();  // the ones tile is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Every role walks the same sequence: step st = 0..nsteps-1, within a step this CTA's tiles
    // it = 0..cnt-1 (tile j = blockIdx.x + it * gridDim.x of the step).  `seq` numbers the CTA's tiles
    // across ALL steps; it fixes the pipeline slot of a tile: X stage seq % nstage, epilogue group
    // seq % NG, accumulator use seq / NG -- so barrier phases simply keep running across steps.
    uint32_t base_seq = 0;
    unsigned int gb_target = 0;
    uint32_t st_flag = 0, st_pairs = 0, st_fp64 = 0, st_fix = 0;  // per-thread statistics
    uint64_t step_t0 = (ACC && blockIdx.x == 0 && threadIdx.x == 0) ? global_timer_ns() : 0;  // grid-barrier instances passed so far x gridDim.x
    for (int st = 0; st < nsteps; ++st) {
    const StepTiles stp = ACC ? step_tiles(p, st) : StepTiles{p.tile_first, p.tile_stride, p.ntiles};
    const int64_t ntiles = stp.count;
    const uint32_t cnt =
        (int64_t)blockIdx.x < ntiles ? (uint32_t)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;

    if (warp == NEPI) {
        // ============================================================ TMA producer
        // The whole warp walks the loop (warp-uniform control flow); one elected lane issues.
        {
            if (elect_one()) {
                // codebook image: linear bulk copies (image is pre-swizzled in global memory); in
                // whole-pass mode it was rewritten by other CTAs through the generic proxy
                asm volatile("fence.proxy.async;" ::: "memory");
                mbar_arrive_expect_tx(bar_w, pl.wimg_bytes);
                for (uint32_t off = 0; off < pl.wimg_bytes; off += 16384u) {
                    const uint32_t sz = min(16384u, pl.wimg_bytes - off);
                    bulk_load(sbase + off, reinterpret_cast<const uint8_t *>(p.wimg) + off, sz, bar_w);
                }
            }
            __syncwarp();
            uint32_t s = base_seq % (uint32_t)nstage;          // one division per step, then
            uint32_t ph = (base_seq / (uint32_t)nstage) & 1u;  // incremental
            for (uint32_t it = 0; it < cnt; ++it, ph ^= (++s == (uint32_t)nstage), s = s == (uint32_t)nstage ? 0u : s) {
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                const int64_t j = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                const int64_t tile = stp.first + j * stp.stride;
                const int32_t row0 = (int32_t)(tile * kTile);
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_full + 8u * s, pl.stage_bytes);
                    for (int b = 0; b < pl.nblkX; ++b)
                        tma_load_2d(sbase + pl.off_x + s * pl.stage_bytes + (uint32_t)b * 16384u,
                                    &tmX, bar_full + 8u * s, b * 32, row0, kEvictFirst);
                }
                __syncwarp();
            }
        }
    } else if (warp == NEPI + 1) {
        // ============================================================ MMA issuer
        // The whole warp walks the loop and waits on the barriers; one elected lane (always the
        // same one, as tcgen05.commit requires) issues the MMAs of a tile and their commit.
        {
            constexpr uint32_t idesc = umma_idesc_tf32(128, (uint32_t)NMMA);
            const uint64_t desc_ones = umma_desc_nosw(sbase + pl.off_ones, 128u, 256u);
            // bias K-step: the no-swizzle block behind the image (8-row groups 256 bytes apart)
            const uint32_t wblk_bytes = (uint32_t)((NCH - 1) * NCHUNK + NMMA) * 128u;
            mbar_wait(bar_w, (uint32_t)st & 1u);
            uint32_t s = base_seq % (uint32_t)nstage;
            uint32_t ph = (base_seq / (uint32_t)nstage) & 1u;
            for (uint32_t it = 0; it < cnt; ++it, ph ^= (++s == (uint32_t)nstage), s = s == (uint32_t)nstage ? 0u : s) {
                const uint32_t seq = base_seq + it;
                mbar_wait(bar_full + 8u * s, ph);
                const uint32_t xs_addr = sbase + pl.off_x + s * pl.stage_bytes;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t q = seq * (uint32_t)NCH + (uint32_t)c;  // accumulator-chunk counter
                    const uint32_t buf = q % NBUF;
                    const uint32_t bph = (q / NBUF) & 1u;
                    mbar_wait(bar_tempty + 8u * buf, bph ^ 1u);
                    tc_fence_after();
                    if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + buf * (uint32_t)NMMA;
                    const uint32_t wrow = (uint32_t)(c * NCHUNK) * 128u;
                    for (int ks = 0; ks < pl.ksteps; ++ks) {
                        const uint32_t blk = (uint32_t)(ks >> 2), ko = (uint32_t)(ks & 3) * 32u;
                        const uint64_t da = umma_desc_sw128(xs_addr + blk * 16384u + ko);
                        const uint64_t db = umma_desc_sw128(sbase + blk * wblk_bytes + wrow + ko);
                        mma_tf32(d_tmem, da, db, idesc, ks > 0 ? 1u : 0u);
                    }
                    const uint64_t dbias = umma_desc_nosw(
                        sbase + pl.off_bias + (uint32_t)(c * NCHUNK / 8) * 256u, 128u, 256u);
                    mma_tf32(d_tmem, desc_ones, dbias, idesc, 1u);
                    mma_commit(bar_tfull + 8u * buf);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ============================================================ epilogue groups
        const int g = warp >> 2;             // group
        const int quad = warp & 3;           // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;    // tile row == TMEM lane
        const uint32_t r7 = (uint32_t)(row & 7);
        const int pair_cap = pl.pair_cap;
        uint32_t *pairs = reinterpret_cast<uint32_t *>(smem + pl.off_pairs) + warp * (2 * pair_cap);
        float *d2buf = reinterpret_cast<float *>(pairs + pair_cap);
        constexpr int Ntot = (NCH - 1) * NCHUNK + NMMA;
        const int nchunks16 = pl.C8 >> 2;    // 16-byte chunks holding real channels
        const float eps32 = (float)(pl.C + 8) * 2.4e-7f;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        mbar_wait(bar_w, (uint32_t)st & 1u);  // codebook image visible to this thread
        // norms of the codebook this step runs against: written by the prep kernel for the first
        // step, by the previous step's in-kernel update (ping-pong slot) afterwards
        const float wmax = __int_as_float((!ACC || st == 0) ? p.ctl->wmax_bits : p.ctl->pp_wmax_bits[st & 1]);
        const float wmax2 = wmax * wmax;
        const bool w_nonneg =
            ((!ACC || st == 0) ? p.ctl->w_has_negative : p.ctl->pp_w_has_negative[st & 1]) == 0;
        // fused accumulation (train mode)
        constexpr bool do_acc = ACC;

        // this group's tiles of the step: local indices it with (base_seq + it) % NG == g
        const uint32_t it0 = ((uint32_t)g + (uint32_t)NG - base_seq % (uint32_t)NG) % (uint32_t)NG;
        uint32_t s = (base_seq + it0) % (uint32_t)nstage;          // one division per step, then
        uint32_t ph = ((base_seq + it0) / (uint32_t)nstage) & 1u;  // incremental (+NG per tile)
        // row / label cursors advance by a constant per tile of this group (no 64-bit multiplies
        // in the tile loop)
        const int64_t j0 = (int64_t)blockIdx.x + (int64_t)it0 * gridDim.x;
        const int64_t jstep = (int64_t)NG * gridDim.x;
        int64_t grow = (stp.first + j0 * stp.stride) * kTile + row;  // global row
        const int64_t grow_step = jstep * stp.stride * kTile;
        int32_t *lab_ptr =
            p.labels ? p.labels + (p.compact_labels ? j0 * kTile + row : grow) : nullptr;
        const int64_t lab_step = p.labels ? (p.compact_labels ? jstep * kTile : grow_step) : 0;
        for (uint32_t it = it0; it < cnt;
             it += (uint32_t)NG, grow += grow_step, lab_ptr += lab_step) {
            const uint32_t seq = base_seq + it;
            const uint32_t use = seq / (uint32_t)NG;  // NG is a power of two: a shift
            const uint8_t *xs = xs0 + (uint32_t)s * pl.stage_bytes;

            mbar_wait(bar_full + 8u * s, ph);  // X tile landed

            // ---- per-row error bound of the tf32 scores.  ||x||^2 and the
            // sign test read whole 128-byte rows in physical order: channels past C are zero-filled
            // by TMA, and neither a sum of squares nor an OR of sign bits cares about chunk order.
            uint64_t xa = 0ull, xb = 0ull;
            uint32_t sgn = 0u;
            {
                // Lane i starts at physical chunk (i & 7) and walks chunks (i & 7) ^ pc, so the 8
                // lanes of a quarter-warp always hit 8 different 16-byte bank groups (reading
                // chunk pc from every row would be an 8-way bank conflict).
                uint32_t xaddr = sbase + pl.off_x + (uint32_t)s * pl.stage_bytes +
                                 (uint32_t)row * 128u + (r7 << 4);
                for (int b = 0; b < pl.nblkX; ++b, xaddr += 16384u) {
#pragma unroll
                    for (uint32_t pc = 0; pc < 8; ++pc) {
                        const uint4 x = lds128(xaddr ^ (pc << 4));
                        const uint64_t lo = pack2u(x.x, x.y), hi = pack2u(x.z, x.w);
                        xa = fma2(lo, lo, xa);
                        xb = fma2(hi, hi, xb);
                        sgn |= (x.x | x.y) | (x.z | x.w);
                    }
                }
            }
            float xn2;
            {
                float a, b, c, d;
                unpack2(xa, a, b);
                unpack2(xb, c, d);
                xn2 = (a + b) + (c + d);
            }
            // |score error| <= E = 2^-8 (1 + 1/16) ||x|| wmax + 2^-18 wmax^2.  In general the error
            // is two-sided and a node can only be the true minimum if its score is within 2E of
            // the smallest one.  When x and the codebook are both non-negative, operand truncation
            // can only RAISE a score (by at most E), so E suffices.
            // The 1e-30 floors keep the bound meaningful when squares or products underflow (the
            // error model above is relative): such rows simply collect candidates and end in fp64.
            const float E = 0.00415039f * sqrtf(xn2 + 1.0e-30f) * 1.000001f * wmax +
                            3.8147e-6f * wmax2 + 1.0e-30f;
            const float delta =
                ((w_nonneg && (int32_t)sgn >= 0) ? 1.03125f * E : 2.0f * E) * p.delta_scale;

            float m_run = __int_as_float(0x7f800000);
            uint32_t mw[NS][NW];
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int w = 0; w < NW; ++w) mw[a][w] = 0u;

#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint32_t buf, bph;
                if constexpr (NCH == 1) {
                    buf = (uint32_t)g;
                    bph = use & 1u;
                } else {
                    // chunk counter q = seq * 2 + c; buffers alternate
                    const uint32_t q = seq * 2u + (uint32_t)c;
                    buf = q & 1u;
                    bph = (q >> 1) & 1u;
                    // Both groups alternate on the same buffer, so this group may get here a whole
                    // phase early, where a parity wait would alias and fall through.  Waiting first
                    // for the previous use's release (made by the OTHER group after it saw the
                    // previous commit) pins the barrier to the right phase.
                    if (q >= 2) mbar_wait(bar_tempty + 8u * buf, bph ^ 1u);
                }
                mbar_wait(bar_tfull + 8u * buf, bph);
                tc_fence_after();
#pragma unroll
                for (int sidx = 0; sidx < SPC; ++sidx) {
                    const int sl = c * SPC + sidx;
                    uint32_t vr[SL];
                    tmem_ld_cols<SL>(tmem_lane + buf * (uint32_t)NMMA + (uint32_t)(sidx * SL), vr);
                    tc_wait_ld();
                    if (sidx == SPC - 1) {
                        // accumulator buffer fully read: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8u * buf);
                    }
                    // pass 1: slice minimum (four independent chains)
                    float a0 = __uint_as_float(vr[0]), a1 = __uint_as_float(vr[1]);
                    float a2 = __uint_as_float(vr[2]), a3 = __uint_as_float(vr[3]);
#pragma unroll
                    for (int i = 4; i + 7 < SL; i += 8) {
                        a0 = fminf(fminf(a0, __uint_as_float(vr[i])), __uint_as_float(vr[i + 4]));
                        a1 = fminf(fminf(a1, __uint_as_float(vr[i + 1])), __uint_as_float(vr[i + 5]));
                        a2 = fminf(fminf(a2, __uint_as_float(vr[i + 2])), __uint_as_float(vr[i + 6]));
                        a3 = fminf(fminf(a3, __uint_as_float(vr[i + 3])), __uint_as_float(vr[i + 7]));
                    }
#pragma unroll
                    for (int i = 4 + ((SL - 4) / 8) * 8; i < SL; ++i)
                        a0 = fminf(a0, __uint_as_float(vr[i]));
                    const float ms = fminf(fminf(a0, a1), fminf(a2, a3));
                    const float m_new = fminf(m_run, ms);
                    if (sl > 0) {
                        // earlier candidates are out of range once the minimum drops by > delta
                        const bool drop = m_new + delta < m_run;
#pragma unroll
                        for (int a = 0; a < NS; ++a)
                            if (a < sl)
#pragma unroll
                                for (int w = 0; w < NW; ++w) mw[a][w] = drop ? 0u : mw[a][w];
                    }
                    m_run = m_new;
                    const float thr = m_run + delta;
                    if (sl == 0 || __any_sync(0xffffffffu, ms < thr)) {
                        // pass 2: sign bit of (v - thr) funnel-shifted into a bit mask; value i of
                        // word w ends at bit (cnt_w - 1 - (i - 32 w))
                        const uint64_t thr2 = pack2(thr, thr);
#pragma unroll
                        for (int i = 0; i < SL; i += 2) {
                            uint32_t d0, d1;
                            unpack2u(sub2(pack2u(vr[i], vr[i + 1]), thr2), d0, d1);  // FADD2
                            mw[sl][i >> 5] = __funnelshift_l(d0, mw[sl][i >> 5], 1);
                            mw[sl][(i + 1) >> 5] = __funnelshift_l(d1, mw[sl][(i + 1) >> 5], 1);
                        }
                    }
                }
            }

            if constexpr (NCH == 2 && !ACC) {
                // Two chunks per tile share their TMEM buffers between the two groups, so the
                // waits above are only alias-free if no warp of this group runs a whole tile ahead
                // of another: a warp that reached tile seq + 2 while a sibling had not yet drained
                // tile seq would find `tempty` two phases behind, fall through both parity waits
                // and read tile seq's scores again (seen with >= 4 X stages: 32 wrong labels, then
                // a dead-locked barrier).  Once all four warps have drained this tile they may
                // drift apart again for the recheck.  (Train mode syncs the group per tile anyway.)
                bar_sync(1u + (uint32_t)g, 128);
            }

            // ---- resolve
            int nc = 0;
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int w = 0; w < NW; ++w) nc += __popc(mw[a][w]);
            const bool finite = fabsf(m_run) <= FLT_MAX;
            int label = kLabelFixup;
            if (finite && nc == 1) {
                int idx = 0;
#pragma unroll
                for (int a = 0; a < NS; ++a)
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const int cnt = (SL - 32 * w) < 32 ? (SL - 32 * w) : 32;
                        if (mw[a][w]) idx = a * SL + 32 * w + cnt - 32 + __clz(mw[a][w]);
                    }
                label = idx + 1;
            }
            bool flagged = finite && nc >= 2 && nc <= kMaxCand;  // kMaxCand <= 15
            const unsigned fmask0 = __ballot_sync(0xffffffffu, flagged);
            // Fast path (warp-uniform): every flagged row of this warp has exactly two candidates.
            // Each such thread settles its own row -- both fp32 distances in one walk over the row,
            // fp64 replica only if they are within the fp32 error of each other -- with no pair
            // list, no compaction and no exchange through shared memory.
            const bool duel_only = fmask0 != 0u && !__any_sync(0xffffffffu, flagged && nc != 2);
            if (duel_only) {
                if (flagged) {
                    int c0 = -1, c1 = -1;  // the two candidates, ascending node index
#pragma unroll
                    for (int a = 0; a < NS; ++a)
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            const int cw = (SL - 32 * w) < 32 ? (SL - 32 * w) : 32;
                            const int base = a * SL + 32 * w + cw - 32;
                            uint32_t m = mw[a][w];
                            if (m) {
                                const int lz = __clz(m);
                                m &= ~(0x80000000u >> lz);
                                if (c0 < 0) c0 = base + lz; else c1 = base + lz;
                                if (m) c1 = base + __clz(m);
                            }
                        }
                    if (c0 >= Ntot) c0 = 0;
                    if (c1 >= Ntot || c1 < 0) c1 = 0;
                    float d0, d1;
                    duel_dist2_f32(xs, ws, Ntot, nchunks16, row, c0, c1, d0, d1);
                    const float best = fminf(d0, d1);
                    const float bound = best * (1.0f + eps32) + 1.0e-30f;
                    ++st_flag;
                    st_pairs += 2;
                    if (d1 > bound) {
                        label = c0 + 1;
                    } else if (d0 > bound) {
                        label = c1 + 1;
                    } else {
                        // stage 3: fp64 replica of the reference loop, ascending node order, strict <
                        ++st_fp64;
                        label = kLabelFixup;
                        if (c0 < pl.K && c1 < pl.K) {
                            const double e0 = pair_dist_f64(xs, ws, Ntot, pl.C, row, c0);
                            const double e1 = pair_dist_f64(xs, ws, Ntot, pl.C, row, c1);
                            label = (e1 < e0 ? c1 : c0) + 1;
                            if (!(e0 == e0) || !(e1 == e1)) label = kLabelFixup;
                        }
                    }
                }
            } else if (fmask0) {
                // warp-local pair list: exclusive prefix of the candidate counts (<= 15, four bits)
                // of the flagged lanes from four ballots -- no dependent shuffle chain
                const int cntf = flagged ? nc : 0;
                const unsigned lt = (1u << lane) - 1u;
                const unsigned b0 = __ballot_sync(0xffffffffu, cntf & 1);
                const unsigned b1 = __ballot_sync(0xffffffffu, cntf & 2);
                const unsigned b2 = __ballot_sync(0xffffffffu, cntf & 4);
                const unsigned b3 = __ballot_sync(0xffffffffu, cntf & 8);
                const int pbase = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt) +
                                  8 * __popc(b3 & lt);
                if (flagged && pbase + nc > pair_cap) flagged = false;  // overflow: fix-up
                // slots [0, total) hold every pair that was written (overflowed lanes leave holes)
                int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
                if (total > pair_cap) total = pair_cap;
                if (flagged) {
                    int t = pbase;
#pragma unroll
                    for (int a = 0; a < NS; ++a)
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            const int cnt = (SL - 32 * w) < 32 ? (SL - 32 * w) : 32;
                            uint32_t m = mw[a][w];
                            while (m) {
                                const int lz = __clz(m);
                                m &= ~(0x80000000u >> lz);
                                pairs[t++] = ((uint32_t)row << 16) |
                                             (uint32_t)(a * SL + 32 * w + cnt - 32 + lz);
                            }
                        }
                    ++st_flag;
                    st_pairs += nc;
                }
                __syncwarp();
                // stage 2: fp32 distances of the warp's pairs, one pair per lane and pass
                // (lanes that overflowed leave holes in [0, total); holes are never read back)
                for (int pi = lane; pi < total; pi += 32) {
                    const uint32_t pr = pairs[pi];
                    const int prow = (int)(pr >> 16) & 127;
                    int pnode = (int)(pr & 0xFFFFu);
                    if (pnode >= Ntot) pnode = 0;
                    d2buf[pi] = pair_dist2_f32(xs, ws, Ntot, nchunks16, prow, pnode, (uint32_t)lane & 7u);
                }
                __syncwarp();
                if (flagged) {
                    float best = __int_as_float(0x7f800000);
#pragma unroll 1
                    for (int t = 0; t < nc; ++t) best = fminf(best, d2buf[pbase + t]);
                    const float bound = best * (1.0f + eps32) + 1.0e-30f;
                    int nsurv = 0, surv0 = -1;
#pragma unroll 1
                    for (int t = 0; t < nc; ++t)
                        if (d2buf[pbase + t] <= bound) {
                            if (nsurv == 0) surv0 = (int)(pairs[pbase + t] & 0xFFFFu);
                            ++nsurv;
                        }
                    if (nsurv == 1) {
                        label = surv0 + 1;
                    } else if (nsurv >= 2) {
                        // stage 3: fp64 replica of the reference loop over the survivors, in node
                        // index order (pairs were written in ascending node order)
                        ++st_fp64;
                        double bestd = DBL_MAX;
                        int bestk = -1;
#pragma unroll 1
                        for (int t = 0; t < nc; ++t) {
                            if (!(d2buf[pbase + t] <= bound)) continue;
                            const int k = (int)(pairs[pbase + t] & 0xFFFFu);
                            if (k >= pl.K) continue;
                            const double d = pair_dist_f64(xs, ws, Ntot, pl.C, row, k);
                            if (d < bestd) {
                                bestd = d;
                                bestk = k;
                            }
                        }
                        label = bestk >= 0 ? bestk + 1 : kLabelFixup;
                    }
                }
                __syncwarp();  // pair buffers are reused by the next tile
            }
            if (label > pl.K) label = kLabelFixup;  // a padded codebook row can only win on garbage
            if constexpr (ACC) {
                // Train mode resolves the rare rows the three stages could not settle right here
                // (their sums must be in this step's table): the warp runs the reference loop for
                // such a row cooperatively, lane l over nodes l, l+32, ...; the lexicographic
                // (distance, index) minimum over lanes is the first minimum of the sequential loop.
                unsigned fixm = __ballot_sync(0xffffffffu, label == kLabelFixup && grow < p.n);
                while (fixm) {
                    const int src = __ffs(fixm) - 1;
                    fixm &= fixm - 1;
                    const int frow = quad * 32 + src;
                    int minid = 0x7fffffff;
                    double mind = DBL_MAX;
                    for (int k = lane; k < pl.K; k += 32) {
                        const double d = pair_dist_f64(xs, ws, Ntot, pl.C, frow, k);
                        if (d < mind) {
                            mind = d;
                            minid = k;
                        }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, mind, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, minid, o);
                        if (od < mind || (od == mind && oi < minid)) {
                            mind = od;
                            minid = oi;
                        }
                    }
                    if (lane == src) {
                        label = minid == 0x7fffffff ? 0 : minid + 1;
                        ++st_fix;
                    }
                }
                if (lab_ptr != nullptr) {
                    if (grow < p.n)
                        *lab_ptr = label;
                    else if (p.compact_labels)
                        *lab_ptr = 0;  // padding row of the last tile: never counted
                }
            } else {
                if (grow < p.n) {
                    if (label == kLabelFixup) {
                        ++st_fix;
                        atomicAdd(&p.ctl->fixup_count, 1);
                    }
                    *lab_ptr = label;
                } else if (p.compact_labels) {
                    *lab_ptr = 0;  // padding row of the last tile: never counted
                }
            }
            if constexpr (do_acc) {
                // ---- fused per-node sums (deterministic, no atomics): see tile_accumulate_sorted
                tile_accumulate_sorted(smem, xs, pl.off_acc, pl.off_lab + (uint32_t)g * pl.sort_stride,
                                       pl.K, pl.C, pl.nblkX, g, quad, lane,
                                       (grow < p.n && label > 0) ? label - 1 : pl.K);
            }
            // all reads of this X stage by this warp are done
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8u * s);
            s += (uint32_t)NG;  // next tile of this group: NG stages further
            if (s >= (uint32_t)nstage) {
                s -= (uint32_t)nstage;
                ph ^= 1u;
            }
        }
    }
    // ================================================================ end of step st
    base_seq += cnt;
    if constexpr (ACC) {
        if (p.apply) {
            __syncthreads();
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                const uint64_t now = global_timer_ns();
                p.ctl->phase_ns[0] += now - step_t0;
            }
            step_update<NG>(p, st, smem, gb_target);
            if (blockIdx.x == 0 && threadIdx.x == 0) step_t0 = global_timer_ns();
        }
    }
    }  // for st

    if (warp < NEPI) {
        if (p.stats) {
            unsigned long long a = st_flag, b = st_pairs, c = st_fp64, d = st_fix;
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(~0u, a, o);
                b += __shfl_xor_sync(~0u, b, o);
                c += __shfl_xor_sync(~0u, c, o);
                d += __shfl_xor_sync(~0u, d, o);
            }
            if (lane == 0) {
                if (a) atomicAdd(p.stats + PIXIE_STAT_ROWS_FLAGGED, a);
                if (b) atomicAdd(p.stats + PIXIE_STAT_PAIRS, b);
                if (c) atomicAdd(p.stats + PIXIE_STAT_ROWS_FP64, c);
                if (d) atomicAdd(p.stats + PIXIE_STAT_ROWS_FIXUP, d);
            }
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == NEPI + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
    }
    if (ACC && !p.apply) {
        // ---- fused sums, part 2: groups are combined in group order into this CTA's partial,
        // then (grid barrier; every CTA is resident: grid <= SM count, one CTA per SM) each CTA
        // folds its slice of the table over all partials in CTA order, in fp64.
        const int len = pl.K * (pl.C + 1);
        const float *a = reinterpret_cast<const float *>(smem + pl.off_acc);
        float *mine = p.partials + (size_t)blockIdx.x * len;
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            float v = a[i];
#pragma unroll
            for (int gg = 1; gg < NG; ++gg) v += a[gg * len + i];
            mine[i] = v;
        }
        unsigned int *sync = p.ctl->sums_sync;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(&sync[0], 1u);
            unsigned spins = 0;
            while (atomicAdd(&sync[0], 0u) < gridDim.x) {
                __nanosleep(64);
                if (++spins > (1u << 24)) __trap();
            }
        }
        __syncthreads();
        __threadfence();
        {
            const int nparts = gridDim.x;
            const int per = (len + nparts - 1) / nparts;
            const int e0 = blockIdx.x * per;
            const int e1 = min(len, e0 + per);
            const int nthr = blockDim.x;           // multiple of 32
            const int oct = threadIdx.x >> 3, q = threadIdx.x & 7;
            const int rounds = (per + nthr / 8 - 1) / (nthr / 8);  // uniform trip count
            for (int it = 0; it < rounds; ++it) {
                const int e = e0 + it * (nthr / 8) + oct;
                // all loads of this thread first (independent, in flight together), then the sum
                // in ascending CTA order
                float v[kFoldMax];
#pragma unroll
                for (int u = 0; u < kFoldMax; ++u) {
                    const int pp = q + 8 * u;
                    v[u] = (e < e1 && pp < nparts) ? __ldcg(p.partials + (size_t)pp * len + e) : 0.f;
                }
                double acc = 0.0;
#pragma unroll
                for (int u = 0; u < kFoldMax; ++u) acc += (double)v[u];
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                if (e < e1 && q == 0) p.SN[e] = acc;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (atomicAdd(&sync[1], 1u) == gridDim.x - 1) {
                sync[0] = 0u;
                sync[1] = 0u;
                __threadfence();
            }
        }
    }
}

template <int SL, int SPC, int NCH, int NG, bool ACC>
static cudaError_t launch_variant(const CUtensorMap &tmX, const TcParams &p, int grid,
                                  cudaStream_t stream)
{
    constexpr int kThreads = NG * 128 + 64;
    cudaError_t e = cudaFuncSetAttribute(bmu_tc_kernel<SL, SPC, NCH, NG, ACC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.plan.smem_bytes);
    if (e != cudaSuccess) return e;
    if constexpr (ACC) {
        // the fused-sums variants use grid-wide barriers: launch cooperatively so the runtime
        // guarantees (or refuses) co-residency of all CTAs
        void *args[] = {const_cast<CUtensorMap *>(&tmX), const_cast<TcParams *>(&p)};
        e = cudaLaunchCooperativeKernel(
            reinterpret_cast<const void *>(&bmu_tc_kernel<SL, SPC, NCH, NG, ACC>),
            dim3(grid), dim3(kThreads), args, p.plan.smem_bytes, stream);
        count_launch();
        return e;
    } else {
        bmu_tc_kernel<SL, SPC, NCH, NG, ACC>
            <<<grid, kThreads, p.plan.smem_bytes, stream>>>(tmX, p);
        count_launch();
        return cudaGetLastError();
    }
}

// Body of a variant family's launcher: dispatch on the plan's template parameters.
#define PIXIE_VARIANT(a_, b_, c_, d_)                               \
    if (pl.SL == a_ && pl.spc == b_ && pl.NCH == c_ && pl.NG == d_) \
        return launch_variant<a_, b_, c_, d_, PIXIE_FAMILY_ACC>(tmX, p, grid, stream);
#define PIXIE_ALL_VARIANTS     \
    PIXIE_VARIANT(32, 1, 1, 4) \
    PIXIE_VARIANT(32, 2, 1, 4) \
    PIXIE_VARIANT(48, 2, 1, 4) \
    PIXIE_VARIANT(50, 2, 1, 4) \
    PIXIE_VARIANT(56, 2, 1, 4) \
    PIXIE_VARIANT(64, 2, 1, 4) \
    PIXIE_VARIANT(32, 1, 1, 2) \
    PIXIE_VARIANT(32, 2, 1, 2) \
    PIXIE_VARIANT(48, 2, 1, 2) \
    PIXIE_VARIANT(50, 2, 1, 2) \
    PIXIE_VARIANT(56, 2, 1, 2) \
    PIXIE_VARIANT(64, 2, 1, 2) \
    PIXIE_VARIANT(80, 2, 1, 2) \
    PIXIE_VARIANT(100, 2, 1, 2) \
    PIXIE_VARIANT(104, 2, 1, 2) \
    PIXIE_VARIANT(128, 2, 1, 2) \
    PIXIE_VARIANT(80, 2, 2, 2) \
    PIXIE_VARIANT(100, 2, 2, 2) \
    PIXIE_VARIANT(104, 2, 2, 2) \
    PIXIE_VARIANT(128, 2, 2, 2)

}  // namespace pixie
