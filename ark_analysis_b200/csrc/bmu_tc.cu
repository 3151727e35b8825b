// bmu_tc.cu -- the tensor-core best-matching-unit (BMU) kernel for sm_100a.
//
// Replaces the N x K x C scalar fp64 loop of pyFlowSOM's C_mapDataToCodes (call site
// /root/reference/src/ark/phenotyping/cluster_helpers.py:152-157) with a three-stage search whose
// RESULT is bit-identical to that loop:
//
//   stage 1 (tensor cores, every row x every node): score[i,k] = ||w_k||^2 - 2 x_i.w_k as one
//           tcgen05.mma.kind::tf32 contraction per 128-row tile.  X tiles arrive by TMA straight
//           into the SWIZZLE_128B K-major layout the MMA consumes (the fp32 bits ARE the tf32
//           operand; no conversion pass), the codebook image (-2 W plus a bias K-step carrying
//           ||w||^2 split into three tf32-exact terms) is resident in shared memory, accumulators
//           live in TMEM.  The epilogue reads each row's K scores with tcgen05.ld, takes the row
//           minimum and keeps every node within a proven error bound delta of it.
//   stage 2 (CUDA cores, only rows with >= 2 candidates): fp32 squared distances of the candidate
//           (row, node) pairs, ~2^-19 relative error; keeps the nodes that can still be the minimum.
//   stage 3 (CUDA cores, fp64, rare): the reference's own operation sequence (subtract, multiply,
//           add in channel order, sqrt, strict <, lowest index first) on the survivors.
//   Rows that defeat this (NaN/Inf, > kMaxCand candidates, pair-buffer overflow) get a sentinel
//   label and are resolved by the exact fp64 kernel (bmu_exact.cu) launched right behind.
//
// Warp roles (320 threads, 1 CTA per SM, persistent over tiles):
//   warps 0-3 / 4-7 : two epilogue groups, alternating tiles (thread <-> tile row <-> TMEM lane)
//   warp 8          : TMA producer (one elected lane)
//   warp 9          : TMEM allocator + MMA issuer (one elected lane)
#include <float.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pixie {

using namespace ptx;

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
TcPlan make_tc_plan(int C, int K)
{
    TcPlan p{};
    p.ok = false;
    p.C = C;
    p.K = K;
    if (C < 1 || C > 128 || K < 1 || K > 512) return p;
    p.C8 = (C + 7) / 8 * 8;
    p.ksteps = p.C8 / 8;
    p.nblkX = (C + 31) / 32;
    p.nblkW = (p.C8 + 8 + 31) / 32;
    static const int kSL[] = {32, 64, 80, 96, 104, 112, 128};
    long best = -1;
    for (int SL : kSL)
        for (int spc = 1; spc <= 8; ++spc) {
            int Nmma = SL * spc;
            if (Nmma > 256 || Nmma % 16) continue;
            for (int NCH = 1; NCH <= 2; ++NCH) {
                int Ntot = NCH * Nmma;
                if (Ntot < K || Ntot > 512) continue;
                long cost = (long)Ntot * 64 + NCH * spc * 8 + NCH;
                if (best < 0 || cost < best) {
                    best = cost;
                    p.SL = SL;
                    p.spc = spc;
                    p.NCH = NCH;
                    p.Nmma = Nmma;
                    p.Ntot = Ntot;
                }
            }
        }
    if (best < 0) return p;
    p.nbuf = 2;
    int need = 2 * p.Nmma;
    p.tmem_cols = 32;
    while (p.tmem_cols < need) p.tmem_cols *= 2;
    p.stage_bytes = (uint32_t)p.nblkX * 16384u;
    p.wimg_bytes = (uint32_t)p.nblkW * (uint32_t)p.Ntot * 128u;
    p.off_ones = p.wimg_bytes;
    p.off_x = p.off_ones + 4096u;
    const uint32_t scratch = 256u + 256u * kMaxCand * 2u + 2u * kPairCap * 4u + 2u * kPairCap * 4u;
    const uint32_t limit = 227u * 1024u - 1024u;
    if (p.off_x + scratch + 2u * p.stage_bytes > limit) return p;
    p.nstage = (int)((limit - p.off_x - scratch) / p.stage_bytes);
    if (p.nstage > kMaxStages) p.nstage = kMaxStages;
    // even depth: stage (it % nstage) is then always consumed by the same epilogue group (it % 2),
    // so no consumer can reach a full-barrier wait one phase early (parity aliasing).
    p.nstage &= ~1;
    p.off_bar = p.off_x + (uint32_t)p.nstage * p.stage_bytes;
    p.off_cand = p.off_bar + 256u;
    p.off_pairs = p.off_cand + 256u * kMaxCand * 2u;
    p.off_d2 = p.off_pairs + 2u * kPairCap * 4u;
    p.smem_bytes = p.off_d2 + 2u * kPairCap * 4u + 1024u;
    p.ok = true;
    return p;
}

// ------------------------------------------------------------------------------------------------
// codebook preparation: W [K x C] fp32  ->  shared-memory image + norms
// ------------------------------------------------------------------------------------------------
// Image layout (bytes): block b (32 columns) at b * Ntot * 128, row r at r * 128, 16-byte chunk c at
// ((c ^ (r & 7)) << 4) -- exactly what TMA SWIZZLE_128B would have written, so one linear
// cp.async.bulk brings it into place.  Column j < C of row k holds -2 * W[k, j]; columns
// C8, C8+1, C8+2 hold ||w_k||^2 split into three tf32-exact terms (multiplied by the all-ones A
// tile of the bias K-step); rows >= K hold zeros and a huge bias so they never win.
__device__ __forceinline__ uint32_t img_offset(int Ntot, int row, int col)
{
    const int b = col >> 5, cc = col & 31;
    return (uint32_t)b * (uint32_t)Ntot * 128u + (uint32_t)row * 128u +
           (uint32_t)((((cc >> 2) ^ (row & 7)) << 4) + ((cc & 3) << 2));
}

__global__ void codebook_prep_kernel(const float *__restrict__ W, int K, int C, int C8, int nblkW,
                                     int Ntot, float *__restrict__ wimg,
                                     CodebookAux *__restrict__ aux)
{
    __shared__ float s_max[32];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const int ncols = nblkW * 32;
    float local_max = 0.f;
    for (int row = threadIdx.x; row < Ntot; row += blockDim.x) {
        double nrm2 = 0.0;
        bool bad = false;
        for (int col = 0; col < ncols; ++col) {
            float v = 0.f;
            if (row < K && col < C) {
                const float w = W[(size_t)row * C + col];
                if (!(fabsf(w) <= FLT_MAX)) bad = true;
                nrm2 += (double)w * (double)w;
                v = -2.0f * w;
            }
            if (col < C8 || col >= C8 + 3)
                *reinterpret_cast<float *>(reinterpret_cast<char *>(wimg) +
                                           img_offset(Ntot, row, col)) = v;
        }
        float bias = (row < K) ? (float)nrm2 : 1.0e30f;
        if (!(bias <= FLT_MAX)) bias = FLT_MAX;  // overflowed norms: row can never win anyway
        // three tf32-exact pieces (11 significant bits each): h + m + l == bias exactly
        const float h = __uint_as_float(__float_as_uint(bias) & 0xFFFFE000u);
        const float r1 = bias - h;
        const float m = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
        const float l = r1 - m;
        char *base = reinterpret_cast<char *>(wimg);
        *reinterpret_cast<float *>(base + img_offset(Ntot, row, C8 + 0)) = h;
        *reinterpret_cast<float *>(base + img_offset(Ntot, row, C8 + 1)) = m;
        *reinterpret_cast<float *>(base + img_offset(Ntot, row, C8 + 2)) = l;
        if (row < K) {
            if (bad) atomicOr(&s_bad, 1);
            const float nr = (float)sqrt(nrm2);
            if (nr <= FLT_MAX) local_max = fmaxf(local_max, nr);
        }
    }
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(~0u, local_max, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = local_max;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, s_max[i]);
        mx *= 1.0000005f;
        aux->wmax = mx;
        aux->wmax2 = mx * mx;
        aux->nonfinite = s_bad;
        aux->pad = 0;
    }
}

cudaError_t launch_codebook_prep(const float *W, int K, int C, const TcPlan &plan, float *wimg,
                                 CodebookAux *aux, cudaStream_t stream)
{
    codebook_prep_kernel<<<1, 256, 0, stream>>>(W, K, C, plan.C8, plan.nblkW, plan.Ntot, wimg, aux);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// device helpers shared by the epilogue stages
// ------------------------------------------------------------------------------------------------
// fp32 value of channel `col` of tile row `row` in an X stage (TMA SWIZZLE_128B layout).
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

// stage 2: fp32 squared distance between tile row `row` and codebook node `node`.
__device__ __forceinline__ float pair_dist2_f32(const uint8_t *xs, const uint8_t *ws, int Ntot,
                                                int nchunks16, int row, int node)
{
    float acc0 = 0.f, acc1 = 0.f;
    for (int q = 0; q < nchunks16; ++q) {
        const float4 x = *x_chunk_ptr(xs, row, q >> 3, q & 7);
        const float4 w = *w_chunk_ptr(ws, Ntot, node, q >> 3, q & 7);
        const float d0 = fmaf(0.5f, w.x, x.x), d1 = fmaf(0.5f, w.y, x.y);  // x - w, w = -0.5 w'
        const float d2 = fmaf(0.5f, w.z, x.z), d3 = fmaf(0.5f, w.w, x.w);
        acc0 = fmaf(d0, d0, acc0);
        acc1 = fmaf(d1, d1, acc1);
        acc0 = fmaf(d2, d2, acc0);
        acc1 = fmaf(d3, d3, acc1);
    }
    return acc0 + acc1;
}

// stage 3: the reference's fp64 operation sequence for one (row, node) pair
// (oracle/pixie_oracle.c nearest_node): tmp = x - w; acc = acc + tmp * tmp (separately rounded),
// in channel order; d = sqrt(acc).
__device__ __noinline__ double pair_dist_f64(const uint8_t *xs, const uint8_t *ws, int Ntot, int C,
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

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int SL>
__global__ void __launch_bounds__(320, 1)
bmu_tc_kernel(const __grid_constant__ CUtensorMap tmX, const TcParams p)
{
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
    const uint32_t bar_full = bar0;                          // [kMaxStages]
    const uint32_t bar_empty = bar0 + 8u * kMaxStages;       // [kMaxStages]
    const uint32_t bar_tfull = bar0 + 16u * kMaxStages;      // [2]
    const uint32_t bar_tempty = bar_tfull + 16u;             // [2]
    const uint32_t bar_w = bar_tempty + 16u;                 // codebook image landed
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + pl.off_bar + 8u * (2 * kMaxStages + 5));
    int *pair_count = reinterpret_cast<int *>(smem + pl.off_bar + 8u * (2 * kMaxStages + 6));  // [2 groups][2 parities]
    uint16_t *cand_all = reinterpret_cast<uint16_t *>(smem + pl.off_cand);
    uint32_t *pairs_all = reinterpret_cast<uint32_t *>(smem + pl.off_pairs);
    float *d2_all = reinterpret_cast<float *>(smem + pl.off_d2);

    const int nstage = pl.nstage;
    const int64_t ntiles = p.ntiles;

    // ---------------------------------------------------------------- one-time setup
    if (warp == 8 && lane == 0) {
        prefetch_tensormap(&tmX);
        for (int s = 0; s < nstage; ++s) {
            mbar_init(bar_full + 8u * s, 1);   // producer's arrive.expect_tx
            mbar_init(bar_empty + 8u * s, 4);  // one arrive per warp of the owning epilogue group
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8u * b, 1);   // tcgen05.commit
            mbar_init(bar_tempty + 8u * b, 4);  // one arrive per epilogue warp
        }
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    if (warp == 9) {
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), (uint32_t)pl.tmem_cols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<float *>(ones)[i] = 1.0f;
    if (threadIdx.x < 4) pair_count[threadIdx.x] = 0;
    fence_proxy_async();  // the ones tile is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ============================================================ TMA producer
        if (lane == 0) {
            // codebook image: linear bulk copies (image is pre-swizzled in global memory)
            mbar_arrive_expect_tx(bar_w, pl.wimg_bytes);
            for (uint32_t off = 0; off < pl.wimg_bytes; off += 16384u) {
                const uint32_t sz = min(16384u, pl.wimg_bytes - off);
                bulk_load(sbase + off, reinterpret_cast<const uint8_t *>(p.wimg) + off, sz, bar_w);
            }
            int64_t it = 0;
            for (int64_t j = blockIdx.x; j < ntiles; j += gridDim.x, ++it) {
                const int s = (int)(it % nstage);
                const uint32_t ph = (uint32_t)((it / nstage) & 1);
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8u * s, pl.stage_bytes);
                const int64_t tile = p.tile_first + j * p.tile_stride;
                const int32_t row0 = (int32_t)(tile * kTile);
                for (int b = 0; b < pl.nblkX; ++b)
                    tma_load_2d(sbase + pl.off_x + (uint32_t)s * pl.stage_bytes + (uint32_t)b * 16384u,
                                &tmX, bar_full + 8u * s, b * 32, row0, kEvictFirst);
            }
        }
    } else if (warp == 9) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)pl.Nmma);
            const uint64_t desc_ones = umma_desc_nosw(sbase + pl.off_ones, 128u, 256u);
            // bias K-step: columns C8..C8+7 of the codebook image
            const uint32_t bias_blk = (uint32_t)(pl.C8 >> 5), bias_off = (uint32_t)(pl.C8 & 31) * 4u;
            mbar_wait(bar_w, 0);
            int64_t it = 0;
            for (int64_t j = blockIdx.x; j < ntiles; j += gridDim.x, ++it) {
                const int s = (int)(it % nstage);
                const uint32_t ph = (uint32_t)((it / nstage) & 1);
                mbar_wait(bar_full + 8u * s, ph);
                const uint32_t xs_addr = sbase + pl.off_x + (uint32_t)s * pl.stage_bytes;
                for (int c = 0; c < pl.NCH; ++c) {
                    const int64_t q = it * pl.NCH + c;
                    const int buf = (int)(q & 1);
                    const uint32_t bph = (uint32_t)((q >> 1) & 1);
                    mbar_wait(bar_tempty + 8u * buf, bph ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * pl.Nmma);
                    const uint32_t wrow = (uint32_t)(c * pl.Nmma) * 128u;
                    for (int ks = 0; ks < pl.ksteps; ++ks) {
                        const uint32_t blk = (uint32_t)(ks >> 2), ko = (uint32_t)(ks & 3) * 32u;
                        const uint64_t da = umma_desc_sw128(xs_addr + blk * 16384u + ko);
                        const uint64_t db =
                            umma_desc_sw128(sbase + blk * (uint32_t)pl.Ntot * 128u + wrow + ko);
                        mma_tf32(d_tmem, da, db, idesc, ks > 0 ? 1u : 0u);
                    }
                    const uint64_t dbias = umma_desc_sw128(
                        sbase + bias_blk * (uint32_t)pl.Ntot * 128u + wrow + bias_off);
                    mma_tf32(d_tmem, desc_ones, dbias, idesc, 1u);
                    mma_commit(bar_tfull + 8u * buf);
                }
            }
        }
    } else {
        // ============================================================ epilogue groups
        const int g = warp >> 2;            // group 0/1
        const int quad = warp & 3;          // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;   // tile row == TMEM lane
        const int gtid = threadIdx.x & 127; // thread index within the group
        const uint32_t bar_id = 1u + (uint32_t)g;
        uint16_t *my_cand = cand_all + (size_t)(g * 128 + row) * kMaxCand;
        uint32_t *pairs = pairs_all + (size_t)g * kPairCap;
        float *d2buf = d2_all + (size_t)g * kPairCap;
        const float wmax = p.aux->wmax, wmax2 = p.aux->wmax2;
        const int nchunks16 = pl.C8 >> 2;   // 16-byte chunks holding real channels (C8 / 4)
        const float eps32 = (float)(pl.C + 8) * 2.4e-7f;
        unsigned long long st_flag = 0, st_pairs = 0, st_fp64 = 0, st_fix = 0;
        mbar_wait(bar_w, 0);  // codebook image visible to this thread (stages 2/3 read it)

        int64_t it = g;
        for (int64_t j = (int64_t)blockIdx.x + (int64_t)g * gridDim.x; j < ntiles;
             j += 2 * (int64_t)gridDim.x, it += 2) {
            const int s = (int)(it % nstage);
            const uint32_t ph = (uint32_t)((it / nstage) & 1);
            const uint8_t *xs = xs0 + (size_t)s * pl.stage_bytes;
            const int64_t tile = p.tile_first + j * p.tile_stride;
            const int64_t grow = tile * kTile + row;  // global row
            const int par = (int)((it >> 1) & 1);
            int *my_pair_count = pair_count + g * 2 + par;

            mbar_wait(bar_full + 8u * s, ph);  // X tile landed

            // ---- per-row error bound of the tf32 scores
            float xn2 = 0.f;
            for (int q = 0; q < nchunks16; ++q) {
                const float4 x = *x_chunk_ptr(xs, row, q >> 3, q & 7);
                xn2 = fmaf(x.x, x.x, xn2);
                xn2 = fmaf(x.y, x.y, xn2);
                xn2 = fmaf(x.z, x.z, xn2);
                xn2 = fmaf(x.w, x.w, xn2);
            }
            // |score error| <= 2^-8 (1+1/16) ||x|| wmax + 2^-18 wmax^2 ; delta = 2 x that
            const float delta =
                2.0f * (0.00415039f * sqrtf(xn2) * 1.000001f * wmax + 3.8147e-6f * wmax2);

            float m_run = __int_as_float(0x7f800000);
            int ncand = 0;
            int cand0 = 0;

            for (int c = 0; c < pl.NCH; ++c) {
                const int64_t q = it * pl.NCH + c;
                const int buf = (int)(q & 1);
                const uint32_t bph = (uint32_t)((q >> 1) & 1);
                // With two chunks per tile both groups alternate on the same accumulator buffer, so
                // this group may get here a whole phase early, where a parity wait would alias and
                // fall through.  Waiting first for the previous use's release (made by the OTHER
                // group, after it saw the previous commit) pins the barrier to the right phase.
                if (pl.NCH > 1 && q >= 2) mbar_wait(bar_tempty + 8u * buf, bph ^ 1u);
                mbar_wait(bar_tfull + 8u * buf, bph);
                tc_fence_after();
                for (int sidx = 0; sidx < pl.spc; ++sidx) {
                    uint32_t vr[SL];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                           (uint32_t)(buf * pl.Nmma + sidx * SL);
                    tmem_ld_cols<SL>(taddr, vr);
                    tc_wait_ld();
                    if (sidx == pl.spc - 1) {
                        // accumulator buffer fully read: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8u * buf);
                    }
                    // pass 1: slice minimum
                    float ms = __uint_as_float(vr[0]);
#pragma unroll
                    for (int i = 1; i < SL; ++i) ms = fminf(ms, __uint_as_float(vr[i]));
                    const float m_new = fminf(m_run, ms);
                    if (m_new + delta < m_run) ncand = 0;  // earlier candidates are out of range
                    m_run = m_new;
                    const float thr = m_run + delta;
                    if (__any_sync(0xffffffffu, ms < thr)) {
                        // pass 2: sign bit of (v - thr) funnel-shifted into a bit mask
                        constexpr int NW = (SL + 31) / 32;
                        uint32_t mw[NW];
#pragma unroll
                        for (int w = 0; w < NW; ++w) mw[w] = 0u;
#pragma unroll
                        for (int i = 0; i < SL; ++i) {
                            const float d = __uint_as_float(vr[i]) - thr;
                            mw[i >> 5] = __funnelshift_l(__float_as_uint(d), mw[i >> 5], 1);
                        }
                        const int colbase = c * pl.Nmma + sidx * SL;
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            const int cnt = (SL - 32 * w) < 32 ? (SL - 32 * w) : 32;
                            uint32_t m = mw[w];
                            while (m) {
                                const int b = 31 - __clz(m);
                                m &= ~(1u << b);
                                const int idx = colbase + 32 * w + (cnt - 1 - b);
                                if (ncand == 0) cand0 = idx;
                                if (ncand < kMaxCand) my_cand[ncand] = (uint16_t)idx;
                                ++ncand;
                            }
                        }
                    }
                }
            }

            // ---- resolve
            int label = kLabelFixup;
            int pbase = -1;
            const bool finite = fabsf(m_run) <= FLT_MAX;
            if (finite && ncand == 1) {
                label = cand0 + 1;
            } else if (finite && ncand >= 2 && ncand <= kMaxCand) {
                const int base = atomicAdd(my_pair_count, ncand);
                if (base + ncand <= kPairCap) {
                    pbase = base;
                    for (int t = 0; t < ncand; ++t)
                        pairs[base + t] = ((uint32_t)row << 16) | (uint32_t)my_cand[t];
                    ++st_flag;
                    st_pairs += ncand;
                }
            }
            bar_sync(bar_id, 128);  // pairs of this tile are complete
            if (gtid == 0) pair_count[g * 2 + (par ^ 1)] = 0;  // reset the next tile's counter
            {
                int P = *reinterpret_cast<volatile int *>(my_pair_count);
                if (P > kPairCap) P = kPairCap;  // overflowing rows did not write their pairs
                // NOTE: rows that overflowed still bumped the counter; their range is unwritten and
                // no row owns it, so evaluating stale pairs there is harmless (indices are masked).
                for (int pi = gtid; pi < P; pi += 128) {
                    const uint32_t pr = pairs[pi];
                    const int prow = (int)(pr >> 16) & 127;
                    int pnode = (int)(pr & 0xFFFFu);
                    if (pnode >= pl.Ntot) pnode = 0;
                    d2buf[pi] = pair_dist2_f32(xs, ws, pl.Ntot, nchunks16, prow, pnode);
                }
            }
            bar_sync(bar_id, 128);  // stage-2 distances are complete
            if (pbase >= 0) {
                float best = __int_as_float(0x7f800000);
                for (int t = 0; t < ncand; ++t) best = fminf(best, d2buf[pbase + t]);
                const float bound = best * (1.0f + eps32) + 1.0e-30f;
                int nsurv = 0, surv0 = -1;
                for (int t = 0; t < ncand; ++t)
                    if (d2buf[pbase + t] <= bound) {
                        if (nsurv == 0) surv0 = (int)my_cand[t];
                        ++nsurv;
                    }
                if (nsurv == 1 && surv0 < pl.K) {
                    label = surv0 + 1;
                } else if (nsurv >= 2) {
                    // stage 3: fp64 replica of the reference loop over the survivors, index order
                    ++st_fp64;
                    double bestd = DBL_MAX;
                    int bestk = -1;
                    for (int t = 0; t < ncand; ++t) {
                        if (!(d2buf[pbase + t] <= bound)) continue;
                        const int k = (int)my_cand[t];
                        if (k >= pl.K) continue;
                        const double d = pair_dist_f64(xs, ws, pl.Ntot, pl.C, row, k);
                        if (d < bestd) {
                            bestd = d;
                            bestk = k;
                        }
                    }
                    label = bestk >= 0 ? bestk + 1 : kLabelFixup;
                }
            }
            if (label == kLabelFixup || label > pl.K) {
                label = kLabelFixup;
                if (grow < p.n) {
                    ++st_fix;
                    atomicAdd(p.fixup_count, 1);
                }
            }
            if (grow < p.n) {
                if (p.compact_labels)
                    p.labels[j * kTile + row] = label;
                else
                    p.labels[grow] = label;
            } else if (p.compact_labels) {
                p.labels[j * kTile + row] = 0;  // padding row of the last tile: never counted
            }
            // all reads of this X stage (and of the pair buffers) by this warp are done
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8u * s);
        }

        if (p.stats) {
            for (int o = 16; o > 0; o >>= 1) {
                st_flag += __shfl_xor_sync(~0u, st_flag, o);
                st_pairs += __shfl_xor_sync(~0u, st_pairs, o);
                st_fp64 += __shfl_xor_sync(~0u, st_fp64, o);
                st_fix += __shfl_xor_sync(~0u, st_fix, o);
            }
            if (lane == 0) {
                if (st_flag) atomicAdd(p.stats + PIXIE_STAT_ROWS_FLAGGED, st_flag);
                if (st_pairs) atomicAdd(p.stats + PIXIE_STAT_PAIRS, st_pairs);
                if (st_fp64) atomicAdd(p.stats + PIXIE_STAT_ROWS_FP64, st_fp64);
                if (st_fix) atomicAdd(p.stats + PIXIE_STAT_ROWS_FIXUP, st_fix);
            }
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
    }
}

template <int SL>
static cudaError_t launch_one(const CUtensorMap &tmX, const TcParams &p, int grid,
                              cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(bmu_tc_kernel<SL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.plan.smem_bytes);
    if (e != cudaSuccess) return e;
    bmu_tc_kernel<SL><<<grid, 320, p.plan.smem_bytes, stream>>>(tmX, p);
    return cudaGetLastError();
}

cudaError_t launch_bmu_tc(const CUtensorMap &tmX, const TcParams &p, int num_sms,
                          cudaStream_t stream)
{
    if (p.ntiles <= 0) return cudaSuccess;
    int grid = num_sms;
    if ((int64_t)grid > p.ntiles) grid = (int)p.ntiles;
    switch (p.plan.SL) {
        case 32: return launch_one<32>(tmX, p, grid, stream);
        case 64: return launch_one<64>(tmX, p, grid, stream);
        case 80: return launch_one<80>(tmX, p, grid, stream);
        case 96: return launch_one<96>(tmX, p, grid, stream);
        case 104: return launch_one<104>(tmX, p, grid, stream);
        case 112: return launch_one<112>(tmX, p, grid, stream);
        case 128: return launch_one<128>(tmX, p, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace pixie
