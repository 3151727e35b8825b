// bmu_tc_kernel.cuh -- the tensor-core BMU kernel template (see bmu_tc.cu for the overview).  Included
// by the bmu_tc_inst_*.cu translation units, each of which instantiates one family of variants.
#pragma once
#include <float.h>
#include <stdio.h>

#include "common.cuh"
#include "ptx.cuh"
#include "som_update.cuh"

// Tile-phase cycle counters of warp 0 of CTA 0 (diagnostic builds: make prof).  They live in
// registers and reach the control block once per step, so that reading them does not stall the
// phases they measure.
#ifdef PIXIE_PROFILE
#define PIXIE_PROF_DECL() long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tick_ = 0; unsigned int trace_n_ = 0
#define PIXIE_TICK(i) tick_ = clock64()
#define PIXIE_TOCK(i)                     \
    do {                                  \
        const long long now_ = clock64(); \
        prof_[i] += now_ - tick_;         \
        tick_ = now_;                     \
    } while (0)
#define PIXIE_TILE_DONE() prof_[6] += 1
#define PIXIE_PROF_FLUSH()                                                          \
    do {                                                                            \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                                  \
            for (int i_ = 0; i_ < 8; ++i_) {                                        \
                p.ctl->tile_cyc[i_] += (unsigned long long)prof_[i_];               \
                prof_[i_] = 0;                                                      \
            }                                                                       \
        }                                                                           \
    } while (0)
// event trace of CTA 0 (one lane per warp): {step, warp, event, tile seq} + globaltimer
#define PIXIE_TRACE(ev, seq_)                                                                   \
    do {                                                                                        \
        if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && st >= p.dbg_step0 &&       \
            st < p.dbg_step0 + 2 && trace_n_ < 1024u) {                                         \
            /* every warp owns 1024 slots: no atomics, the two stores are fire-and-forget */    \
            unsigned long long *slot_ = p.trace + 2u * ((threadIdx.x >> 5) * 1024u + trace_n_); \
            slot_[0] = ((unsigned long long)st << 48) | ((unsigned long long)(threadIdx.x >> 5) << 40) | \
                       ((unsigned long long)(ev) << 32) | (unsigned int)(seq_);                 \
            slot_[1] = global_timer_ns();                                                       \
            ++trace_n_;                                                                         \
        }                                                                                       \
    } while (0)
#else
#define PIXIE_TRACE(ev, seq_)
#define PIXIE_PROF_DECL()
#define PIXIE_TICK(i)
#define PIXIE_TOCK(i)
#define PIXIE_TILE_DONE()
#define PIXIE_PROF_FLUSH()
#endif

namespace pixie {

using namespace ptx;

#ifdef PIXIE_PROFILE
constexpr unsigned int kTraceCap = 1u << 15;
#endif

// Image layout (bytes): block b (32 columns) at b * Ntot * 128, row r at r * 128, 16-byte chunk c at
// ((c ^ (r & 7)) << 4) -- exactly what TMA SWIZZLE_128B would have written, so one linear
// cp.async.bulk brings it into place.  Column j < C of row k holds -2 * W[k, j].  Behind the
// blocks sits the bias block (plan.off_bias): one 8-column K-step per row in the no-swizzle
// core-matrix layout (bias_offset), columns 0..2 = ||w_k||^2 split into three tf32-exact terms
// (multiplied by the all-ones A tile of the bias K-step); rows >= K hold zeros and a huge bias so
// they never win.
// `lay` = rows per image block (Ntot, < 65536) in the low half and, when the last block is a
// "tail8" block (TcPlan::tail8: 32-byte rows, SWIZZLE_32B), its index + 1 in the high half.
__host__ __device__ __forceinline__ int lay_pack(int Ntot, int tail_blk_or_neg)
{
    return Ntot | ((tail_blk_or_neg + 1) << 16);
}
__device__ __forceinline__ uint32_t img_offset(int lay, int row, int col)
{
    const int Ntot = lay & 0xFFFF, tb = (lay >> 16) - 1;
    const int b = col >> 5, cc = col & 31;
    if (b == tb)
        return (uint32_t)b * (uint32_t)Ntot * 128u + (uint32_t)row * 32u +
               (uint32_t)((((cc >> 2) ^ ((row >> 2) & 1)) << 4) + ((cc & 3) << 2));
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
template <bool T8>
__device__ __forceinline__ const float4 *x_chunk_ptr(const uint8_t *xs, int lay, int row, int blk,
                                                     int chunk)
{
    if (T8 && blk == (lay >> 16) - 1)
        return reinterpret_cast<const float4 *>(xs + (size_t)blk * 16384u + (size_t)row * 32u +
                                                (size_t)(((chunk ^ (row >> 2)) & 1) << 4));
    return reinterpret_cast<const float4 *>(xs + (size_t)blk * 16384u + (size_t)row * 128u +
                                            (size_t)(((chunk ^ (row & 7)) & 7) << 4));
}
template <bool T8>
__device__ __forceinline__ const float4 *w_chunk_ptr(const uint8_t *ws, int lay, int node, int blk,
                                                     int chunk)
{
    const int Ntot = lay & 0xFFFF;
    if (T8 && blk == (lay >> 16) - 1)
        return reinterpret_cast<const float4 *>(ws + (size_t)blk * (size_t)Ntot * 128u +
                                                (size_t)node * 32u +
                                                (size_t)(((chunk ^ (node >> 2)) & 1) << 4));
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
template <bool T8 = false>
__device__ __forceinline__ float pair_dist2_f32(const uint8_t *xs, const uint8_t *ws, int lay,
                                                int nchunks16, int row, int node, uint32_t rot)
{
    const int Ntot = lay & 0xFFFF;
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
        const float4 x = *x_chunk_ptr<T8>(xs, lay, row, nfull, lc);
        const float4 w = *w_chunk_ptr<T8>(ws, lay, node, nfull, lc);
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
template <bool T8 = false>
__device__ __forceinline__ void duel_dist2_f32(const uint8_t *xs, const uint8_t *ws, int lay,
                                               int nchunks16, int row, int na, int nb, float &da,
                                               float &db)
{
    const int Ntot = lay & 0xFFFF;
    constexpr bool tail = T8;
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
            // partial last block; a tail8 block keeps 32-byte rows (chunk c at c ^ ((r >> 2) & 1))
            uint32_t xr = xrow, ra = wa, rb = wb;
            if (tail) {
                xr = (xrow & ~0x7Fu) - (uint32_t)row * 96u + ((((uint32_t)row >> 2) & 1u) << 4);
                ra = (wa & ~0x7Fu) - (uint32_t)na * 96u + ((((uint32_t)na >> 2) & 1u) << 4);
                rb = (wb & ~0x7Fu) - (uint32_t)nb * 96u + ((((uint32_t)nb >> 2) & 1u) << 4);
            }
            for (uint32_t c = 0; c < (uint32_t)left; ++c) {
                const uint4 xu = lds128(xr ^ (c << 4));
                const uint4 au = lds128(ra ^ (c << 4));
                const uint4 bu = lds128(rb ^ (c << 4));
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
template <bool T8 = false>
static __device__ __noinline__ double pair_dist_f64(const uint8_t *xs, const uint8_t *ws, int lay, int C,
                                             int row, int node)
{
    double acc = 0.0;
    for (int j = 0; j < C; ++j) {
        const int blk = j >> 5, cc = j & 31;
        const float xf = reinterpret_cast<const float *>(x_chunk_ptr<T8>(xs, lay, row, blk, cc >> 2))[cc & 3];
        const float wf = reinterpret_cast<const float *>(w_chunk_ptr<T8>(ws, lay, node, blk, cc >> 2))[cc & 3];
        const double tmp = __dsub_rn((double)xf, (double)(-0.5f * wf));
        acc = __dadd_rn(acc, __dmul_rn(tmp, tmp));
    }
    return __dsqrt_rn(acc);
}

// ------------------------------------------------------------------------------------------------
// train mode (ACC variants): synchronisation
// ------------------------------------------------------------------------------------------------
// The TMA producer warp never takes part in the end-of-step work: X does not depend on the
// codebook, so it keeps streaming the NEXT step's tiles into free stages while the other warps
// (epilogue groups + MMA warp, `nthr` threads) fold, exchange and update.  Those warps meet at
// named barrier kBarStep.
constexpr uint32_t kBarStep = 8;

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier (every CTA is resident: cooperative launch, grid <= SM count).  `counter` only
// grows during a launch; `target` is its value once every CTA has arrived at this instance.  One
// thread per CTA talks to L2; the CTA barriers on both sides make its fences cumulative for the
// whole CTA (the scheme cooperative groups uses).
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target,
                                             uint32_t nthr, bool leader)
{
    bar_sync(kBarStep, nthr);
    if (leader) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned spins = 0;
        uint64_t t0 = 0;
        while (ld_acquire_gpu(counter) < target) {
            if ((++spins & 1023u) == 0u) {  // a wedged grid traps after ~4 s instead of hanging
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > 4000000000ull) __trap();
            }
        }
        __threadfence();
    }
    bar_sync(kBarStep, nthr);
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
// train mode: the per-node statistics.  Each epilogue group owns a private table of channel sums
// -- in shared memory ([K][tab_pitch(C)] fp32) when NG of them fit beside the pipeline, else in
// global memory (L2-resident: [K][part_pitch(C)], group g of CTA b is table b * NG + g of
// p.partials) -- and an int32 count per node in shared memory.
//
// publish_labels: the tile's labels go to the group's ring (two slots: a warp may start the next
// tile while a sibling still reads this one's) and every row counts itself (integer shared-memory
// atomic: order-free); then the four warps of the group meet.
//
// tile_accumulate: warp `quad` owns the nodes with (node & 3) == quad.  It compacts the tile's rows
// of those nodes (ascending row order) into a list and walks it FOUR ROWS PER INSTRUCTION: lane =
// (row of the batch, 16-byte chunk of the 128-byte row), so one 128-bit access moves four rows of
// 32 channels.  No other warp, group or CTA touches this warp's table rows, and the adds into a
// cell happen in list order: the sums are deterministic.
//   shared table: 128-bit read-modify-write; rows of one node inside a batch go in turn
//   global table: red.global.add.v4.f32, fire and forget (L2 performs ~1 fp32 atomic per clock
//                 and slice: fine for the big shapes that have no room on chip, but at cfg2 the
//                 5.4 M atomics of a step take longer to drain than the step's tiles take)
//
// What this costs is instructions: a warp runs it serially once per tile.  History (cfg2, us per
// tile): counting sort + segment walk 3.5; one row per instruction with counts in registers 3.3
// (950 instructions); 128-bit lanes with an all-or-nothing serial fallback 4+ (63 % of the
// four-row batches hold a duplicate node); rows dealt into four hazard-free lists: costs more
// to build than it saves.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void publish_labels(uint8_t *smem, uint32_t lab_off, int *cnt_s, int K,
                                               int g, int quad, int lane, int L, uint32_t parity)
{
    uint16_t *lab = reinterpret_cast<uint16_t *>(smem + lab_off) + parity * 128u;
    lab[quad * 32 + lane] = (uint16_t)L;
    if (L < K) atomicAdd(cnt_s + L, 1);
    bar_sync(1u + (uint32_t)g, 128);
}

template <int NBLK, bool TABG, bool T8>
__device__ __noinline__ void tile_accumulate_blk(const TcParams &p, uint8_t *smem,
                                                    uint32_t xs_addr, uint32_t tab_addr,
                                                    float *tab_g, uint32_t lab_off,
                                                    uint32_t list_addr, int quad, int lane,
                                                    uint32_t parity)
{
    const TcPlan &pl = p.plan;
    const int K = pl.K, C = pl.C;
    const uint32_t pitch4 = (uint32_t)(TABG ? part_pitch(C) : tab_pitch(C)) * 4u;
    const uint16_t *lab = reinterpret_cast<const uint16_t *>(smem + lab_off) + parity * 128u;
    // 1. this warp's rows as a compact list of 8-byte entries {row * 128 | (row & 7) << 4,
    //    node * pitch bytes} (the list lives in the warp's idle pair buffer: 128 x 8 bytes)
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t total = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t l = lab[j * 32 + lane];
        const bool mine = (int)l < K && (int)(l & 3u) == quad;
        const uint32_t m = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            const uint32_t row = (uint32_t)(j * 32 + lane);
            const uint32_t pos = total + (uint32_t)__popc(m & lt);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list_addr + pos * 8u),
                         "r"(row * 128u | (row & 7u) << 4), "r"(l * pitch4)
                         : "memory");
        }
        total += (uint32_t)__popc(m);
    }
    __syncwarp();
    // 2. the walk.  Physical address of chunk ck of tile row r in a 32-channel block:
    //    (block + r * 128 + ck * 16) ^ ((r & 7) << 4)  (SWIZZLE_128B).
    const uint32_t r4 = (uint32_t)lane >> 3, ck = (uint32_t)lane & 7u;
    const uint32_t A = xs_addr + ck * 16u;
    const bool last_ok = (NBLK - 1) * 32 + (int)ck * 4 < C;  // the last block may be partial
    auto ldsf4 = [](uint32_t addr) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(addr));
        return v;
    };
    const uint32_t my = list_addr + r4 * 8u;
    // source of chunk ck of a row in block blk; a tail8 block keeps 32-byte rows (row = w0 >> 7,
    // chunk ck at ck ^ ((row >> 2) & 1); only ck < 2 gets here: last_ok)
    constexpr bool t8 = T8;
    const uint32_t XT = xs_addr + pl.x_tail_off;
    auto xsrc = [&](uint32_t xa, uint32_t w0, int blk) -> uint32_t {
        if (t8 && blk == NBLK - 1) return XT + ((w0 >> 7) << 5) + (((ck ^ (w0 >> 9)) & 1u) << 4);
        return xa + (uint32_t)blk * 16384u;
    };
    if constexpr (TABG) {
        // global table: fire and forget, nothing to order
        char *const G = reinterpret_cast<char *>(tab_g) + ck * 16u;
        auto add_row = [&](uint32_t w0, uint32_t w1) {
            const uint32_t xa = (A ^ (w0 & 0x70u)) + (w0 & 0xFFFFFF80u);
            char *cell = G + w1;
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                if (blk < NBLK - 1 || last_ok) {
                    const float4 x = ldsf4(xsrc(xa, w0, blk));
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                     cell + blk * 128),
                                 "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w)
                                 : "memory");
                }
            }
        };
        const uint32_t full = total & ~3u;
        for (uint32_t i = 0; i < full; i += 4u) {
            uint32_t w0, w1;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(my + i * 8u));
            add_row(w0, w1);
        }
        if (full + r4 < total) {  // the last, short batch
            uint32_t w0, w1;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(my + full * 8u));
            add_row(w0, w1);
        }
    } else {
        // shared table: read-modify-write.  Rows of one node inside a batch must go one after the
        // other: rank = how many EARLIER rows of the batch share my node (one match.any); round r
        // takes the rows of rank r.  Most batches need one or two rounds.
        const uint32_t T = tab_addr + ck * 16u;
        const uint32_t below = (1u << (r4 * 8u)) - 1u;
        auto rmw = [&](uint32_t xa, uint32_t ta, uint32_t w0) {
            float4 x[NBLK], t[NBLK];
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk)
                if (blk < NBLK - 1 || last_ok) x[blk] = ldsf4(xsrc(xa, w0, blk));
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk)
                if (blk < NBLK - 1 || last_ok) t[blk] = ldsf4(ta + (uint32_t)blk * 128u);
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk)
                if (blk < NBLK - 1 || last_ok)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta + (uint32_t)blk * 128u),
                                 "f"(t[blk].x + x[blk].x), "f"(t[blk].y + x[blk].y),
                                 "f"(t[blk].z + x[blk].z), "f"(t[blk].w + x[blk].w)
                                 : "memory");
        };
        for (uint32_t i = 0; i < total; i += 4u) {
            const bool on = i + r4 < total;  // the last batch may be short
            uint32_t w0 = 0u, w1 = 0x80000000u | r4;  // an idle lane group matches nobody
            if (on)
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(my + i * 8u));
            const unsigned same = __match_any_sync(0xffffffffu, w1);
            const int rank = __popc(same & below) >> 3;
            const uint32_t xa = (A ^ (w0 & 0x70u)) + (w0 & 0xFFFFFF80u), ta = T + w1;
            if (on && rank == 0) rmw(xa, ta, w0);
            if (__any_sync(0xffffffffu, rank >= 1)) {
                __syncwarp();
                if (on && rank == 1) rmw(xa, ta, w0);
                if (__any_sync(0xffffffffu, rank >= 2)) {
                    __syncwarp();
                    if (on && rank == 2) rmw(xa, ta, w0);
                    __syncwarp();
                    if (on && rank == 3) rmw(xa, ta, w0);
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();  // the list is the warp's pair buffer again from the next tile on
}

template <bool T8>
__device__ __forceinline__ void tile_accumulate(const TcParams &p, uint8_t *smem, uint32_t xs_addr,
                                                uint32_t tab_addr, float *tab_g, uint32_t lab_off,
                                                uint32_t list_addr, int quad, int lane,
                                                uint32_t parity)
{
    // warp-uniform dispatch: the block loop is unrolled and the table kind fixed per case
#define PIXIE_ACC_CASE(n_)                                                                        \
    case n_:                                                                                      \
        if (tab_g != nullptr)                                                                     \
            tile_accumulate_blk<n_, true, T8>(p, smem, xs_addr, tab_addr, tab_g, lab_off, list_addr,  \
                                          quad, lane, parity);                                    \
        else                                                                                      \
            tile_accumulate_blk<n_, false, T8>(p, smem, xs_addr, tab_addr, tab_g, lab_off, list_addr, \
                                           quad, lane, parity);                                   \
        break;
    switch (p.plan.nblkX) {
        PIXIE_ACC_CASE(1)
        PIXIE_ACC_CASE(2)
        PIXIE_ACC_CASE(3)
        default:
        PIXIE_ACC_CASE(4)
    }
#undef PIXIE_ACC_CASE
}

// ------------------------------------------------------------------------------------------------
// train mode, end of a step (called by the `nthr` non-producer threads of every CTA, tid < nthr):
//   1. [shared tables] the NG group tables are added in group order into the CTA's part of
//      p.partials and cleared; [global tables] they already are parts
//   2. grid barrier; every CTA folds ITS slice of the K x (C+1) table over all parts, part order,
//      fp64 (global tables are cleared by the thread that just read them)
//   3. [N > 1] cross-GPU sum over NVLink peer memory: see exchange_slice()
//   4. grid barrier; batch update (som_update_nodes), CTA j takes nodes j, j + grid, ...: W64, W32,
//      the codebook image rows and the norms of the next step; grid barrier
// Without p.apply (one step per launch: pixie_som_accum_f32) it stops after the fold.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fold_slice_sum(const TcParams &p, int nparts, int len, int e0,
                                               int e1, int per, int tid, int nthr, double *dst)
{
    const int ld = p.plan.C + 1, ldp = part_pitch(p.plan.C);
    const size_t plen = (size_t)p.plan.K * ldp;  // floats per part
    const int oct = tid >> 3, q = tid & 7, noct = nthr >> 3;
    const int rounds = (per + noct - 1) / noct;  // uniform trip count: shuffles below
    for (int it = 0; it < rounds; ++it) {
        const int e = e0 + it * noct + oct;  // logical element (node k, column c) of K x (C+1)
        double a = 0.0;
        if (e < e1) {
            const int k = e / ld, c = e - k * ld;
            const float *col = p.parts + (size_t)k * ldp + c;
            // thread q of the octet sums parts q, q + 8, ... (ascending), eight loads in flight
            for (int pp = q; pp < nparts; pp += 64) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[u] = pp + 8 * u < nparts ? __ldcg(col + (size_t)(pp + 8 * u) * plen) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) a += (double)v[u];
            }
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (e < e1 && q == 0) dst[e - e0] = a;
    }
}

// Cross-GPU sum of this CTA's slice [e0, e1) (s_slice holds this rank's folded values).
// Exchange buffer of a rank (torch symmetric memory, mapped on every rank):
//   cell uint32[4] [2][8][len]   parity of the step x SOURCE rank x element
// a cell = {low word, tag, high word, tag} of one fp64 value, written with ONE 16-byte store.
// Every CTA PUSHES each value of its slice into the cell [parity][my rank][e] of every other rank
// (one-way NVLink traffic, no round trip); the owner of the same slice there polls its OWN memory
// until both tags of the cell carry this step's number (8-byte halves are single-copy atomic, so a
// half with the right tag holds the right data: the scheme NCCL's LL protocol uses), and adds the
// `world` values in rank order -- every rank adds the same values in the same order: bit-identical
// totals.  No fence, no flag round trip, no grid barrier, no NCCL call, no host round trip: the
// cost of a step's exchange is one NVLink store latency plus the skew between the GPUs.
// Reuse of a parity slot two steps later is safe: a rank can only reach the push of step t + 2
// after it received this rank's values of step t + 1, which this CTA sent after it had read step t.
__device__ __forceinline__ void exchange_slice(const TcParams &p, int st, int len, int e0, int e1,
                                               int tid, int nthr, const double *s_slice)
{
    const int par = st & 1, ne = e1 - e0;
    const uint32_t tag = p.flag_base + (uint32_t)st + 1u;
    const size_t cell0 = ((size_t)par * 8 + (size_t)p.rank) * len;
    for (int i = tid; i < ne * p.world; i += nthr) {
        const int r = i / ne, e = e0 + (i - r * ne);
        if (r == p.rank) continue;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(s_slice[e - e0]);
        uint4 *dst = reinterpret_cast<uint4 *>(p.peer_buf[r]) + cell0 + e;
        asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst),
                     "r"((uint32_t)bits), "r"(tag), "r"((uint32_t)(bits >> 32)), "r"(tag)
                     : "memory");
    }
    const uint4 *mine = reinterpret_cast<const uint4 *>(p.peer_buf[p.rank]) + (size_t)par * 8 * len;
    for (int e = e0 + tid; e < e1; e += nthr) {
        double a = 0.0;
        for (int r = 0; r < p.world; ++r) {  // rank order
            double v = s_slice[e - e0];
            if (r != p.rank) {
                const uint4 *src = mine + (size_t)r * len + e;
                uint4 c;
                uint32_t spins = 0;
                uint64_t t0 = 0;
                do {
                    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
                                 : "l"(src)
                                 : "memory");
                    if (c.y == tag && c.w == tag) break;
                    if ((++spins & 1023u) == 0u) {
                        // a peer that never arrives (died, or took the other code path) makes this
                        // rank fail after 60 s instead of hanging
                        const uint64_t now = global_timer_ns();
                        if (t0 == 0) t0 = now;
                        if (now - t0 > 60000000000ull) __trap();
                    }
                } while (true);
                v = __longlong_as_double((long long)(((unsigned long long)c.z << 32) | c.x));
            }
            a += v;
        }
        p.SN[e] = a;
    }
}

template <int NG>
__device__ __noinline__ void step_finish(const TcParams &p, int st, uint8_t *smem, int tid,
                                         unsigned int &gb_target, uint64_t &tm)
{
    const TcPlan &pl = p.plan;
    const int K = pl.K, C = pl.C, len = K * (C + 1);
    constexpr uint32_t nthr = (uint32_t)NG * 128u + 32u;
    const bool leader = tid == 0;
    unsigned int *gsync = &p.ctl->grid_sync;
    const bool timing = blockIdx.x == 0 && tid == 0;
    auto lap = [&](int slot) {
        if (timing) {
            const uint64_t now = global_timer_ns();
            p.ctl->phase_ns[slot] += now - tm;
            tm = now;
        }
    };
    lap(0);  // tiles
#ifdef PIXIE_PROFILE
    // arrival time of every CTA at the end of its tiles, steps dbg_step0 .. +1: slots behind the
    // 32 x 1024 event slots of the trace buffer
    if (p.trace && tid == 0 && st >= p.dbg_step0 && st < p.dbg_step0 + 2)
        p.trace[2u * 32u * 1024u + (unsigned)(st - p.dbg_step0) * 256u + blockIdx.x] = global_timer_ns();
#endif

    // 1. this CTA's part of the statistics: its NG group tables are added in group order, the
    //    counts joined in, tables and counts cleared for the next step.
    const bool tabg = pl.tab_global != 0;
    if (tabg) __threadfence();  // this thread's red operations (fire-and-forget) are performed
    bar_sync(kBarStep, nthr);   // every group is done with its last tile
    const int ldp = part_pitch(C), plen = K * ldp;
    int *cnt_s = reinterpret_cast<int *>(smem + pl.off_cnt);
    {
        float4 *grp = reinterpret_cast<float4 *>(p.partials + (size_t)blockIdx.x * NG * plen);
        float4 *mine = reinterpret_cast<float4 *>(p.parts + (size_t)blockIdx.x * plen);
        const float4 *acc4 = reinterpret_cast<const float4 *>(smem + pl.off_acc);
        const int ldp4 = ldp >> 2, plen4 = plen >> 2, Cp4 = tab_pitch(C) >> 2;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < plen4; i += (int)nthr) {
            const int k = i / ldp4, q4 = i - k * ldp4, c4 = q4 * 4;
            float4 v = zero4;
            if (tabg) {
#pragma unroll
                for (int gg = 0; gg < NG; ++gg) {
                    const float4 u = __ldcg(grp + (size_t)gg * plen4 + i);
                    __stcg(grp + (size_t)gg * plen4 + i, zero4);
                    v.x += u.x, v.y += u.y, v.z += u.z, v.w += u.w;
                }
            } else if (q4 < Cp4) {
#pragma unroll
                for (int gg = 0; gg < NG; ++gg) {
                    float4 *cell = const_cast<float4 *>(acc4) + ((size_t)gg * K + k) * Cp4 + q4;
                    const float4 u = *cell;
                    *cell = zero4;
                    v.x += u.x, v.y += u.y, v.z += u.z, v.w += u.w;
                }
            }
            if (c4 <= C && C < c4 + 4) {  // the chunk that holds the count column
                int n = 0;
#pragma unroll
                for (int gg = 0; gg < NG; ++gg) {
                    n += cnt_s[gg * K + k];
                    cnt_s[gg * K + k] = 0;
                }
                (&v.x)[C - c4] = (float)n;
            }
            __stcg(mine + i, v);
        }
    }
    gb_target += gridDim.x;
    grid_barrier(gsync, gb_target, nthr, leader);
    lap(1);

    // 2. fold this CTA's slice
    const int nparts = (int)gridDim.x;
    const int per = (len + (int)gridDim.x - 1) / (int)gridDim.x;
    const int e0 = min(len, (int)blockIdx.x * per), e1 = min(len, e0 + per);
    double *s_slice = reinterpret_cast<double *>(smem + pl.off_pairs);  // pair lists are idle
    const bool direct = p.world <= 1;
    fold_slice_sum(p, nparts, len, e0, e1, per, tid, (int)nthr, direct ? p.SN + e0 : s_slice);
    if (!direct) {
        bar_sync(kBarStep, nthr);
        exchange_slice(p, st, len, e0, e1, tid, (int)nthr, s_slice);
    }
    if (!p.apply) return;
    if (blockIdx.x == 0 && tid == 0) {
        // slot the update below publishes the next step's norms into
        p.ctl->pp_wmax_bits[(st + 1) & 1] = 0;
        p.ctl->pp_w_has_negative[(st + 1) & 1] = 0;
    }
    lap(2);
    gb_target += gridDim.x;
    grid_barrier(gsync, gb_target, nthr, leader);
    lap(3);

    // 3. batch update; schedule of step t = t0 + st of T
    {
        const double frac = (double)(p.t0 + st) / (double)p.T;
        const double r = p.r0 - (p.r0 - p.r1) * frac;
        const double r_eff = r < 1.0 ? 0.5 : r;
        const double sigma = 0.5 * r_eff;
        const double inv2s2 = 1.0 / (2.0 * sigma * sigma);
        const double alpha = p.a0 - (p.a0 - p.a1) * frac;
        char *img = reinterpret_cast<char *>(p.wimg_rw);
        int *wmax_slot = &p.ctl->pp_wmax_bits[(st + 1) & 1];
        int *neg_slot = &p.ctl->pp_w_has_negative[(st + 1) & 1];
        som_update_nodes(
            p.SN, p.W64, p.W32, K, C, p.ydim, inv2s2, alpha, (int)blockIdx.x, (int)gridDim.x, tid,
            (int)nthr, reinterpret_cast<double *>(smem + pl.off_pairs),
            [&] { bar_sync(kBarStep, nthr); },
            [&](int k, int c, float wf) {
                *reinterpret_cast<float *>(
                    img + img_offset(lay_pack(pl.Ntot, pl.tail8 ? pl.nblkW - 1 : -1), k, c)) = -2.0f * wf;
            },
            [&](int k, int lane, double nrm2, bool neg) {
                if (lane != 0) return;
                float bias = (float)nrm2;
                if (!(bias <= FLT_MAX)) bias = FLT_MAX;
                const float h = __uint_as_float(__float_as_uint(bias) & 0xFFFFE000u);
                const float r1 = bias - h;
                const float m = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
                const float l = r1 - m;
                float *bb = reinterpret_cast<float *>(img + pl.off_bias + bias_offset(k, 0));
                bb[0] = h;
                bb[1] = m;
                bb[2] = l;
                float nr = (float)sqrt(nrm2) * 1.0000005f;
                if (!(nr <= FLT_MAX)) nr = FLT_MAX;
                atomicMax(wmax_slot, __float_as_int(nr));
                if (neg) atomicOr(neg_slot, 1);
            });
    }
    lap(4);
    gb_target += gridDim.x;
    grid_barrier(gsync, gb_target, nthr, leader);
    lap(5);
#ifdef PIXIE_PROFILE
    if ((p.dbg_flags & 4) && timing) {
        const unsigned long long *c = p.ctl->tile_cyc;
        printf("step %2d tiles %llu | cyc: waitX %llu norm %llu passes %llu resolve %llu fix %llu accwork %llu accwait %llu | ns: tiles %llu b1 %llu fold %llu b2 %llu upd %llu b3 %llu\n",
               st, c[6], c[0], c[1], c[2], c[3], c[4], c[5], c[7], p.ctl->phase_ns[0],
               p.ctl->phase_ns[1], p.ctl->phase_ns[2], p.ctl->phase_ns[3], p.ctl->phase_ns[4],
               p.ctl->phase_ns[5]);
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// T8: the plan's "tail8" layout (TcPlan::tail8) as a compile-time switch -- as a run-time flag its
// branches and the extra live values cost the 96-register epilogue of every OTHER shape 7 % (cfg2
// assign 1.85 -> 1.98 ms, same box), so shapes without a tail block run instantiations that do
// not contain it.
template <int SL, int SPC, int NCH, int NG, bool ACC, bool T8 = false>
__global__ void __launch_bounds__(NG * 128 + 64, 1)
bmu_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ TcParams p)
{
    constexpr int NEPI = NG * 4;            // epilogue warps
    constexpr int NCHUNK = SL * SPC;        // codebook rows per accumulator chunk
    constexpr int NMMA = (NCHUNK + 15) / 16 * 16;  // UMMA N (columns past NCHUNK are never read)
    static_assert(NCH == 1 || NCHUNK % 8 == 0, "chunk base must stay on a swizzle-atom row");
    constexpr int NS = SPC * NCH;           // slices per tile
    constexpr int NW = (SL + 31) / 32;      // mask words per slice
    // TMEM accumulator buffers: one per epilogue group (NCH = 1: the tile's only chunk; NCH = 4:
    // the tile's chunks go through it one after the other), or two shared by the two groups
    // (NCH = 2: both chunks of a tile in flight at once)
    constexpr int NBUF = NCH == 2 ? 2 : NG;
    static_assert(NCH == 1 || NCH == 2 || NCH == 4, "accumulator chunks per tile");
    static_assert(NCH != 2 || NG == 2, "two chunks per tile only with two epilogue groups");

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
        // shared-memory tables and the counts start at zero (global tables are cleared by the host
        // before the launch and by their owner after every step)
        if (!pl.tab_global) {
            float *a = reinterpret_cast<float *>(smem + pl.off_acc);
            for (int i = threadIdx.x; i < NG * pl.K * tab_pitch(pl.C); i += blockDim.x) a[i] = 0.f;
        }
        int *cz = reinterpret_cast<int *>(smem + pl.off_cnt);
        for (int i = threadIdx.x; i < NG * pl.K; i += blockDim.x) cz[i] = 0;
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
    unsigned int gb_target = 0;  // grid-barrier instances passed so far x gridDim.x
    uint32_t st_flag = 0, st_pairs = 0, st_fp64 = 0, st_fix = 0;  // per-thread statistics
    uint64_t lap_t = (ACC && blockIdx.x == 0 && threadIdx.x == 0) ? global_timer_ns() : 0;
    PIXIE_PROF_DECL();
    // thread index among the non-producer threads (epilogue warps, then the MMA warp)
    const int tid_np = warp < NEPI ? (int)threadIdx.x : (int)threadIdx.x - 32;

    // codebook image -> shared memory: linear bulk copies (the image is pre-swizzled in global
    // memory); in whole-pass mode it was rewritten by other CTAs through the generic proxy
    auto load_image = [&]() {
        if (elect_one()) {
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_arrive_expect_tx(bar_w, pl.wimg_bytes);
            for (uint32_t off = 0; off < pl.wimg_bytes; off += 16384u) {
                const uint32_t sz = min(16384u, pl.wimg_bytes - off);
                bulk_load(sbase + off, reinterpret_cast<const uint8_t *>(p.wimg) + off, sz, bar_w);
            }
        }
        __syncwarp();
    };
    // X tiles of one step -> stages.  The whole warp walks the loop (warp-uniform control flow);
    // one elected lane issues.
    // Tile it + nstage -- the next occupant of the stage being filled -- is prefetched into L2 at
    // the same time (past the step's last tile: the first tiles of the next step, `nxt`).  Worth
    // 2 % where the pipeline is shallow (K = 400, C = 40: three stages), nothing elsewhere
    // (profiles/r02_notes.md section 8); PIXIE_DBG_FLAGS=8 switches it off for A/B runs.
    auto produce_step = [&](const StepTiles &stp, uint32_t cnt, uint32_t seq0, int st,
                            const StepTiles &nxt, uint32_t nxt_cnt) {
        uint32_t s = seq0 % (uint32_t)nstage;          // one division per step, then
        uint32_t ph = (seq0 / (uint32_t)nstage) & 1u;  // incremental
        for (uint32_t it = 0; it < cnt; ++it, ph ^= (++s == (uint32_t)nstage), s = s == (uint32_t)nstage ? 0u : s) {
            mbar_wait(bar_empty + 8u * s, ph ^ 1u);
            const int64_t j = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t tile = stp.first + j * stp.stride;
            const int32_t row0 = (int32_t)(tile * kTile);
            const uint32_t ita = it + (uint32_t)nstage;
            int64_t ptile = -1;
            if (!(p.dbg_flags & 8)) {
                if (ita < cnt)
                    ptile = stp.first + ((int64_t)blockIdx.x + (int64_t)ita * gridDim.x) * stp.stride;
                else if (ita - cnt < nxt_cnt)
                    ptile = nxt.first +
                            ((int64_t)blockIdx.x + (int64_t)(ita - cnt) * gridDim.x) * nxt.stride;
            }
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full + 8u * s, pl.stage_bytes);
                const int nfullb = pl.nblkX - (T8 ? 1 : 0);  // a tail8 block has its own tensor map
                for (int b = 0; b < nfullb; ++b)
                    tma_load_2d(sbase + pl.off_x + s * pl.stage_bytes + (uint32_t)b * 16384u,
                                &tmX, bar_full + 8u * s, b * 32, row0, kEvictFirst);
                if constexpr (T8)
                    tma_load_2d(sbase + pl.off_x + s * pl.stage_bytes + pl.x_tail_off, &p.tm_tail,
                                bar_full + 8u * s, nfullb * 32, row0, kEvictFirst);
                if (ptile >= 0)
                    for (int b = 0; b < pl.nblkX; ++b)
                        tma_prefetch_2d(&tmX, b * 32, (int32_t)(ptile * kTile));
            }
            __syncwarp();
            PIXIE_TRACE(0, seq0 + it);
        }
    };
    auto tiles_of = [&](const StepTiles &stp) {
        return (int64_t)blockIdx.x < stp.count
                   ? (uint32_t)((stp.count - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    };

    if (ACC && warp == NEPI) {
        // ============================================================ TMA producer, train mode
        // Runs through ALL steps on its own: X does not depend on the codebook, so the next step's
        // first tiles are already in flight while the other warps finish the current step.
        StepTiles stp = step_tiles(p, 0);
        uint32_t cnt = tiles_of(stp);
        for (int st = 0; st < nsteps; ++st) {
            const bool more = st + 1 < nsteps;
            const StepTiles nxt = more ? step_tiles(p, st + 1) : StepTiles{0, 0, 0};
            const uint32_t nxt_cnt = more ? tiles_of(nxt) : 0u;
            produce_step(stp, cnt, base_seq, st, nxt, nxt_cnt);
            base_seq += cnt;
            stp = nxt;
            cnt = nxt_cnt;
        }
    } else {
    for (int st = 0; st < nsteps; ++st) {
    const StepTiles stp = ACC ? step_tiles(p, st) : StepTiles{p.tile_first, p.tile_stride, p.ntiles};
    const uint32_t cnt = tiles_of(stp);

    if (warp == NEPI) {
        // ============================================================ TMA producer (assignment)
        load_image();
        produce_step(stp, cnt, base_seq, st, StepTiles{0, 0, 0}, 0u);
    } else if (warp == NEPI + 1) {
        // ============================================================ MMA issuer
        // The whole warp walks the loop and waits on the barriers; one elected lane (always the
        // same one, as tcgen05.commit requires) issues the MMAs of a tile and their commit.
        {
            constexpr uint32_t idesc = umma_idesc_tf32(128, (uint32_t)NMMA);
            const uint64_t desc_ones = umma_desc_nosw(sbase + pl.off_ones, 128u, 256u);
            // bias K-step: the no-swizzle block behind the image (8-row groups 256 bytes apart)
            const uint32_t wblk_bytes = (uint32_t)((NCH - 1) * NCHUNK + NMMA) * 128u;
            if constexpr (ACC) load_image();  // this step's codebook (the producer warp is busy ahead)
            mbar_wait(bar_w, (uint32_t)st & 1u);
            // the MMAs of one accumulator chunk (all K-steps + the bias K-step) and their commit
            auto issue_chunk = [&](uint32_t xs_addr, int c, uint32_t buf) {
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + buf * (uint32_t)NMMA;
                    const uint32_t wrow = (uint32_t)(c * NCHUNK) * 128u;
                    const int kfull = pl.ksteps - (T8 ? 1 : 0);
                    for (int ks = 0; ks < kfull; ++ks) {
                        const uint32_t blk = (uint32_t)(ks >> 2), ko = (uint32_t)(ks & 3) * 32u;
                        const uint64_t da = umma_desc_sw128(xs_addr + blk * 16384u + ko);
                        const uint64_t db = umma_desc_sw128(sbase + blk * wblk_bytes + wrow + ko);
                        mma_tf32(d_tmem, da, db, idesc, ks > 0 ? 1u : 0u);
                    }
                    if constexpr (T8) {  // the last K-step: 32-byte rows
                        const uint64_t da = umma_desc_sw32(xs_addr + pl.x_tail_off);
                        const uint64_t db =
                            umma_desc_sw32(sbase + pl.w_tail_off + (uint32_t)(c * NCHUNK) * 32u);
                        mma_tf32(d_tmem, da, db, idesc, 1u);
                    }
                    const uint64_t dbias = umma_desc_nosw(
                        sbase + pl.off_bias + (uint32_t)(c * NCHUNK / 8) * 256u, 128u, 256u);
                    mma_tf32(d_tmem, desc_ones, dbias, idesc, 1u);
                    mma_commit(bar_tfull + 8u * buf);
                }
                __syncwarp();
            };
            if constexpr (NCH <= 2) {
                uint32_t s = base_seq % (uint32_t)nstage;
                uint32_t ph = (base_seq / (uint32_t)nstage) & 1u;
                for (uint32_t it = 0; it < cnt; ++it, ph ^= (++s == (uint32_t)nstage), s = s == (uint32_t)nstage ? 0u : s) {
                    const uint32_t seq = base_seq + it;
                    mbar_wait(bar_full + 8u * s, ph);
                    PIXIE_TRACE(1, seq);
                    const uint32_t xs_addr = sbase + pl.off_x + s * pl.stage_bytes;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const uint32_t q = seq * (uint32_t)NCH + (uint32_t)c;  // accumulator-chunk counter
                        const uint32_t buf = q % NBUF;
                        const uint32_t bph = (q / NBUF) & 1u;
                        mbar_wait(bar_tempty + 8u * buf, bph ^ 1u);
                        tc_fence_after();
                        issue_chunk(xs_addr, c, buf);
                        PIXIE_TRACE(2, seq);
                    }
                }
            } else {
                // Four chunks per tile through the group's ONE buffer: chunk c + 1 can only be
                // issued once the group has drained chunk c.  Issuing tile by tile would stall the
                // tensor pipe on every drain, so the tiles go in rounds of NG (one per group) and
                // the round is walked chunk-major: while group g drains chunk c, chunk c of the
                // other groups' tiles is issued.
                for (uint32_t it0 = 0; it0 < cnt; it0 += (uint32_t)NG) {
                    const uint32_t nt = cnt - it0 < (uint32_t)NG ? cnt - it0 : (uint32_t)NG;
#pragma unroll 1
                    for (int c = 0; c < NCH; ++c) {
                        for (uint32_t t = 0; t < nt; ++t) {
                            const uint32_t seq = base_seq + it0 + t;
                            const uint32_t s = seq % (uint32_t)nstage;
                            const uint32_t g = seq % (uint32_t)NG, use = seq / (uint32_t)NG;
                            if (c == 0) mbar_wait(bar_full + 8u * s, (seq / (uint32_t)nstage) & 1u);
                            const uint32_t q = use * (uint32_t)NCH + (uint32_t)c;  // per-group chunk counter
                            mbar_wait(bar_tempty + 8u * g, (q & 1u) ^ 1u);
                            tc_fence_after();
                            issue_chunk(sbase + pl.off_x + s * pl.stage_bytes, c, g);
                        }
                    }
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
        const int lay = T8 ? lay_pack(Ntot, pl.nblkX - 1) : Ntot;  // image / tile layout key
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

            PIXIE_TICK(0);
            mbar_wait(bar_full + 8u * s, ph);  // X tile landed
            PIXIE_TOCK(0);
            PIXIE_TRACE(3, seq);

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
                for (int b = 0; b < pl.nblkX - (T8 ? 1 : 0); ++b, xaddr += 16384u) {
#pragma unroll
                    for (uint32_t pc = 0; pc < 8; ++pc) {
                        const uint4 x = lds128(xaddr ^ (pc << 4));
                        const uint64_t lo = pack2u(x.x, x.y), hi = pack2u(x.z, x.w);
                        xa = fma2(lo, lo, xa);
                        xb = fma2(hi, hi, xb);
                        sgn |= (x.x | x.y) | (x.z | x.w);
                    }
                }
                if constexpr (T8) {  // 32-byte rows: both chunks, either order
                    const uint32_t ta = sbase + pl.off_x + (uint32_t)s * pl.stage_bytes +
                                        pl.x_tail_off + (uint32_t)row * 32u;
#pragma unroll
                    for (uint32_t pc = 0; pc < 2; ++pc) {
                        const uint4 x = lds128(ta + (pc << 4));
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

            PIXIE_TOCK(1);
            float m_run = __int_as_float(0x7f800000);
            uint32_t mw[NS][NW];
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int w = 0; w < NW; ++w) mw[a][w] = 0u;

            if constexpr (NCH == 2 && !ACC && T8) {
                // K > 256 with a tail block (cfg3: 40 channels, 20 x 20): the four slices of a
                // tile as ONE loop body instead of four unrolled copies.  The unrolled stream of
                // this instantiation is ~2000 instructions (32 KB) per tile, more than the 32 KB
                // instruction cache behind the L0s holds: its hit rate was 79 % and
                // instruction-fetch stalls 17 % of the epilogue warps' time (profiles/r02_notes.md
                // section 12).  The masks of the slice in hand live in cur[] and are committed to
                // mw[sl] by a chain of selects (register arrays need static indices).  Measured:
                // C = 40 5.02 -> 4.86 ms; without a tail block the rolled loop LOSES (C = 16
                // 1.20 -> 1.28 ms, C = 64 1.57 -> 1.59 ms), so those keep the unrolled form.
#pragma unroll 1
                for (int sl = 0; sl < NS; ++sl) {
                    const int c = sl / SPC, sidx = sl - c * SPC;
                    const uint32_t q = seq * 2u + (uint32_t)c;  // chunk counter; buffers alternate
                    const uint32_t buf = q & 1u, bph = (q >> 1) & 1u;
                    if (sidx == 0) {
                        // (see below: the previous use's release pins the barrier's phase)
                        if (q >= 2) mbar_wait(bar_tempty + 8u * buf, bph ^ 1u);
                        mbar_wait(bar_tfull + 8u * buf, bph);
                        tc_fence_after();
                    }
                    uint32_t vr[SL];
                    tmem_ld_cols<SL>(tmem_lane + buf * (uint32_t)NMMA + (uint32_t)(sidx * SL), vr);
                    tc_wait_ld();
                    if (sidx == SPC - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8u * buf);
                    }
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
                    // earlier candidates are out of range once the minimum drops by > delta
                    // (the masks of this and the later slices are still zero)
                    const bool drop = m_new + delta < m_run;
#pragma unroll
                    for (int a = 0; a < NS; ++a)
#pragma unroll
                        for (int w = 0; w < NW; ++w) mw[a][w] = drop ? 0u : mw[a][w];
                    m_run = m_new;
                    const float thr = m_run + delta;
                    uint32_t cur[NW];
#pragma unroll
                    for (int w = 0; w < NW; ++w) cur[w] = 0u;
                    if (sl == 0 || __any_sync(0xffffffffu, ms < thr)) {
                        const uint64_t thr2 = pack2(thr, thr);
#pragma unroll
                        for (int i = 0; i < SL; i += 2) {
                            uint32_t d0, d1;
                            unpack2u(sub2(pack2u(vr[i], vr[i + 1]), thr2), d0, d1);  // FADD2
                            cur[i >> 5] = __funnelshift_l(d0, cur[i >> 5], 1);
                            cur[(i + 1) >> 5] = __funnelshift_l(d1, cur[(i + 1) >> 5], 1);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < NS; ++a)
#pragma unroll
                        for (int w = 0; w < NW; ++w) mw[a][w] = a == sl ? cur[w] : mw[a][w];
                }
            } else {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint32_t buf, bph;
                if constexpr (NCH != 2) {
                    // the group's own buffer; its uses are numbered use * NCH + c
                    buf = (uint32_t)g;
                    bph = (use * (uint32_t)NCH + (uint32_t)c) & 1u;
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
                PIXIE_TRACE(4, seq);
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

            PIXIE_TOCK(2);
            PIXIE_TRACE(5, seq);
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
                    duel_dist2_f32<T8>(xs, ws, lay, nchunks16, row, c0, c1, d0, d1);
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
                            const double e0 = pair_dist_f64<T8>(xs, ws, lay, pl.C, row, c0);
                            const double e1 = pair_dist_f64<T8>(xs, ws, lay, pl.C, row, c1);
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
                    d2buf[pi] = pair_dist2_f32<T8>(xs, ws, lay, nchunks16, prow, pnode, (uint32_t)lane & 7u);
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
                            const double d = pair_dist_f64<T8>(xs, ws, lay, pl.C, row, k);
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
            PIXIE_TOCK(3);
            if constexpr (ACC) {
                // Train mode resolves the rare rows the three stages could not settle right here
                // (their sums must be in this step's table): the warp runs the reference loop for
                // such a row cooperatively, lane l over nodes l, l+32, ...; the lexicographic
                // (distance, index) minimum over lanes is the first minimum of the sequential loop.
                // A row that only had too many candidates (a collapsed map early in training: more
                // than kMaxCand nodes inside the window, or a full pair list) does not need all K
                // nodes: its candidate masks are a proven superset of the minimum, so the lanes
                // share out THOSE nodes (up to 64; node order = mask order) -- K / nc times less
                // fp64 work per row.  Rows with non-finite scores take the full loop.
                unsigned fixm = __ballot_sync(0xffffffffu, label == kLabelFixup && grow < p.n);
                while (fixm) {
                    const int src = __ffs(fixm) - 1;
                    fixm &= fixm - 1;
                    const int frow = quad * 32 + src;
                    int minid = 0x7fffffff;
                    double mind = DBL_MAX;
                    const int nc_src = __shfl_sync(0xffffffffu, finite ? nc : 0, src);
                    if (nc_src >= 2 && nc_src <= 64) {
                        int myk0 = -1, myk1 = -1, rank = 0;
#pragma unroll
                        for (int a = 0; a < NS; ++a)
#pragma unroll
                            for (int w = 0; w < NW; ++w) {
                                const int cw = (SL - 32 * w) < 32 ? (SL - 32 * w) : 32;
                                const int base = a * SL + 32 * w + cw - 32;
                                uint32_t m = __shfl_sync(0xffffffffu, mw[a][w], src);
                                while (m) {  // warp-uniform
                                    const int lz = __clz(m);
                                    m &= ~(0x80000000u >> lz);
                                    if ((rank & 31) == lane) {
                                        if (rank < 32) myk0 = base + lz; else myk1 = base + lz;
                                    }
                                    ++rank;
                                }
                            }
                        // fp32 first (stage 2 of the row, a candidate per lane): only the nodes
                        // within the fp32 error of the smallest distance can be the minimum.  In
                        // the first pass from a sampled codebook (hundreds of nodes drawn from a
                        // few dozen populations) nearly every row comes through here with 20+
                        // candidates whose distances differ in the second digit: one survivor,
                        // no fp64 at all (it used to be an fp64 distance per candidate: 10 us per
                        // K = 400 tile, a third of the training pass).
                        const bool v0 = myk0 >= 0 && myk0 < pl.K, v1 = myk1 >= 0 && myk1 < pl.K;
                        float f0 = __int_as_float(0x7f800000), f1 = f0;
                        if (v0) f0 = pair_dist2_f32<T8>(xs, ws, lay, nchunks16, frow, myk0, (uint32_t)lane & 7u);
                        if (__any_sync(0xffffffffu, v1) && v1)
                            f1 = pair_dist2_f32<T8>(xs, ws, lay, nchunks16, frow, myk1, (uint32_t)lane & 7u);
                        float best = fminf(f0, f1);
                        for (int o = 16; o > 0; o >>= 1)
                            best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
                        const float bound = best * (1.0f + eps32) + 1.0e-30f;
                        const bool s0 = v0 && f0 <= bound, s1 = v1 && f1 <= bound;
                        const unsigned b0 = __ballot_sync(0xffffffffu, s0);
                        const unsigned b1 = __ballot_sync(0xffffffffu, s1);
                        const int nsurv = __popc(b0) + __popc(b1);
                        if (nsurv == 1) {
                            const int wl = b0 ? __ffs(b0) - 1 : __ffs(b1) - 1;
                            minid = __shfl_sync(0xffffffffu, b0 ? myk0 : myk1, wl);
                            mind = 0.0;  // every lane agrees: the reduction below returns minid
                        } else if (nsurv >= 2) {
                            // stage 3 over the survivors: the reference's own fp64 sequence
                            if (s0) {
                                mind = pair_dist_f64<T8>(xs, ws, lay, pl.C, frow, myk0);
                                minid = myk0;
                            }
                            if (s1) {
                                const double d = pair_dist_f64<T8>(xs, ws, lay, pl.C, frow, myk1);
                                if (d < mind) {  // myk1 > myk0: strict < keeps the lower index on ties
                                    mind = d;
                                    minid = myk1;
                                }
                            }
                            if (!(mind == mind)) mind = DBL_MAX, minid = 0x7fffffff;
                        }  // no survivor (every distance NaN): the full loop below
                    }
                    if (!__any_sync(0xffffffffu, minid != 0x7fffffff)) {
                        for (int k = lane; k < pl.K; k += 32) {
                            const double d = pair_dist_f64<T8>(xs, ws, lay, pl.C, frow, k);
                            if (d < mind) {
                                mind = d;
                                minid = k;
                            }
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
            PIXIE_TOCK(4);
            PIXIE_TRACE(6, seq);
            if constexpr (do_acc) {
                // ---- fused per-node sums (deterministic, no atomics between warps)
                float *tab_s = reinterpret_cast<float *>(smem + pl.off_acc) +
                               (size_t)g * pl.K * tab_pitch(pl.C);
                float *tab_g = pl.tab_global
                                   ? p.partials + ((size_t)blockIdx.x * NG + (size_t)g) * pl.K *
                                                      part_pitch(pl.C)
                                   : nullptr;
                publish_labels(smem, pl.off_lab + (uint32_t)g * 512u,
                               reinterpret_cast<int *>(smem + pl.off_cnt) + g * pl.K, pl.K, g, quad,
                               lane, (grow < p.n && label > 0) ? label - 1 : pl.K, use & 1u);
                PIXIE_TOCK(7);
                PIXIE_TRACE(7, seq);
                tile_accumulate<T8>(p, smem, sbase + pl.off_x + (uint32_t)s * pl.stage_bytes,
                                smem_u32(tab_s), tab_g, pl.off_lab + (uint32_t)g * 512u,
                                smem_u32(pairs), quad, lane, use & 1u);
            }
            PIXIE_TOCK(5);
            PIXIE_TRACE(8, seq);
            PIXIE_TILE_DONE();
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
        PIXIE_PROF_FLUSH();
        PIXIE_TRACE(9, 0);
        step_finish<NG>(p, st, smem, tid_np, gb_target, lap_t);
        PIXIE_TRACE(10, 0);
    }
    }  // for st
    }  // non-producer roles

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
}

template <int SL, int SPC, int NCH, int NG, bool ACC, bool T8>
static cudaError_t launch_variant(const CUtensorMap &tmX, const TcParams &p, int grid,
                                  cudaStream_t stream)
{
    constexpr int kThreads = NG * 128 + 64;
    cudaError_t e = cudaFuncSetAttribute(bmu_tc_kernel<SL, SPC, NCH, NG, ACC, T8>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.plan.smem_bytes);
    if (e != cudaSuccess) return e;
    if constexpr (ACC) {
        // the fused-sums variants use grid-wide barriers: launch cooperatively so the runtime
        // guarantees (or refuses) co-residency of all CTAs
        void *args[] = {const_cast<CUtensorMap *>(&tmX), const_cast<TcParams *>(&p)};
        e = cudaLaunchCooperativeKernel(
            reinterpret_cast<const void *>(&bmu_tc_kernel<SL, SPC, NCH, NG, ACC, T8>),
            dim3(grid), dim3(kThreads), args, p.plan.smem_bytes, stream);
        count_launch();
        return e;
    } else {
        bmu_tc_kernel<SL, SPC, NCH, NG, ACC, T8>
            <<<grid, kThreads, p.plan.smem_bytes, stream>>>(tmX, p);
        count_launch();
        return cudaGetLastError();
    }
}

// Body of a variant family's launcher: dispatch on the plan's template parameters.  Only the plain
// (assignment) family has tail8 instantiations: make_tc_plan() never sets tail8 for train-mode plans.
#ifdef PIXIE_FAMILY_NO_T8
#define PIXIE_LAUNCH_T8(a_, b_, c_, d_) \
    launch_variant<a_, b_, c_, d_, PIXIE_FAMILY_ACC, false>(tmX, p, grid, stream)
#else
#define PIXIE_LAUNCH_T8(a_, b_, c_, d_)                                                            \
    (pl.tail8 ? launch_variant<a_, b_, c_, d_, PIXIE_FAMILY_ACC, true>(tmX, p, grid, stream)      \
              : launch_variant<a_, b_, c_, d_, PIXIE_FAMILY_ACC, false>(tmX, p, grid, stream))
#endif
#define PIXIE_VARIANT(a_, b_, c_, d_)                               \
    if (pl.SL == a_ && pl.spc == b_ && pl.NCH == c_ && pl.NG == d_) \
        return PIXIE_LAUNCH_T8(a_, b_, c_, d_);
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
