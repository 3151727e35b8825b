// bmu_x3_kernel.cuh -- BMU assignment with SPLIT tf32 operands for the default Pixie shape
// (K <= 104 nodes, C <= 32 channels: the 10 x 10 SOM of cluster_helpers.py:106-109 / :152-157 on a
// MIBI panel).  Same contract as bmu_tc_kernel (labels bit-identical to the reference's fp64
// loop), same roles and pipeline, one difference in stage 1:
//
//   bmu_tc_kernel feeds X and -2 W to kind::tf32 as they are; the tensor core reads 11 significant
//   bits of each, so a score is only known to ~2^-8 ||x|| max||w|| and 14 % of the rows (99 % of
//   the warps) have a second node inside that window and go through the fp32 / fp64 recheck --
//   0.24 of the HBM roofline (profiles/r01_notes.md, r02_notes.md section 6).
//
//   Here every operand is split into the part the tensor core reads and the part it drops,
//   x = xh + xl, w' = wh + wl (w' = -2 w; xl = x - trunc(x) and wl likewise are exact in fp32), and
//   three MMAs per K-step accumulate  xh.wh + xh.wl + xl.wh  into the same TMEM columns.  What is
//   lost is xl.wl and the truncation of xl and wl themselves, 3 * 2^-20 |x||w| per product, plus
//   the accumulator's own rounding.  The window shrinks by two orders of magnitude, ~0.1 % of
//   the rows keep a second candidate, and those few are settled by the thread that owns the row
//   (two candidates: both fp32 distances in one walk, the fp64 replica of the reference loop only
//   on an fp32 tie; more: the fp64 replica over the candidates) -- no pair lists, no compaction.
//
//   The low part of an X tile is produced by the epilogue group that will consume the tile, in
//   the same walk over its rows that yields ||x||^2 for the bound, one tile ahead: after draining
//   the scores of tile t the group converts tile t + 4 (its next) into its private low-part
//   buffer and signals the MMA warp, then resolves tile t while the tensor core works.  The low
//   part of -2 W is a second set of image blocks written by codebook_prep_kernel.
//
// Cost: 3 MMAs per K-step instead of 1, and they are slow for this shape (M 128 x N 112 x K 8:
// ~100 cycles each, measured), plus 64 KB of shared memory for the low-part buffers.  Measured on
// the bench data, same box, plain kernel -> this one (profiles/r02_notes.md section 9):
//   C = 16, K = 100:  1.79 -> 1.42 ms        C = 24, K = 64:  1.61 -> 1.46 ms
//   C = 32, K = 100:  1.85 -> 2.05 ms (13 MMAs per tile: the tensor pipe becomes the bound)
// so make_x3_plan() takes it for C <= 24 (three K-steps) and K <= 104; PIXIE_X3=2 forces it up
// to C = 32, PIXIE_X3=0 switches it off.
#pragma once
#include "bmu_tc_kernel.cuh"

namespace pixie {

using namespace ptx;

constexpr int kX3Groups = 4;
constexpr int kX3Stages = 8;  // two per group: the tile being resolved and the one being converted

// |score - exact| <= kX3Repr * ||x|| wmax + kX3Acc * (wmax^2 + 2 ||x|| wmax)
//   representation (proven): with 11 significant bits read per operand, |xl| <= 2^-10 |x|, what
//       the tensor core drops of xl is < 2^-21 |x|, likewise for w'; the three lost terms
//       xl.wl + xh.(wl - tf32(wl)) + (xl - tf32(xl)).wh are <= (2^-20 + 2 * 2^-21) sum |x_c||w'_c|
//       <= 2^-19 ||x|| ||w'|| = 2^-18 ||x|| ||w||.
//   accumulation (model + measurement): 13 MMAs of 8 exact products each into an fp32
//       accumulator whose partial sums are bounded by M = wmax^2 + 2 ||x|| wmax.  If every MMA
//       truncates once and aligns its 8 addends with >= 24 bits kept, the error is <= 13 * 5 ulp
//       = 65 * 2^-23 M < 2^-17 M.  Measured (scripts/x3_margin.py, 90 M rows of five data sets
//       incl. codebooks of near-identical node pairs): labels stay identical to the exact kernel
//       with the window shrunk to 2^-22 M and first differ at 2^-24 M, i.e. the real error is
//       ~1 ulp of M; 2^-17 keeps a 32-fold margin over the smallest window that still held.
constexpr float kX3Repr = 3.814697265625e-6f;   // 2^-18
constexpr float kX3Acc = 7.62939453125e-6f;     // 2^-17

template <int SL, int SPC>
__global__ void __launch_bounds__(kX3Groups * 128 + 64, 1)
bmu_x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ TcParams p)
{
    constexpr int NG = kX3Groups;
    constexpr int NEPI = NG * 4;
    constexpr int NCHUNK = SL * SPC;
    constexpr int NMMA = (NCHUNK + 15) / 16 * 16;
    constexpr int NS = SPC;
    constexpr int NW = (SL + 31) / 32;
    constexpr uint32_t NST = kX3Stages;
    static_assert(NG * NMMA <= 512, "four accumulator buffers must fit in TMEM");

    // K = 97..104 needs every byte a CTA can have, so the 1 KiB alignment the swizzled tiles need
    // is asked of the declaration instead of being padded for; should a toolchain not honour it,
    // the launch fails (trap) rather than run past the allocation.
    extern __shared__ __align__(1024) uint8_t smem_x3[];
    const TcPlan &pl = p.plan;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t raw_u32 = smem_u32(smem_x3);
    const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
    if (pad + pl.smem_need > pl.smem_bytes) __trap();
    uint8_t *smem = smem_x3 + pad;
    const uint32_t sbase = raw_u32 + pad;
    uint8_t *ws = smem;  // codebook image: hi block (full-precision -2 W), lo block, bias block
    uint8_t *ones = smem + pl.off_ones;
    uint8_t *xs0 = smem + pl.off_x;
    const uint32_t bar0 = sbase + pl.off_bar;
    const uint32_t bar_full = bar0;               // [8]  X tile landed (TMA)
    const uint32_t bar_empty = bar0 + 64u;        // [8]  X tile no longer needed (4 warps)
    const uint32_t bar_tfull = bar0 + 128u;       // [4]  scores of the group's tile complete (commit)
    const uint32_t bar_tempty = bar0 + 160u;      // [4]  accumulator buffer drained (4 warps)
    const uint32_t bar_xl = bar0 + 192u;          // [4]  low parts of the group's next tile written
    const uint32_t bar_w = bar0 + 224u;           //      codebook image landed
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + pl.off_bar + 232u);
    static_assert(236 <= kBarBlock, "barrier block");

    if (warp == NEPI && lane == 0) {
        prefetch_tensormap(&tmX);
        for (uint32_t s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8u * s, 1);
            mbar_init(bar_empty + 8u * s, 4);
        }
        for (uint32_t g = 0; g < (uint32_t)NG; ++g) {
            mbar_init(bar_tfull + 8u * g, 1);
            mbar_init(bar_tempty + 8u * g, 4);
            mbar_init(bar_xl + 8u * g, 4);
        }
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    if (warp == NEPI + 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), (uint32_t)pl.tmem_cols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<float *>(ones)[i] = 1.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's tiles: j = blockIdx.x + it * gridDim.x, it < cnt; tile it goes to group it % 4,
    // X stage it % 8
    const uint32_t cnt = (int64_t)blockIdx.x < p.ntiles
                             ? (uint32_t)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    uint32_t st_flag = 0, st_pairs = 0, st_fp64 = 0, st_fix = 0;

    if (warp == NEPI) {
        // ============================================================ TMA producer
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_w, pl.wimg_bytes);
            for (uint32_t off = 0; off < pl.wimg_bytes; off += 16384u) {
                const uint32_t sz = min(16384u, pl.wimg_bytes - off);
                bulk_load(sbase + off, reinterpret_cast<const uint8_t *>(p.wimg) + off, sz, bar_w);
            }
        }
        __syncwarp();
        for (uint32_t it = 0; it < cnt; ++it) {
            const uint32_t s = it & (NST - 1u);
            mbar_wait(bar_empty + 8u * s, ((it >> 3) & 1u) ^ 1u);
            const int64_t j = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int32_t row0 = (int32_t)((p.tile_first + j * p.tile_stride) * kTile);
            const bool ahead = it + NST < cnt && !(p.dbg_flags & 8);
            const int64_t ja = (int64_t)blockIdx.x + (int64_t)(it + NST) * gridDim.x;
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full + 8u * s, pl.stage_bytes);
                tma_load_2d(sbase + pl.off_x + s * pl.stage_bytes, &tmX, bar_full + 8u * s, 0, row0,
                            kEvictFirst);
                if (ahead)
                    tma_prefetch_2d(&tmX, 0, (int32_t)((p.tile_first + ja * p.tile_stride) * kTile));
            }
            __syncwarp();
        }
    } else if (warp == NEPI + 1) {
        // ============================================================ MMA issuer
        // Descriptors are built once and advanced by adding byte offsets >> 4 to their address
        // field (all operands sit below 256 KiB, the field never carries out).
        constexpr uint32_t idesc = umma_idesc_tf32(128, (uint32_t)NMMA);
        const uint64_t desc_ones = umma_desc_nosw(sbase + pl.off_ones, 128u, 256u);
        const uint64_t desc_bias = umma_desc_nosw(sbase + pl.off_bias, 128u, 256u);
        const uint64_t d_wh = umma_desc_sw128(sbase);
        const uint64_t d_wl = umma_desc_sw128(sbase + pl.off_wlo);
        const uint64_t d_x0 = umma_desc_sw128(sbase + pl.off_x);
        const uint64_t d_xl0 = umma_desc_sw128(sbase + pl.off_xl);
        const uint64_t stage16 = (uint64_t)(pl.stage_bytes >> 4);
        const int ksteps = pl.ksteps;
        mbar_wait(bar_w, 0u);
        for (uint32_t it = 0; it < cnt; ++it) {
            const uint32_t g = it & 3u, n = it >> 2, s = it & (NST - 1u);
            // One wait per tile.  (Issuing the eight high-part products as soon as the tile has
            // landed and the accumulator is drained, and the rest after the conversion, was
            // measured SLOWER -- 1.42 -> 1.69 ms at C = 16: the in-order issuer then sits in the
            // second wait while other groups are ready.)
            mbar_wait(bar_xl + 8u * g, n & 1u);              // low parts written => X landed too
            mbar_wait(bar_tempty + 8u * g, (n & 1u) ^ 1u);   // (complete by then: drained first)
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d_tmem = tmem_base + g * (uint32_t)NMMA;
                const uint64_t d_xh = d_x0 + (uint64_t)s * stage16;
                const uint64_t d_xl = d_xl0 + (uint64_t)g * stage16;
#pragma unroll 4
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t ko = (uint64_t)(2 * ks);  // 32 bytes per K-step, >> 4
                    mma_tf32(d_tmem, d_xh + ko, d_wh + ko, idesc, ks > 0 ? 1u : 0u);
                    mma_tf32(d_tmem, d_xh + ko, d_wl + ko, idesc, 1u);
                    mma_tf32(d_tmem, d_xl + ko, d_wh + ko, idesc, 1u);
                }
                mma_tf32(d_tmem, desc_ones, desc_bias, idesc, 1u);
                mma_commit(bar_tfull + 8u * g);
            }
            __syncwarp();
        }
    } else {
        // ============================================================ epilogue groups
        const int g = warp >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t r7 = (uint32_t)(row & 7);
        const int Nrows = pl.Ntot;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * NMMA);
        mbar_wait(bar_w, 0u);
        const float wmax = __int_as_float(p.ctl->wmax_bits);
        const float wmax2 = wmax * wmax;
        const float eps32 = (float)(pl.C + 8) * 2.4e-7f;  // relative error of the fp32 distances
        const uint32_t xl_row = sbase + pl.off_xl + (uint32_t)g * pl.stage_bytes +
                                (uint32_t)row * 128u + (r7 << 4);

        // low parts of this thread's row of stage s -> the group's buffer; returns the candidate
        // window of the row.  Step pc touches physical chunk (row & 7) ^ pc of both tiles, so the
        // 8 lanes of a quarter-warp hit 8 different 16-byte bank groups; chunk order is irrelevant
        // to a sum of squares.  Channels past C are zero (TMA fill) and stay zero.
        auto convert = [&](uint32_t s) -> float {
            const uint32_t xaddr = sbase + pl.off_x + s * pl.stage_bytes + (uint32_t)row * 128u +
                                   (r7 << 4);
            uint64_t xa = 0ull, xb = 0ull;
#pragma unroll
            for (uint32_t pc = 0; pc < 8; ++pc) {
                const uint4 x = lds128(xaddr ^ (pc << 4));
                const uint64_t lo = pack2u(x.x, x.y), hi = pack2u(x.z, x.w);
                xa = fma2(lo, lo, xa);
                xb = fma2(hi, hi, xb);
                uint32_t l0, l1, l2, l3;
                unpack2u(sub2(lo, pack2u(x.x & 0xFFFFE000u, x.y & 0xFFFFE000u)), l0, l1);
                unpack2u(sub2(hi, pack2u(x.z & 0xFFFFE000u, x.w & 0xFFFFE000u)), l2, l3);
                sts128(xl_row ^ (pc << 4), l0, l1, l2, l3);
            }
            fence_proxy_async();  // the tensor core reads the buffer through the async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xl + 8u * (uint32_t)g);
            float a, b, c, d;
            unpack2(xa, a, b);
            unpack2(xb, c, d);
            const float xn = sqrtf((a + b) + (c + d) + 1.0e-30f) * 1.000001f;
            const float E = kX3Repr * xn * wmax + kX3Acc * (wmax2 + 2.0f * xn * wmax) + 1.0e-30f;
            return 2.0f * E * p.delta_scale;  // the error is two-sided
        };

        uint32_t it = (uint32_t)g, n = 0;
        float delta = 0.f;
        if (it < cnt) {
            mbar_wait(bar_full + 8u * it, 0u);
            delta = convert(it);
        }
        for (; it < cnt; it += (uint32_t)NG, ++n) {
            const uint32_t s = it & (NST - 1u);
            const uint8_t *xs = xs0 + s * pl.stage_bytes;
            const int64_t j = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t grow = (p.tile_first + j * p.tile_stride) * kTile + row;
            int32_t *lab_ptr = p.labels + (p.compact_labels ? j * kTile + row : grow);

            // ---- scores: slice minimum, then the bit mask of the values inside the window
            mbar_wait(bar_tfull + 8u * (uint32_t)g, n & 1u);
            tc_fence_after();
            float m_run = __int_as_float(0x7f800000);
            uint32_t mw[NS][NW];
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int w = 0; w < NW; ++w) mw[a][w] = 0u;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                uint32_t vr[SL];
                tmem_ld_cols<SL>(tmem_lane + (uint32_t)(sl * SL), vr);
                tc_wait_ld();
                if (sl == NS - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8u * (uint32_t)g);
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
                if (sl > 0) {
                    const bool drop = m_new + delta < m_run;  // earlier candidates fall out
#pragma unroll
                    for (int a = 0; a < NS; ++a)
                        if (a < sl)
#pragma unroll
                            for (int w = 0; w < NW; ++w) mw[a][w] = drop ? 0u : mw[a][w];
                }
                m_run = m_new;
                const float thr = m_run + delta;
                if (sl == 0 || __any_sync(0xffffffffu, ms < thr)) {
                    // sign bit of (v - thr) funnel-shifted into the mask; value i of word w ends
                    // at bit (cnt_w - 1 - (i - 32 w))
                    const uint64_t thr2 = pack2(thr, thr);
#pragma unroll
                    for (int i = 0; i < SL; i += 2) {
                        uint32_t d0, d1;
                        unpack2u(sub2(pack2u(vr[i], vr[i + 1]), thr2), d0, d1);
                        mw[sl][i >> 5] = __funnelshift_l(d0, mw[sl][i >> 5], 1);
                        mw[sl][(i + 1) >> 5] = __funnelshift_l(d1, mw[sl][(i + 1) >> 5], 1);
                    }
                }
            }

            // ---- the group's next tile: low parts + window, so its MMAs run during the resolve
            float delta_next = 0.f;
            if (it + (uint32_t)NG < cnt) {
                const uint32_t it2 = it + (uint32_t)NG;
                mbar_wait(bar_full + 8u * (it2 & (NST - 1u)), (it2 >> 3) & 1u);
                delta_next = convert(it2 & (NST - 1u));
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
                        const int cw = (SL - 32 * w) < 32 ? (SL - 32 * w) : 32;
                        if (mw[a][w]) idx = a * SL + 32 * w + cw - 32 + __clz(mw[a][w]);
                    }
                label = idx + 1;
            } else if (finite && nc == 2) {
                // two candidates (nearly every flagged row): both fp32 distances in one walk over
                // the row; the fp64 replica only if they are within the fp32 error of each other
                ++st_flag;
                st_pairs += 2u;
                int c0 = -1, c1 = -1;  // ascending node index
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
                if (c0 >= Nrows) c0 = 0;
                if (c1 >= Nrows || c1 < 0) c1 = 0;
                float d0, d1;
                duel_dist2_f32(xs, ws, Nrows, pl.C8 >> 2, row, c0, c1, d0, d1);
                const float bound = fminf(d0, d1) * (1.0f + eps32) + 1.0e-30f;
                if (d1 > bound) {
                    label = c0 + 1;
                } else if (d0 > bound) {
                    label = c1 + 1;
                } else {
                    ++st_fp64;
                    if (c0 < pl.K && c1 < pl.K) {
                        const double e0 = pair_dist_f64(xs, ws, Nrows, pl.C, row, c0);
                        const double e1 = pair_dist_f64(xs, ws, Nrows, pl.C, row, c1);
                        label = (e1 < e0 ? c1 : c0) + 1;
                        if (!(e0 == e0) || !(e1 == e1)) label = kLabelFixup;
                    }
                }
            } else if (finite && nc >= 3 && nc <= kMaxCand) {
                // the reference loop itself over the candidates, ascending node order, strict <
                ++st_flag;
                ++st_fp64;
                st_pairs += (uint32_t)nc;
                double bestd = DBL_MAX;
                int bestk = -1;
                bool isnan = false;
#pragma unroll
                for (int a = 0; a < NS; ++a)
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const int cw = (SL - 32 * w) < 32 ? (SL - 32 * w) : 32;
                        uint32_t m = mw[a][w];
                        while (m) {
                            const int lz = __clz(m);
                            m &= ~(0x80000000u >> lz);
                            const int k = a * SL + 32 * w + cw - 32 + lz;
                            if (k < pl.K) {
                                const double d = pair_dist_f64(xs, ws, Nrows, pl.C, row, k);
                                if (!(d == d)) isnan = true;
                                if (d < bestd) {
                                    bestd = d;
                                    bestk = k;
                                }
                            }
                        }
                    }
                label = (bestk >= 0 && !isnan) ? bestk + 1 : kLabelFixup;
            }
            if (label > pl.K) label = kLabelFixup;  // a padded codebook row can only win on garbage
            if (grow < p.n) {
                if (label == kLabelFixup) {
                    ++st_fix;
                    atomicAdd(&p.ctl->fixup_count, 1);
                }
                *lab_ptr = label;
            } else if (p.compact_labels) {
                *lab_ptr = 0;  // padding row of the last tile
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8u * s);
            delta = delta_next;
        }
    }

    if (warp < NEPI && p.stats) {
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

    tc_fence_before();
    __syncthreads();
    if (warp == NEPI + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
    }
}

template <int SL, int SPC>
static cudaError_t launch_x3_variant(const CUtensorMap &tmX, const TcParams &p, int grid,
                                     cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(bmu_x3_kernel<SL, SPC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.plan.smem_bytes);
    if (e != cudaSuccess) return e;
    bmu_x3_kernel<SL, SPC><<<grid, kX3Groups * 128 + 64, p.plan.smem_bytes, stream>>>(tmX, p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
