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
//   label and are resolved by the exact fp64 kernel (som_kernels.cu) launched right behind.
//
// Warp roles (NG*128 + 64 threads, 1 CTA per SM, persistent over tiles):
//   warps 0 .. 4*NG-1 : NG epilogue groups of 4 warps; group g owns tiles g, g+NG, ... of the CTA
//                       (thread <-> tile row <-> TMEM lane; warp w reads lane quadrant w % 4)
//   warp 4*NG         : TMA producer (one elected lane)
//   warp 4*NG+1       : TMEM allocator + MMA issuer (one elected lane)
// NG = 4 with 56-column slices (<= 112 registers/thread) for K <= 128, NG = 2 with wider slices
// above that.  The epilogue is instruction-latency bound, so warps per scheduler matter.
//
// Two relatives of this kernel, both planned here:
//   * bmu_x3_kernel.cuh -- assignment for K <= 104, C <= 24 with SPLIT tf32 operands (three MMAs per
//     K-step): the candidate window is ~100x narrower and the recheck all but disappears;
//   * the "tail8" layout (TcPlan::tail8, a template parameter of bmu_tc_kernel) -- for C8 % 32 == 8
//     the last eight channels of the X tile and of the image are 32-byte SWIZZLE_32B rows instead
//     of a 128-byte block that is three quarters zero fill (cfg3: 2 -> 6 pipeline stages).
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include "bmu_tc_kernel.cuh"
#include "som_update.cuh"

namespace pixie {

using namespace ptx;

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
namespace {
struct Variant {
    int SL, spc, NCH, NG;
};
// every entry has a template instantiation in launch_bmu_tc()
const Variant kVariants[] = {
    {32, 1, 1, 4}, {32, 2, 1, 4}, {48, 2, 1, 4}, {50, 2, 1, 4}, {56, 2, 1, 4}, {64, 2, 1, 4},
    {32, 1, 1, 2}, {32, 2, 1, 2}, {48, 2, 1, 2}, {50, 2, 1, 2}, {56, 2, 1, 2}, {64, 2, 1, 2},
    {80, 2, 1, 2}, {100, 2, 1, 2}, {104, 2, 1, 2}, {128, 2, 1, 2},
    {80, 2, 2, 2}, {100, 2, 2, 2}, {104, 2, 2, 2}, {128, 2, 2, 2},
    // ({40,52,64}, 2, 4, 4) -- four chunks per tile through ONE buffer per group, four groups for
    // K up to 512 -- exist in the kernel template, are bit-exact, and were measured 4-8 % SLOWER than
    // the two-group variants (profiles/r02_notes.md): not instantiated.
};
}  // namespace

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && v[0]) ? atoi(v) : dflt;
}

TcPlan make_tc_plan(int C, int K, bool acc)
{
    TcPlan best{};
    best.ok = false;
    if (C < 1 || C > 128 || K < 1 || K > 512) return best;
    const int cap_stages = env_int("PIXIE_TC_STAGES", kMaxStages);
    // experiments: PIXIE_TC_VARIANT=SL,SPC,NCH,NG restricts the choice to one variant
    int only[4] = {0, 0, 0, 0};
    if (const char *ov = getenv("PIXIE_TC_VARIANT"))
        if (sscanf(ov, "%d,%d,%d,%d", &only[0], &only[1], &only[2], &only[3]) != 4) only[0] = 0;
    long best_cost = -1;
    for (const Variant &v : kVariants) {
        if (only[0] && (v.SL != only[0] || v.spc != only[1] || v.NCH != only[2] || v.NG != only[3]))
            continue;
        TcPlan p{};
        p.C = C;
        p.K = K;
        p.C8 = (C + 7) / 8 * 8;
        p.ksteps = p.C8 / 8;
        p.nblkX = (C + 31) / 32;
        p.nblkW = (p.C8 + 31) / 32;
        p.SL = v.SL;
        p.spc = v.spc;
        p.NCH = v.NCH;
        p.NG = v.NG;
        p.Nchunk = v.SL * v.spc;
        if (v.NCH > 1 && p.Nchunk % 8) continue;  // chunk base must stay on a swizzle-atom row
        p.Nmma = (p.Nchunk + 15) / 16 * 16;
        p.Ntot = (v.NCH - 1) * p.Nchunk + p.Nmma;
        if (p.Nchunk * v.NCH < K) continue;
        p.nbuf = v.NCH == 2 ? 2 : v.NG;
        const int need = p.nbuf * p.Nmma;
        if (need > 512) continue;
        p.tmem_cols = 32;
        while (p.tmem_cols < need) p.tmem_cols *= 2;
        // (assignment only: in train mode the stage count stays at one per group anyway -- the
        // L1 carve-out rule below -- and the tail8 code paths cost the fused kernel more than the
        // smaller tiles bring: cfg3 shard pass 22.6 ms without, 22.8 ms with, same box)
        p.tail8 = (!acc && p.C8 % 32 == 8 && p.nblkX >= 2 && p.nblkX == p.nblkW &&
                   env_int("PIXIE_TAIL8", 1) != 0) ? 1 : 0;
        p.stage_bytes = (uint32_t)p.nblkX * 16384u;
        p.off_bias = (uint32_t)p.nblkW * (uint32_t)p.Ntot * 128u;
        if (p.tail8) {
            p.x_tail_off = (uint32_t)(p.nblkX - 1) * 16384u;
            p.stage_bytes = p.x_tail_off + 4096u;
            p.w_tail_off = (uint32_t)(p.nblkW - 1) * (uint32_t)p.Ntot * 128u;
            p.off_bias = (p.w_tail_off + (uint32_t)p.Ntot * 32u + 1023u) / 1024u * 1024u;
        }
        p.wimg_bytes = p.off_bias + (uint32_t)p.Ntot * 32u;
        p.off_ones = (p.wimg_bytes + 1023u) / 1024u * 1024u;
        p.off_x = p.off_ones + 4096u;
        // fused accumulation: the groups' node counts and label rings (the tables themselves live
        // in global memory); the pair-list area doubles as the accumulate lists and as scratch of
        // the end-of-step update (som_update_nodes) and of the cross-GPU exchange
        const uint32_t pair_cap = acc ? kWarpPairCapAcc : kWarpPairCap;
        uint32_t pairs_bytes = (uint32_t)(4 * v.NG) * pair_cap * 8u;
        if (acc) {
            uint32_t need = som_update_scratch_bytes(C, K);
            const uint32_t slice = (uint32_t)((K * (C + 1) + kSumParts - 1) / kSumParts) * 8u;
            if (slice > need) need = slice;
            if (need > pairs_bytes) pairs_bytes = (need + 15u) / 16u * 16u;
        }
        const uint32_t limit = 227u * 1024u - 1024u;
        const uint32_t tables = (uint32_t)v.NG * (uint32_t)K * (uint32_t)tab_pitch(C) * 4u;
        const uint32_t counts = ((uint32_t)v.NG * (uint32_t)K * 4u + 15u) / 16u * 16u;
        const uint32_t rings = (uint32_t)v.NG * 512u;
        uint32_t acc_bytes = 0;
        p.tab_global = 0;
        if (acc) {
            const uint32_t fixed = p.off_x + (uint32_t)kBarBlock + pairs_bytes + counts + rings +
                                   (uint32_t)v.NG * p.stage_bytes;
            const int force = env_int("PIXIE_TAB_GLOBAL", -1);
            const bool fits = fixed + tables <= limit;
            if (force == 1 || !fits) {
                if (force == 0) continue;
                p.tab_global = 1;
                acc_bytes = counts + rings;
            } else {
                acc_bytes = tables + counts + rings;
            }
        }
        const uint32_t scratch = (uint32_t)kBarBlock + pairs_bytes + acc_bytes;
        if (p.off_x + scratch + (uint32_t)v.NG * p.stage_bytes > limit) continue;
        p.nstage = (int)((limit - p.off_x - scratch) / p.stage_bytes);
        if (p.nstage > kMaxStages) p.nstage = kMaxStages;
        if (p.nstage > cap_stages && cap_stages >= v.NG) p.nstage = cap_stages;
        // a multiple of NG: stage (it % nstage) is then always consumed by the same epilogue group
        // (it % NG), so no consumer can reach a full-barrier wait a phase early (parity aliasing).
        // (Measured: a free stage count -- 7 instead of 4 stages in train mode, made safe by an
        // extra wait for the stage's previous release -- bought nothing and cost assign 1.7 %.)
        p.nstage = p.nstage / v.NG * v.NG;
        if (acc) {
            // Train mode keeps loop state on the stack around the per-tile accumulate call (~370
            // bytes x 576 threads): with all of shared memory taken the L1 left over (228 KB
            // carve-out -> 28 KB) thrashes on it.  A pipeline one tile per group deep is enough here
            // (the producer runs a whole step ahead anyway), so stay under the 164 KB carve-out
            // when that is possible: cfg2 8 -> 4 stages, training pass 1.72 -> 1.56 ms.
            const uint32_t soft = env_int("PIXIE_TC_SOFTCAP", 1) ? 164u * 1024u - 1024u : limit;  // 0: experiments
            while (p.nstage > v.NG && p.off_x + scratch + (uint32_t)p.nstage * p.stage_bytes > soft)
                p.nstage -= v.NG;
        }
        p.off_bar = p.off_x + (uint32_t)p.nstage * p.stage_bytes;
        p.off_pairs = p.off_bar + (uint32_t)kBarBlock;
        p.pair_cap = (int)pair_cap;
        p.pairs_bytes = pairs_bytes;
        p.off_acc = p.off_pairs + pairs_bytes;
        p.off_cnt = p.off_acc + ((acc && !p.tab_global) ? tables : 0u);
        p.off_lab = p.off_cnt + (acc ? counts : 0u);
        p.acc = acc ? 1 : 0;
        p.smem_bytes = p.off_acc + acc_bytes + 1024u;
        p.ok = true;
        // fewest padded codebook rows first; then more epilogue groups; then tables on chip and a
        // deeper pipeline
        const long cost = (long)(p.Nchunk * v.NCH) * 1000 - v.NG * 10 - p.nstage +
                          (p.tab_global ? 15 : 0);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = p;
        }
    }
    return best;
}

// Split-operand assignment kernel (bmu_x3_kernel.cuh).  Shared memory, from the 1 KiB-aligned base:
// image = [hi block | lo block | bias block] (rows rounded up to 8, not to the MMA's N: the
// columns of the rows it reads past the block are never looked at), the barriers in the gap up to
// the next KiB, the ones tile, eight X stages, four low-part buffers.  At K = 100, C = 32 that is
// 231,424 bytes, all a CTA can have: hence the shape limits.
TcPlan make_x3_plan(int C, int K)
{
    TcPlan p{};
    p.ok = false;
    if (C < 1 || C > 32 || K < 1 || K > 104) return p;
    const int mode = env_int("PIXIE_X3", 1);  // 0 = never, 1 = where it is faster, 2 = wherever it fits
    if (mode == 0) return p;
    if (mode == 1 && C > 24) return p;  // four K-steps x 3 MMAs: tensor-pipe bound, slower than plain
    static const int kSl[][2] = {{32, 1}, {32, 2}, {48, 2}, {50, 2}, {52, 2}};
    int pick = -1;
    for (int i = 0; i < 5; ++i)
        if (kSl[i][0] * kSl[i][1] >= K) {
            pick = i;
            break;
        }
    if (pick < 0) return p;
    p.C = C;
    p.K = K;
    p.C8 = (C + 7) / 8 * 8;
    p.ksteps = p.C8 / 8;
    p.nblkX = 1;
    p.nblkW = 1;
    p.SL = kSl[pick][0];
    p.spc = kSl[pick][1];
    p.NCH = 1;
    p.NG = 4;
    p.Nchunk = p.SL * p.spc;
    p.Nmma = (p.Nchunk + 15) / 16 * 16;
    p.Ntot = (p.Nchunk + 7) / 8 * 8;  // image rows per block
    p.nbuf = 4;
    p.tmem_cols = 32;
    while (p.tmem_cols < 4 * p.Nmma) p.tmem_cols *= 2;
    p.stage_bytes = 16384u;
    p.x3 = 1;
    p.off_wlo = (uint32_t)p.Ntot * 128u;
    p.off_bias = 2u * p.off_wlo;
    p.wimg_bytes = p.off_bias + (uint32_t)p.Ntot * 32u;
    p.off_bar = (p.wimg_bytes + 15u) / 16u * 16u;
    p.off_ones = (p.off_bar + (uint32_t)kBarBlock + 1023u) / 1024u * 1024u;
    p.off_x = p.off_ones + 4096u;
    p.nstage = 8;
    p.off_xl = p.off_x + 8u * p.stage_bytes;
    p.smem_need = p.off_xl + 4u * p.stage_bytes;
    // 227 KiB per CTA minus the KiB the toolchain reserves (cuobjdump: SHARED:1024).  K = 100 needs
    // exactly this much: no alignment slack is left, the kernel traps if the base is not aligned.
    const uint32_t limit = 226u * 1024u;
    if (p.smem_need > limit) return p;
    p.smem_bytes = p.smem_need + 1024u > limit ? limit : p.smem_need + 1024u;
    p.ok = true;
    return p;
}

// ------------------------------------------------------------------------------------------------
// codebook preparation: W [K x C] fp32  ->  shared-memory image + norms
// ------------------------------------------------------------------------------------------------
// one warp per codebook row; lanes stride over the image columns (coalesced reads of W)
__global__ void __launch_bounds__(256)
codebook_prep_kernel(const float *__restrict__ W, int K, int C, int nblkW, int lay,
                     uint32_t off_bias, uint32_t off_wlo, float *__restrict__ wimg,
                     CodebookAux *__restrict__ aux)
{
    const int Ntot = lay & 0xFFFF;  // rows per block; the high half marks a tail8 last block
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= Ntot) return;
    const int tb = (lay >> 16) - 1;
    const int ncols = tb >= 0 ? tb * 32 + 8 : nblkW * 32;  // a tail8 block is 8 columns wide
    char *base = reinterpret_cast<char *>(wimg);
    double nrm2 = 0.0;
    bool bad = false, neg = false;
    for (int col = lane; col < ncols; col += 32) {
        float v = 0.f;
        if (row < K && col < C) {
            const float w = W[(size_t)row * C + col];
            if (!(fabsf(w) <= FLT_MAX)) bad = true;
            if (__float_as_int(w) < 0) neg = true;  // sign bit set (includes -0.0: conservative)
            nrm2 += (double)w * (double)w;
            v = -2.0f * w;
        }
        *reinterpret_cast<float *>(base + img_offset(lay, row, col)) = v;
        // split-operand kernel: what the tensor core drops from v (it reads the top 19 bits)
        if (off_wlo)
            *reinterpret_cast<float *>(base + off_wlo + img_offset(lay, row, col)) =
                v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
    for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(~0u, nrm2, o);
    bad = __any_sync(~0u, bad);
    neg = __any_sync(~0u, neg);
    if (lane == 0) {
        float bias = (row < K) ? (float)nrm2 : 1.0e30f;
        if (!(bias <= FLT_MAX)) bias = FLT_MAX;  // overflowed norms: row can never win anyway
        // three tf32-exact pieces (11 significant bits each): h + m + l == bias exactly
        const float h = __uint_as_float(__float_as_uint(bias) & 0xFFFFE000u);
        const float r1 = bias - h;
        const float m = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
        const float l = r1 - m;
        float *b = reinterpret_cast<float *>(base + off_bias + bias_offset(row, 0));
        b[0] = h;
        b[1] = m;
        b[2] = l;
        b[3] = 0.f;
        float *b2 = reinterpret_cast<float *>(base + off_bias + bias_offset(row, 4));
        b2[0] = b2[1] = b2[2] = b2[3] = 0.f;
        if (row < K) {
            if (bad) atomicOr(&aux->nonfinite, 1);
            if (neg) atomicOr(&aux->w_has_negative, 1);
            float nr = (float)sqrt(nrm2) * 1.0000005f;
            if (!(nr <= FLT_MAX)) nr = FLT_MAX;
            atomicMax(&aux->wmax_bits, __float_as_int(nr));  // non-negative floats order as ints
        }
    }
}

// `aux` must have been zeroed on the stream (capi.cu does it with the fix-up counter).
cudaError_t launch_codebook_prep(const float *W, int K, int C, const TcPlan &plan, float *wimg,
                                 CodebookAux *aux, cudaStream_t stream)
{
    const int rows_per_block = 8;
    codebook_prep_kernel<<<(plan.Ntot + rows_per_block - 1) / rows_per_block, 256, 0, stream>>>(
        W, K, C, plan.nblkW, lay_pack(plan.Ntot, plan.tail8 ? plan.nblkW - 1 : -1), plan.off_bias,
        plan.x3 ? plan.off_wlo : 0u, wimg, aux);
    count_launch();
    return cudaGetLastError();
}


// variant families, one translation unit each (bmu_tc_inst_*.cu)
cudaError_t launch_tc_family_acc(const CUtensorMap &tmX, const TcParams &p, int grid, cudaStream_t s);
cudaError_t launch_tc_family_plain(const CUtensorMap &tmX, const TcParams &p, int grid, cudaStream_t s);

cudaError_t launch_bmu_tc(const CUtensorMap &tmX, const TcParams &p, int num_sms,
                          cudaStream_t stream)
{
    if (p.ntiles <= 0) return cudaSuccess;
    int grid = num_sms;
    if ((int64_t)grid > p.ntiles) grid = (int)p.ntiles;
    if (p.partials != nullptr) return launch_tc_family_acc(tmX, p, grid, stream);
    return launch_tc_family_plain(tmX, p, grid, stream);
}

}  // namespace pixie
