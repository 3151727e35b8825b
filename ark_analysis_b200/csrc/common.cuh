// common.cuh -- shared declarations of the Pixie B200 kernels (internal; the public surface is
// include/pixie_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/pixie_b200.h"

namespace pixie {

constexpr int kTile = PIXIE_TILE;  // rows per tile == UMMA M == TMEM lanes
constexpr int kLabelFixup = -1;    // sentinel: row must be resolved by the exact fix-up kernel
constexpr int kMaxCand = 15;       // candidate nodes kept per row before giving up to the fix-up
constexpr int kWarpPairCap = 256;  // (row, node) pairs re-evaluated per tile by one epilogue warp
constexpr int kWarpPairCapAcc = 128;  // ... in train mode, where shared memory also holds the sums
                                      // (the buffer doubles as the accumulate list: 128 x 8 bytes)
constexpr int kMaxStages = 8;
constexpr int kBarBlock = 256;     // bytes of shared memory reserved for mbarriers

// Device-side result of the codebook preparation kernel, read by the BMU kernel.
struct CodebookAux {
    int wmax_bits;  // float bits of max_k ||w_k|| (atomicMax on the non-negative float pattern)
    int nonfinite;
    int fixup_count;  // rows the tensor-core kernel handed to the exact fix-up kernel
    int w_has_negative;  // any codebook entry with the sign bit set (disables the one-sided bound)
    unsigned int sums_sync[2];  // grid barrier of the cluster-sums kernel (self-resetting)
    // whole-pass kernel: codebook norms/flags of step t live in slot t & 1 (written by step t-1)
    int pp_wmax_bits[2];
    int pp_w_has_negative[2];
    unsigned int grid_sync;  // its grid-barrier counter (zeroed per launch, only grows)
    int pad[3];
    // whole-pass kernel, CTA 0: nanoseconds spent in [tiles, barrier 1, fold, barrier 2, update,
    // barrier 3], summed over steps (diagnostics; scripts/prof_train_pass.py reads them)
    unsigned long long phase_ns[6];
    // PIXIE_PROFILE builds only (make prof): SM cycles warp 0 of CTA 0 spent per tile in
    // [wait X, norm pass, accumulator wait + two passes, resolve, fix-ups, accumulate work], the
    // number of tiles it handled, and [7] the whole accumulate call (group barrier wait + work)
    unsigned long long tile_cyc[8];
};
static_assert(sizeof(CodebookAux) <= 256, "the workspace reserves 256 bytes for the control block");

// Host-side plan of the tensor-core BMU kernel for one (C, K).
struct TcPlan {
    bool ok;
    int C, K;
    int C8;      // C rounded up to the tf32 MMA K-step (8)
    int ksteps;  // C8 / 8 data K-steps (+1 bias K-step)
    int nblkX;   // 32-channel (128-byte) blocks of one X tile  = ceil(C / 32)
    int nblkW;   // 32-column blocks of the codebook image      = ceil(C8 / 32)
    int SL;      // TMEM columns per epilogue slice            (template parameter)
    int spc;     // slices per accumulator chunk               (template parameter)
    int NCH;     // accumulator chunks per tile, 1 or 2        (template parameter)
    int NG;      // epilogue groups of 4 warps, 2 or 4         (template parameter)
    int Nchunk;  // codebook rows covered by one accumulator chunk = SL * spc
    int Nmma;    // UMMA N = Nchunk rounded up to 16 (<= 256); columns past Nchunk are never read
    int Ntot;    // codebook rows in the image = (NCH - 1) * Nchunk + Nmma (rows >= K never win)
    int nbuf;    // TMEM accumulator buffers: NG when NCH == 1, else 2
    int tmem_cols;
    int nstage;  // X tile pipeline depth (multiple of NG)
    uint32_t stage_bytes, wimg_bytes;
    uint32_t off_bias;  // bias block inside the image: Ntot rows x 8 columns, no-swizzle core matrices
    uint32_t off_ones, off_x, off_bar, off_pairs;
    uint32_t off_acc, off_cnt, off_lab;  // fused accumulation (train mode), at off_acc: NG x K x
                                // tab_pitch(C) fp32 group tables when they fit (tab_global == 0),
                                // NG x K int32 node counts (off_cnt), NG x 512 bytes of label rings
    int tab_global;             // 1 = the group tables live in global memory (TcParams::partials)
    int acc;                    // 1 = this plan has room for the fused accumulation
    int pair_cap;               // pair-list capacity per epilogue warp
    uint32_t pairs_bytes;       // bytes at off_pairs (pair lists; scratch of the end-of-step work)
    uint32_t smem_bytes;  // dynamic shared memory to request (includes 1 KiB alignment slack)
    // split-operand assignment kernel (bmu_x3_kernel.cuh; make_x3_plan): the image holds a second
    // set of blocks with the low parts of -2 W at off_wlo, and every epilogue group owns a buffer
    // for the low parts of its next X tile at off_xl + g * stage_bytes
    // tail8: C8 % 32 == 8 (C = 33..40, 65..72, 97..104) -- the last 32-channel block of the X tile
    // and of the codebook image would be three quarters zero fill.  It is kept as 32-byte rows
    // instead (one tf32 K-step per row, SWIZZLE_32B: 16-byte chunk c of row r at c ^ ((r >> 2) & 1)),
    // loaded by a second tensor map: X stage = full blocks + 4 KiB at x_tail_off, image = full
    // blocks + Ntot x 32 bytes at w_tail_off.  At C = 40, K = 400 (cfg3) that is 20 KiB instead of
    // 32 per stage and 65 instead of 104 KiB of image: six pipeline stages instead of two.
    int tail8;
    uint32_t x_tail_off, w_tail_off;
    int x3;
    uint32_t off_wlo, off_xl;
    uint32_t smem_need;   // bytes used from the 1 KiB-aligned base (smem_bytes - alignment slack)
};

// Train-mode statistics.  Global tables (group tables that do not fit on chip, CTA parts):
// [K][part_pitch(C)] fp32 -- channel sums, the count in column C.  Shared-memory group tables:
// [K][tab_pitch(C)] fp32, sums only.  Both pitches are multiples of four floats: the accumulate
// moves whole 16-byte chunks.
__host__ __device__ inline int part_pitch(int C) { return (C + 1 + 3) & ~3; }
__host__ __device__ inline int tab_pitch(int C) { return (C + 3) & ~3; }

// acc = true: a plan for the fused per-node sums (train mode): per-group fp32 tables in shared
// memory when they fit beside the pipeline, else in global memory (tab_global).  PIXIE_TC_STAGES
// caps the pipeline depth, PIXIE_TAB_GLOBAL=1/0 forces / forbids global tables (environment,
// experiments only).
TcPlan make_tc_plan(int C, int K, bool acc = false);
// Plan of the split-operand ("3 x tf32") assignment kernel: K <= 104 and C <= 24 (C <= 32 fits in
// shared memory -- beside eight X stages and four low-part buffers -- but is slower than the plain
// kernel; PIXIE_X3=2 takes it anyway, PIXIE_X3=0 never: A/B runs and tests).
TcPlan make_x3_plan(int C, int K);

struct TcParams {
    int64_t n;             // rows of X
    int64_t tile_first;    // first tile visited
    int64_t tile_stride;   // distance between visited tiles
    int64_t ntiles;        // number of tiles visited by this launch
    const float *wimg;     // prepared codebook image (global)
    int32_t *labels;       // labels[row] (assign) or labels[j * 128 + r] (compact, accum)
    int compact_labels;
    unsigned long long *stats;  // may be null
    CodebookAux *ctl;           // fixup_count lives here
    // fused per-node sums (partials == null: plain assignment).  partials: the group tables,
    // [grid x NG] tables accumulated into by red.global.add during the tiles; parts: [grid] tables,
    // each CTA's groups added up at the end of a step; SN [K][C+1] fp64: the fold of all parts
    // (behind a grid barrier on ctl->grid_sync)
    float *partials;
    float *parts;
    double *SN;
    // whole-pass mode (nsteps > 1 or apply != 0): the kernel runs `nsteps` mini-batch steps itself,
    // step t visiting the tiles whose GLOBAL index is congruent to (t0 + t) % B, and after each
    // step folds the sums, applies the batch update to W64/W32 and rewrites the codebook image.
    int nsteps;            // 0/1 = single step described by tile_first/tile_stride/ntiles
    int apply;             // 1 = fold + apply + re-prep inside the kernel
    int B;                 // mini-batches per pass
    int t0;                // index of the first step (for the schedule)
    int T;                 // total steps of the run (schedule denominator)
    int xdim, ydim;
    int64_t tile_offset;   // global tile index of this shard's tile 0
    int64_t tiles_total;   // ceil(n / 128)
    double a0, a1, r0, r1; // learning-rate and radius ranges
    double *W64;           // [K x C] master codebook
    float *W32;            // [K x C] fp32 copy
    float *wimg_rw;        // the codebook image again, writable (rows < K are rewritten per step)
    // multi-GPU whole-pass mode: every rank's exchange buffer (layout: exchange_slice() in
    // bmu_tc_kernel.cuh; size: pixie_peer_buffer_bytes), mapped into this process (NVLink peer memory)
    int world, rank;
    uint32_t flag_base;    // flags only grow: step st of this launch signals flag_base + st + 1
    double *peer_buf[8];
    int dbg_step0;         // PIXIE_PROFILE builds: first of the two steps whose events are traced
    unsigned long long *trace;  // ... into this buffer (null in production builds)
    unsigned int *trace_count;
    int dbg_flags;         // experiments (PIXIE_DBG_FLAGS); 4 = per-step printf in PIXIE_PROFILE builds, 8 = no L2 prefetch
    float delta_scale;     // 1 in production; tests shrink the candidate window to probe its margin
    TcPlan plan;
    alignas(64) CUtensorMap tm_tail;  // plan.tail8: channels 32 (nblk - 1) .. + 7, box 8 x 128, SWIZZLE_32B
};

// launchers (each enqueues on `stream` and returns the cudaError_t of the launch)
cudaError_t launch_codebook_prep(const float *W, int K, int C, const TcPlan &plan, float *wimg,
                                 CodebookAux *aux, cudaStream_t stream);
cudaError_t launch_bmu_tc(const CUtensorMap &tmX, const TcParams &p, int num_sms,
                          cudaStream_t stream);
cudaError_t launch_bmu_x3(const CUtensorMap &tmX, const TcParams &p, int num_sms,
                          cudaStream_t stream);
cudaError_t launch_bmu_exact(const float *X, int64_t n, int C, int64_t ldX, const float *W, int K,
                             int32_t *labels, int64_t tile_first, int64_t tile_stride,
                             int64_t ntiles, int compact_labels, const int *fixup_count_or_null,
                             double *SN_add_or_null, cudaStream_t stream);
cudaError_t launch_bmu_dist(const float *X, int64_t n, int C, int64_t ldX, const float *W, int K,
                            const int32_t *labels, double *dists, cudaStream_t stream);
cudaError_t launch_cluster_sums(const float *X, int64_t n, int C, int64_t ldX,
                                const int32_t *labels, int compact_labels, int K,
                                int64_t tile_first, int64_t tile_stride, int64_t ntiles,
                                float *partials, int nparts, double *SN, unsigned int *ticket,
                                cudaStream_t stream);
cudaError_t launch_som_apply(double *W64, float *W32, const double *SN, int xdim, int ydim, int C,
                             double sigma, double alpha, cudaStream_t stream);

cudaError_t launch_columns_to_rows(const double *cols, int64_t col_stride, int64_t n, int C,
                                   const double *divisor, float *X, int64_t ldX, int num_sms,
                                   cudaStream_t stream);

cudaError_t launch_label_histogram(const int32_t *seg, const int32_t *clu, int64_t n, int32_t n_seg,
                                   int32_t n_clu, int32_t *counts, unsigned long long *bad,
                                   int num_sms, cudaStream_t stream);
cudaError_t launch_scatter_labels(const int32_t *row_index, const int32_t *col_index,
                                  const int32_t *clu, int64_t n, const int16_t *id_map,
                                  int32_t map_len, int32_t H, int32_t W, int16_t *img,
                                  int32_t *winner, unsigned long long *bad, int num_sms,
                                  cudaStream_t stream);

constexpr int kMaxBlurRadius = 32;  // gaussian taps either side (scipy: int(4 * sigma + 0.5))
size_t preprocess_scan_bytes(int64_t n);
cudaError_t launch_preprocess(const void *img, int img_is_f64, int H, int W, int C, const double *norm,
                              const double *taps_host, int radius, double thresh,
                              const int32_t *seg, double *blurred, double *tmp, double *rowsum,
                              int32_t *flags, int32_t *pos, void *scan_tmp, size_t scan_bytes,
                              double *X64, float *X32, int64_t ldX32, int32_t *row_index,
                              int32_t *col_index, int32_t *labels_out, int64_t *n_kept,
                              int stop_after_blur, int num_sms, cudaStream_t stream);

size_t quantile_workspace_bytes(int C);
cudaError_t launch_column_quantile(const double *X, int64_t n, int C, int64_t ldX, double q,
                                   double *lo, double *hi, int64_t *m_out, void *workspace,
                                   int num_sms, cudaStream_t stream);

size_t som_online_smem_bytes(int C, int K);
cudaError_t launch_som_online(const float *X, int64_t n, int C, int64_t ldX, double *W, int xdim,
                              int ydim, const int64_t *sample_idx, int64_t niter,
                              int64_t n_per_pass, double a0, double a1, double r0, double r1,
                              long long *iters_done, cudaStream_t stream);

// process-wide count of kernels this library has launched (pixie_kernel_launches())
void count_launch(int n = 1);

// number of per-CTA partial buffers the cluster-sums kernel uses
constexpr int kSumParts = 148;
// partials folded per thread of an octet: ceil(kSumParts / 8)
constexpr int kFoldMax = (kSumParts + 7) / 8;

}  // namespace pixie
