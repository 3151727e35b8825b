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
constexpr int kWarpPairCapAcc = 256;  // ... in train mode, where shared memory also holds the sums
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
};

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
    uint32_t off_acc, off_lab;  // fused accumulation (train mode): NG x K x (C+1) fp32 tables, then
                                // NG x sort_stride bytes of per-group sort scratch (SortLayout)
    uint32_t sort_stride;
    int acc;                    // 1 = this plan has room for the fused accumulation
    int pair_cap;               // pair-list capacity per epilogue warp
    uint32_t smem_bytes;  // dynamic shared memory to request (includes 1 KiB alignment slack)
};

// Per-group scratch of the fused accumulation (train mode): the tile's 128 rows are counting-sorted
// by label so that the rows of a node are contiguous and their sum is a register accumulation.
//   hist32  uint32[bins]     byte w of word b = rows of warp w labelled b (bin K = rows to skip)
//   wbase   uint8 [4][bins]  sorted position of the first row of warp w labelled b
//   order   uint8 [128]      sorted position -> tile row
//   slab    uint16[128]      sorted position -> label bin
//   side    float [4][C+1]   sum (and count) of a warp's FIRST segment: its node may continue from
//   side_lab int[4]          the previous warp's range, so it is merged after a group barrier
struct SortLayout {
    uint32_t nbl, bins, off_wbase, off_order, off_slab, off_side, off_sidelab, bytes;
};
__host__ __device__ inline SortLayout sort_layout(int C, int K)
{
    SortLayout s;
    s.nbl = ((uint32_t)K + 1u + 31u) / 32u;  // bins scanned per lane
    s.bins = 32u * s.nbl;
    s.off_wbase = s.bins * 4u;
    s.off_order = s.off_wbase + 4u * s.bins;
    s.off_slab = s.off_order + 128u;
    s.off_side = s.off_slab + 256u;
    s.off_sidelab = s.off_side + (4u * (uint32_t)(C + 1) * 4u + 15u) / 16u * 16u;
    s.bytes = s.off_sidelab + 16u;
    return s;
}

// acc = true: also reserve per-group fp32 accumulators for the fused per-node sums (train mode).
// PIXIE_TC_STAGES (environment, experiments only) caps the pipeline depth.
TcPlan make_tc_plan(int C, int K, bool acc = false);

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
    // fused per-node sums (null = plain assignment): per-CTA partials [grid][K][C+1] fp32, folded
    // into SN [K][C+1] fp64 behind a grid barrier on ctl->sums_sync
    float *partials;
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
    // multi-GPU whole-pass mode: every rank's exchange buffer ([2][K x (C+1)] fp64 ping-pong + one
    // uint32 flag per source rank), mapped into this process (NVLink peer memory)
    int world, rank;
    uint32_t flag_base;    // flags only grow: step st of this launch signals flag_base + st + 1
    double *peer_buf[8];
    float delta_scale;     // 1 in production; tests shrink the candidate window to probe its margin
    TcPlan plan;
};

// launchers (each enqueues on `stream` and returns the cudaError_t of the launch)
cudaError_t launch_codebook_prep(const float *W, int K, int C, const TcPlan &plan, float *wimg,
                                 CodebookAux *aux, cudaStream_t stream);
cudaError_t launch_bmu_tc(const CUtensorMap &tmX, const TcParams &p, int num_sms,
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
