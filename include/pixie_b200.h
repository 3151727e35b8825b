/*
 * pixie_b200.h -- C ABI of libpixie_b200.so: the B200 (sm_100a) implementation of the Pixie SOM
 * hot path of angelolab/ark-analysis.
 *
 * The reference's arithmetic for this path is two functions of the third-party pyFlowSOM module,
 * bound at src/ark/phenotyping/cluster_helpers.py:14 and called at :106-109 (`som`) and :152-157
 * (`map_data_to_nodes`).  Every entry point below cites the reference interface it replaces.
 *
 * Conventions
 *  - plain C types only; no torch / C++ types cross this boundary.
 *  - "device" entry points take device pointers valid on the CURRENT CUDA device and a
 *    cudaStream_t passed as void* (NULL = default stream).  They only enqueue work (asynchronous,
 *    never synchronise, never allocate) and are re-entrant across streams provided each call is
 *    given its own workspace.
 *  - "host" entry points take host pointers, do their own staging and synchronise before returning.
 *  - return value: 0 = PIXIE_OK, negative = error (pixie_error_string()).  Nothing throws.
 *  - matrices are row-major fp32; `ld*` is the row pitch in elements.  The tensor-core path needs
 *    X 16-byte aligned and ldX % 4 == 0; anything else is routed to the (slow, exact) fp64 kernel.
 *  - labels are 1-indexed int32 exactly as pyFlowSOM returns them (0 = row holds NaN/Inf).
 */
#ifndef PIXIE_B200_H
#define PIXIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIXIE_OK 0
#define PIXIE_ERR_INVALID_ARG (-1)
#define PIXIE_ERR_WORKSPACE (-2)
#define PIXIE_ERR_CUDA (-3)
#define PIXIE_ERR_UNSUPPORTED (-4)
#define PIXIE_ERR_NO_DEVICE (-5)

/* rows per tile: the unit in which mini-batches of the batch SOM are interleaved */
#define PIXIE_TILE 128

/* flags for pixie_bmu_f32 / pixie_som_accum_f32 */
#define PIXIE_FLAG_AUTO 0u        /* tensor-core kernel when the shape allows, else exact kernel */
#define PIXIE_FLAG_FORCE_EXACT 1u /* fp64 brute-force kernel (bit-exact by construction; slow) */
#define PIXIE_FLAG_FORCE_TC 2u    /* fail with PIXIE_ERR_UNSUPPORTED instead of falling back */

/* slots of the optional device-side statistics array (uint64[PIXIE_NSTATS], accumulated) */
#define PIXIE_STAT_ROWS_FLAGGED 0  /* rows with >= 2 tensor-core candidates */
#define PIXIE_STAT_PAIRS 1         /* (row, node) pairs re-evaluated in fp32 */
#define PIXIE_STAT_ROWS_FP64 2     /* rows resolved by the fp64 replica of the reference loop */
#define PIXIE_STAT_ROWS_FIXUP 3    /* rows sent to the exact fix-up kernel (NaN/Inf/overflow) */
#define PIXIE_STAT_KERNEL 4        /* 1 = tensor-core kernel ran, 2 = exact kernel ran */
#define PIXIE_NSTATS 8

int pixie_version(void);
const char *pixie_error_string(int code);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
unsigned long long pixie_kernel_launches(void);
/* number of CUDA devices visible, or a negative error */
int pixie_device_count(void);
/* Diagnostic builds only (make prof, -DPIXIE_PROFILE): copies the kernel event trace to the host
 * (two uint64 per event: key, globaltimer ns) and resets it.  Returns the number of events;
 * always 0 in the production build. */
int pixie_debug_trace(unsigned long long *out_host, int max_events);

/* Which kernel and shared-memory layout the library would use for a (C, K) codebook: pure host
 * logic, no device needed (the CPU test-suite checks the planner with it).  train != 0: the plan of
 * the fused training kernel, else of assignment.  Writes up to 16 int32 to out:
 *   [0] 1 = a tensor-core plan exists   [1] kernel: 0 plain, 1 split-operand (3 x tf32)
 *   [2] SL  [3] slices per chunk  [4] accumulator chunks per tile  [5] epilogue groups
 *   [6] X pipeline stages  [7] bytes per stage  [8] dynamic shared memory requested
 *   [9] codebook image bytes  [10] TMEM columns  [11] tail8 layout  [12] sum tables in global memory
 *   [13] MMA N  [14] image rows per block  [15] K-steps
 * Returns PIXIE_OK or PIXIE_ERR_INVALID_ARG. */
int pixie_plan_describe(int32_t C, int32_t K, int32_t train, int32_t *out_host);

/* Bytes of device workspace pixie_bmu_f32 / pixie_som_accum_f32 need for these shapes. */
size_t pixie_workspace_bytes(int64_t n, int32_t C, int32_t K);

/*
 * BMU assignment.  Replaces pyFlowSOM.map_data_to_nodes(nodes, newdata)[0]
 * (cluster_helpers.py:152-157): labels[i] = 1 + argmin_k sqrt(sum_j (X[i,j] - W[k,j])^2), fp64
 * arithmetic on the fp32 inputs, first minimum wins, 0 for rows with NaN.
 *   X [n x C] ldX, W [K x C] contiguous, labels int32[n], all device memory.
 *   SN_or_null: optional double[K x (C+1)] -- per-node channel sums and count of the rows
 *   assigned to it ([S_k0..S_k,C-1, n_k]); overwritten.  This is the aggregate
 *   compute_pixel_cluster_channel_avg builds per FOV (pixel_cluster_utils.py:369-374).
 *   stats_or_null: optional uint64[PIXIE_NSTATS], accumulated (caller zeroes it).
 */
int pixie_bmu_f32(const float *X, int64_t n, int32_t C, int64_t ldX, const float *W, int32_t K,
                  int32_t *labels, double *SN_or_null, void *workspace, size_t ws_bytes,
                  uint32_t flags, unsigned long long *stats_or_null, void *stream);

/* Exact fp64 distance to the assigned node: dists[i] = sqrt(sum_j (X[i,j]-W[labels[i]-1,j])^2),
 * DBL_MAX when labels[i] == 0.  The second array map_data_to_nodes returns. */
int pixie_bmu_dist_f64(const float *X, int64_t n, int32_t C, int64_t ldX, const float *W, int32_t K,
                       const int32_t *labels, double *dists, void *stream);

/*
 * Per-node channel sums and counts for an EXISTING label array (labels 1..K; anything else is
 * skipped): SN = double[K x (C+1)], row k = [sum of X rows labelled k+1 | their count].  This is
 * the per-FOV aggregate of compute_pixel_cluster_channel_avg (pixel_cluster_utils.py:369-374:
 * groupby(cluster)[channels].sum() and .size()).  Deterministic (fixed summation order).
 */
int pixie_cluster_sums_f32(const float *X, int64_t n, int32_t C, int64_t ldX, const int32_t *labels,
                           int32_t K, double *SN, void *workspace, size_t ws_bytes, void *stream);

/*
 * N2 -- Feather column buffers -> device matrix.  `cols` holds C float64 columns of n values,
 * column c starting at cols + c * col_stride (the per-channel Arrow buffers of a FOV file,
 * pixie_preprocessing.py:172-183, uploaded as they are).  Writes
 *   X[i, c] = (float)(cols[c][i] / divisor[c])      (divisor NULL: no division)
 * i.e. normalize_data (cluster_helpers.py:244-246), the column gather (:151-156) and the
 * float64 -> device fp32 cast in one pass; bit-identical to rounding the reference's normalised
 * float64 table to fp32.  All pointers are device pointers; divisor is double[C].
 */
int pixie_columns_to_rows_f32(const double *cols, int64_t col_stride, int64_t n, int32_t C,
                              const double *divisor_or_null, float *X, int64_t ldX, void *stream);

/*
 * Parity mode: pyFlowSOM.som's own ONLINE rule (cluster_helpers.py:106-109; FlowSOM C_SOM as
 * restated in oracle/pixie_oracle.c) on the device, with the reference's floating-point operation
 * sequence -- the codebook it returns is bit-identical to that restatement run on the same fp32
 * inputs.  niter = rlen * n sequential single-sample updates: iteration k takes row
 * sample_idx[k] (device int64[niter], drawn by the caller -- pixie_libc_sample_indices reproduces
 * the reference's srand(seed)/rand() stream), moves every node within the shrinking Chebyshev
 * radius of its BMU towards it by alpha_k, and stops early between passes exactly as C_SOM does.
 * W64 [K x C] holds the initial codebook on entry and the trained one when the stream drains.
 * Sequential by nature: one CTA, does not shard (replicas only).  PIXIE_ERR_UNSUPPORTED when
 * K > 1024, C > 1024 or the fp64 codebook does not fit in shared memory.
 */
int pixie_som_online_f64(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64,
                         int32_t xdim, int32_t ydim, const int64_t *sample_idx, int64_t niter,
                         double alpha0, double alpha1, double radius0, double radius1,
                         long long *iters_done_or_null, void *stream);
/* HOST helper: out_host[k] = (int64)(n * (rand() / (RAND_MAX + 1.0))) after srand(seed) -- the
 * sample sequence of C_SOM.  Uses (and reseeds) the process-wide libc generator, as the reference
 * does. */
int pixie_libc_sample_indices(uint32_t seed, int64_t n, int64_t count, int64_t *out_host);

/*
 * One mini-batch step of the batch SOM (the B200 replacement for pyFlowSOM.som's inner loop,
 * cluster_helpers.py:106-109; algorithm in DESIGN.md section 4).  Visits tiles
 * tile_first, tile_first + tile_stride, ... (< ceil(n / PIXIE_TILE)), finds each row's BMU against
 * W32 and writes SN = double[K x (C+1)] per-node sums and counts.  In a multi-GPU run the caller
 * all-reduces SN (sum) across ranks before pixie_som_apply_f64.
 */
int pixie_som_accum_f32(const float *X, int64_t n, int32_t C, int64_t ldX, const float *W32,
                        int32_t K, int64_t tile_first, int64_t tile_stride, double *SN,
                        void *workspace, size_t ws_bytes, uint32_t flags,
                        unsigned long long *stats_or_null, void *stream);

/* Applies one batch update to the fp64 master codebook and refreshes its fp32 copy:
 *   H[k,b] = exp(-cheb(k,b)^2 / (2 sigma^2)); num_k = sum_b H[k,b] S_b; den_k = sum_b H[k,b] n_b;
 *   den_k > 0: W64_k += (1 - (1 - alpha)^den_k) (num_k / den_k - W64_k);  W32 = (float)W64. */
int pixie_som_apply_f64(double *W64, float *W32, const double *SN, int32_t xdim, int32_t ydim,
                        int32_t C, double sigma, double alpha, void *stream);

/*
 * Whole single-GPU training run: rlen passes x batches_per_pass steps of accum + apply, enqueued
 * back to back with no host synchronisation.  W64 [K x C] holds the initial codebook on entry and
 * the trained codebook when the stream drains; W32 [K x C] is scratch/out.  SN is double[K x (C+1)]
 * scratch.  Replaces pyFlowSOM.som for a device-resident fp32 matrix.
 */
int pixie_som_train_f32(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64, float *W32,
                        double *SN, int32_t xdim, int32_t ydim, int32_t rlen,
                        int32_t batches_per_pass, double alpha0, double alpha1, double radius0,
                        double radius1, void *workspace, size_t ws_bytes, uint32_t flags,
                        void *stream);

/*
 * Multi-GPU form of pixie_som_train_f32: one call per rank (one process per GPU), all ranks at the
 * same time.  X is this rank's row shard, whose first row is GLOBAL tile `tile_offset` (n == 0 is
 * allowed: a rank without rows still takes part in every exchange).  The per-step statistics are
 * summed across ranks INSIDE the training kernel over NVLink peer memory (no NCCL call, no host
 * synchronisation): every CTA pushes the values of its slice of the folded table into every rank's
 * exchange buffer -- value and step tag in one 16-byte store, so the receiver only polls its own
 * memory -- and adds the `world` values of each element in rank order.
 * `peer_bufs` is a HOST array of `world` device pointers, entry r being rank r's exchange buffer
 * mapped into this process (CUDA IPC / symmetric memory), each at least
 * pixie_peer_buffer_bytes(C, K) bytes, zero-initialised once.  `flag_base` must grow by at least
 * rlen * batches_per_pass between successive calls on the same buffers.  Every rank ends with the
 * bit-identical codebook.  The ranks must AGREE to call it: pixie_som_train_peers_supported()
 * tells a rank whether its shape / alignment fits (1) or not (0); a caller all-reduces (min) that
 * answer and otherwise runs pixie_som_accum_f32 + all-reduce + pixie_som_apply_f64 per step.
 * Returns PIXIE_ERR_UNSUPPORTED when called for a shape without a persistent plan.
 */
int pixie_som_train_peers_supported(int32_t C, int32_t K, int64_t ldX, int32_t x_aligned16);
size_t pixie_peer_buffer_bytes(int32_t C, int32_t K);
int pixie_som_train_peers_f32(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64,
                              float *W32, double *SN, int32_t xdim, int32_t ydim, int32_t rlen,
                              int32_t batches_per_pass, double alpha0, double alpha1,
                              double radius0, double radius1, int64_t tile_offset, int32_t world,
                              int32_t rank, const uint64_t *peer_bufs, uint32_t flag_base,
                              void *workspace, size_t ws_bytes, uint32_t flags, void *stream);

/*
 * N3 -- pixel preprocessing on the device: the arithmetic of
 * create_fov_pixel_data / preprocess_fov (/root/reference/src/ark/phenotyping/
 * pixie_preprocessing.py:18-80, :154-161) and normalize_rows (pixel_cluster_utils.py:109-142).
 *
 * img [H x W x C] fp32 (fp64 with PIXIE_PREPROCESS_IMG_F64; channel fastest, device) -> x = float64(img) / norm_vect[c]
 * -> per-channel 2-D gaussian filter (scipy.ndimage.gaussian_filter: separable, axis 0 then axis 1,
 * mode 'reflect'; `taps_host` is a HOST array of radius + 1 fp64 weights, taps_host[j] = weight at
 * distance j, computed by the caller exactly as scipy's _gaussian_kernel1d does; radius 0 = no
 * blur) -> `blurred` [H x W x C] fp64 (device, caller-owned)
 * -> keep pixels with sum_c x > pixel_thresh_val and any x != 0
 * -> kept pixels in image order, each divided by its channel sum:
 *      X64 [n_kept x C] fp64 and/or X32 [n_kept x C] fp32 at row pitch ldX32 (capacity H*W rows),
 *      row_index / column_index int32 [H*W], labels_out (seg_labels of the kept pixels) and the
 *      device scalar *n_kept.
 * fp64 throughout, in the reference's operation order: `blurred`, the kept set and X64 are
 * bit-identical to the scipy + pandas route.  PIXIE_PREPROCESS_BLUR_ONLY stops after the blur.
 * Asynchronous on `stream`; workspace of pixie_preprocess_workspace_bytes(H, W, C) bytes.
 */
#define PIXIE_PREPROCESS_BLUR_ONLY 1u
#define PIXIE_PREPROCESS_IMG_F64 2u /* img holds fp64 values (already normalised images) */
size_t pixie_preprocess_workspace_bytes(int32_t H, int32_t W, int32_t C);
int pixie_preprocess_fov_f64(const void *img, int32_t H, int32_t W, int32_t C,
                             const double *norm_vect_or_null, const double *taps_host,
                             int32_t radius, double pixel_thresh_val,
                             const int32_t *seg_labels_or_null, double *blurred, double *X64_or_null,
                             float *X32_or_null, int64_t ldX32, int32_t *row_index,
                             int32_t *column_index, int32_t *labels_out_or_null, int64_t *n_kept,
                             void *workspace, size_t ws_bytes, uint32_t flags, void *stream);

/*
 * N3, second half: the order statistics behind
 *     fov_full_pixel_data.replace(0, np.nan).quantile(q, axis=0)
 * (/root/reference/src/ark/phenotyping/pixie_preprocessing.py:405-410, :424-427), whose mean over
 * the FOVs is the normalisation row of the SOM.  For every column c of X [n x C] fp64 (row pitch
 * ldX): m[c] = number of valid entries (non-zero, non-NaN), lo[c] / hi[c] = the valid entries of
 * rank floor((m-1) q) and that rank + 1 (clipped to the last rank), NaN when m = 0 -- exact (radix
 * select, no sorting).  numpy's 'linear' quantile is lerp(lo, hi, (m-1) q - floor((m-1) q)); the
 * caller does that on the host (ark_analysis_b200.pixie_preprocessing.column_quantile).
 */
size_t pixie_column_quantile_workspace_bytes(int32_t C);
int pixie_column_quantile_f64(const double *X, int64_t n, int32_t C, int64_t ldX, double q,
                              double *lo, double *hi, int64_t *m, void *workspace,
                              size_t ws_bytes, void *stream);

/*
 * N4 -- consumers of the label array.
 *
 * pixie_label_histogram_i32: counts[s * n_clusters + c] += 1 for every pixel i with
 * s = seg_labels[i], c = clusters[i]: the per-cell histogram of pixel cluster labels that
 * create_c2pc_data builds with groupby(['label', cluster]).size() + pivot
 * (/root/reference/src/ark/phenotyping/cell_cluster_utils.py:119-132).  `counts` is
 * [n_seg x n_clusters] int32, ACCUMULATED into (zero it first; several FOV chunks may be added).
 * Pixels whose pair lies outside [0, n_seg) x [0, n_clusters) are skipped and counted in
 * *out_of_range_or_null.  seg_labels and clusters must be 16-byte aligned.
 *
 * pixie_scatter_labels_i16: img[row_index[i] * W + column_index[i]] = id_map[clusters[i]]
 * (or (int16) clusters[i] when id_map is null): the cluster mask of generate_pixel_cluster_mask
 * (/root/reference/src/ark/utils/data_utils.py:523-551; int16 "to allow for Photoshop loading").
 * `img` is [H x W] int16, written in place (zero it first).  With `winner_ws` (an [H x W] int32
 * scratch) duplicate coordinates resolve as numpy's fancy assignment does -- the LAST row wins --
 * in two passes; with null, coordinates must be unique (as they are in a pixel table).
 */
int pixie_label_histogram_i32(const int32_t *seg_labels, const int32_t *clusters, int64_t n,
                              int32_t n_seg, int32_t n_clusters, int32_t *counts,
                              unsigned long long *out_of_range_or_null, void *stream);
int pixie_scatter_labels_i16(const int32_t *row_index, const int32_t *column_index,
                             const int32_t *clusters, int64_t n, const int16_t *id_map_or_null,
                             int32_t map_len, int32_t H, int32_t W, int16_t *img,
                             int32_t *winner_ws_or_null, unsigned long long *out_of_range_or_null,
                             void *stream);

/*
 * Host-buffer entry point with pyFlowSOM.map_data_to_nodes' shape (what a ctypes/cgo/JNI binding
 * of the reference would call): nodes [K x C] and data [n x C] in HOST memory (fp32, row-major,
 * contiguous), labels int32[n] and optional dists double[n] in host memory.  Streams the rows
 * through the GPU in chunks (H2D, kernel, D2H overlapped on two streams) and returns when the
 * labels are on the host.  device < 0 = current device.
 */
int pixie_map_data_to_nodes_host_f32(const float *nodes, int32_t K, const float *data, int64_t n,
                                     int32_t C, int32_t *labels, double *dists_or_null,
                                     int32_t device, int64_t chunk_rows);

/* Same boundary with the reference's fp64 arrays (cluster_helpers.py:153-156 casts to float64):
 * values are rounded to fp32 on the host while staging. */
int pixie_map_data_to_nodes_host_f64(const double *nodes, int32_t K, const double *data, int64_t n,
                                     int32_t C, int32_t *labels, double *dists_or_null,
                                     int32_t device, int64_t chunk_rows);

#ifdef __cplusplus
}
#endif
#endif /* PIXIE_B200_H */
