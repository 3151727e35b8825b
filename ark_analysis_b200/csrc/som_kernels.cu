// som_kernels.cu -- the CUDA-core kernels of the Pixie SOM path:
//   bmu_exact_kernel     fp64 replica of the reference BMU loop (fix-up of sentinel rows, forced
//                        exact mode, and shapes the tensor-core kernel does not take)
//   bmu_dist_kernel      exact fp64 distance to the assigned node (map_data_to_nodes()[1])
//   cluster_sums_kernel  per-node channel sums + counts for a label array (batch-SOM statistics
//                        and the SOM-cluster channel averages of pixel_cluster_utils.py:369-404)
//   reduce_partials / som_apply  the batch-SOM update
#include <float.h>

#include "common.cuh"

namespace pixie {

// ------------------------------------------------------------------------------------------------
// exact BMU: one thread per row; identical operation sequence to oracle/pixie_oracle.c
// nearest_node() (cluster_helpers.py:152-157 semantics): fp64 subtract, multiply, add in channel
// order (no FMA contraction), sqrt, strict '<', nodes in index order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bmu_exact_kernel(const float *__restrict__ X, int64_t n, int C, int64_t ldX,
                 const float *__restrict__ W, int K, int32_t *__restrict__ labels,
                 int64_t tile_first, int64_t tile_stride, int64_t ntiles, int compact_labels,
                 const int *__restrict__ fixup_count)
{
    const bool fix_only = fixup_count != nullptr;
    if (fix_only && *fixup_count == 0) return;  // nothing was flagged: the common case
    for (int64_t j = blockIdx.x; j < ntiles; j += gridDim.x) {
        const int64_t tile = tile_first + j * tile_stride;
        const int64_t row = tile * kTile + threadIdx.x;
        const int64_t lidx = compact_labels ? j * kTile + threadIdx.x : row;
        if (row >= n) continue;
        if (fix_only && labels[lidx] != kLabelFixup) continue;
        const float *x = X + (size_t)row * ldX;
        int minid = -1;
        double mindist = DBL_MAX;
        for (int k = 0; k < K; ++k) {
            const float *w = W + (size_t)k * C;
            double acc = 0.0;
            for (int c = 0; c < C; ++c) {
                const double tmp = __dsub_rn((double)x[c], (double)__ldg(w + c));
                acc = __dadd_rn(acc, __dmul_rn(tmp, tmp));
            }
            const double d = __dsqrt_rn(acc);
            if (d < mindist) {
                mindist = d;
                minid = k;
            }
        }
        labels[lidx] = minid + 1;
    }
}

cudaError_t launch_bmu_exact(const float *X, int64_t n, int C, int64_t ldX, const float *W, int K,
                             int32_t *labels, int64_t tile_first, int64_t tile_stride,
                             int64_t ntiles, int compact_labels, const int *fixup_count_or_null,
                             cudaStream_t stream)
{
    if (ntiles <= 0) return cudaSuccess;
    int64_t grid = ntiles < 148 * 16 ? ntiles : 148 * 16;
    bmu_exact_kernel<<<(unsigned)grid, 128, 0, stream>>>(X, n, C, ldX, W, K, labels, tile_first,
                                                        tile_stride, ntiles, compact_labels,
                                                        fixup_count_or_null);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__global__ void bmu_dist_kernel(const float *__restrict__ X, int64_t n, int C, int64_t ldX,
                                const float *__restrict__ W, int K,
                                const int32_t *__restrict__ labels, double *__restrict__ dists)
{
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n;
         row += (int64_t)gridDim.x * blockDim.x) {
        const int k = labels[row] - 1;
        double d = DBL_MAX;
        if (k >= 0 && k < K) {
            const float *x = X + (size_t)row * ldX;
            const float *w = W + (size_t)k * C;
            double acc = 0.0;
            for (int c = 0; c < C; ++c) {
                const double tmp = __dsub_rn((double)x[c], (double)__ldg(w + c));
                acc = __dadd_rn(acc, __dmul_rn(tmp, tmp));
            }
            d = __dsqrt_rn(acc);
        }
        dists[row] = d;
    }
}

cudaError_t launch_bmu_dist(const float *X, int64_t n, int C, int64_t ldX, const float *W, int K,
                            const int32_t *labels, double *dists, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    bmu_dist_kernel<<<(unsigned)blocks, 256, 0, stream>>>(X, n, C, ldX, W, K, labels, dists);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// cluster sums.  Each CTA walks its tiles; per tile the 128 rows are counting-sorted by label in
// shared memory, then each warp sums whole label segments (lanes = channels, fixed row order) and
// adds the segment sum into the CTA's private fp32 partial buffer in global memory (L2-resident,
// plain read-modify-write: a node's segment of one tile is owned by exactly one warp, and tiles are
// separated by a CTA barrier, so no atomics and a fixed summation order).  A second kernel folds
// the per-CTA partials into fp64 in fixed order => results are run-to-run deterministic.
// partials layout: [nparts][K][C+1] fp32 (last column = count).
// ------------------------------------------------------------------------------------------------
constexpr int kSumThreads = 256;

__global__ void __launch_bounds__(kSumThreads)
cluster_sums_kernel(const float *__restrict__ X, int64_t n, int C, int64_t ldX,
                    const int32_t *__restrict__ labels, int compact_labels, int K,
                    int64_t tile_first, int64_t tile_stride, int64_t ntiles,
                    float *__restrict__ partials)
{
    extern __shared__ int s_int[];
    int *s_cnt = s_int;            // [K]   rows per node in this tile
    int *s_start = s_cnt + K;      // [K+1] segment starts
    int *s_lab = s_start + K + 1;  // [128] label-1 per row (-1 = skip)
    int *s_rank = s_lab + kTile;   // [128] rank of the row within its node
    int *s_perm = s_rank + kTile;  // [128] rows sorted by node
    int *s_seg = s_perm + kTile;   // [128] list of non-empty nodes
    __shared__ int s_nseg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = kSumThreads / 32;
    float *mine = partials + (size_t)blockIdx.x * K * (C + 1);
    for (int i = tid; i < K * (C + 1); i += kSumThreads) mine[i] = 0.f;
    __syncthreads();

    for (int64_t j = blockIdx.x; j < ntiles; j += gridDim.x) {
        const int64_t tile = tile_first + j * tile_stride;
        const int64_t row0 = tile * kTile;
        for (int i = tid; i < K; i += kSumThreads) s_cnt[i] = 0;
        __syncthreads();
        if (tid < kTile) {
            const int64_t row = row0 + tid;
            int lab = -1;
            if (row < n) {
                lab = labels[compact_labels ? j * kTile + tid : row] - 1;
                if (lab < 0 || lab >= K) lab = -1;
            }
            s_lab[tid] = lab;
        }
        __syncthreads();
        // ranks must not depend on thread scheduling: row order within a node is the row index.
        if (tid < kTile) {
            const int lab = s_lab[tid];
            int rank = 0;
            if (lab >= 0) {
                for (int t = 0; t < tid; ++t) rank += (s_lab[t] == lab);
                atomicAdd(&s_cnt[lab], 1);
            }
            s_rank[tid] = rank;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of s_cnt and list of non-empty nodes, in node order
            int carry = 0, nseg = 0;
            for (int base = 0; base < K; base += 32) {
                const int k = base + lane;
                const int c = k < K ? s_cnt[k] : 0;
                int incl = c;
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(~0u, incl, o);
                    if (lane >= o) incl += t;
                }
                if (k < K) s_start[k] = carry + incl - c;
                const unsigned nz = __ballot_sync(~0u, c > 0);
                if (c > 0) s_seg[nseg + __popc(nz & ((1u << lane) - 1u))] = k;
                nseg += __popc(nz);
                carry += __shfl_sync(~0u, incl, 31);
            }
            if (lane == 0) {
                s_start[K] = carry;
                s_nseg = nseg;
            }
        }
        __syncthreads();
        if (tid < kTile && s_lab[tid] >= 0) s_perm[s_start[s_lab[tid]] + s_rank[tid]] = tid;
        __syncthreads();
        const int nseg = s_nseg;
        for (int sg = warp; sg < nseg; sg += nwarps) {
            const int k = s_seg[sg];
            const int lo = s_start[k], hi = lo + s_cnt[k];
            float *dst = mine + (size_t)k * (C + 1);
            for (int c0 = 0; c0 < C; c0 += 32) {
                const int c = c0 + lane;
                if (c < C) {
                    float acc = 0.f;
                    for (int q = lo; q < hi; ++q)
                        acc += __ldg(X + (size_t)(row0 + s_perm[q]) * ldX + c);
                    dst[c] += acc;
                }
            }
            if (lane == 0) dst[C] += (float)(hi - lo);
        }
        __syncthreads();
    }
}

// SN[i] = sum over parts (fixed order, fp64) of partials[part][i]
__global__ void reduce_partials_kernel(const float *__restrict__ partials, int nparts, int len,
                                       double *__restrict__ SN)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += (double)partials[(size_t)p * len + i];
    SN[i] = acc;
}

cudaError_t launch_cluster_sums(const float *X, int64_t n, int C, int64_t ldX,
                                const int32_t *labels, int compact_labels, int K,
                                int64_t tile_first, int64_t tile_stride, int64_t ntiles,
                                float *partials, int nparts, double *SN, cudaStream_t stream)
{
    const int len = K * (C + 1);
    const size_t smem = (size_t)(K + K + 1 + 4 * kTile) * sizeof(int);
    // every partial buffer is (re)initialised by its CTA, so always launch all nparts CTAs
    cluster_sums_kernel<<<nparts, kSumThreads, smem, stream>>>(X, n, C, ldX, labels,
                                                              compact_labels, K, tile_first,
                                                              tile_stride, ntiles, partials);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    reduce_partials_kernel<<<(len + 255) / 256, 256, 0, stream>>>(partials, nparts, len, SN);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// batch-SOM update (fp64 restatement in oracle/pixie_oracle.c
// oracle_som_batch).  One CTA per node k, threads over channels.
// ------------------------------------------------------------------------------------------------
__global__ void som_apply_kernel(double *__restrict__ W64, float *__restrict__ W32,
                                 const double *__restrict__ SN, int xdim, int ydim, int C,
                                 double inv2s2, double alpha)
{
    const int K = xdim * ydim;
    const int k = blockIdx.x;
    const int kx = k / ydim, ky = k % ydim;
    __shared__ double s_den;
    if (threadIdx.x == 0) {
        double den = 0.0;
        for (int b = 0; b < K; ++b) {
            const double cnt = SN[(size_t)b * (C + 1) + C];
            if (cnt == 0.0) continue;
            const int dx = abs(kx - b / ydim), dy = abs(ky - b % ydim);
            const double d = (double)(dx > dy ? dx : dy);
            den += exp(-d * d * inv2s2) * cnt;
        }
        s_den = den;
    }
    __syncthreads();
    const double den = s_den;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double w = W64[(size_t)k * C + c];
        if (den > 0.0) {
            double num = 0.0;
            for (int b = 0; b < K; ++b) {
                const double cnt = SN[(size_t)b * (C + 1) + C];
                if (cnt == 0.0) continue;
                const int dx = abs(kx - b / ydim), dy = abs(ky - b % ydim);
                const double d = (double)(dx > dy ? dx : dy);
                num += exp(-d * d * inv2s2) * SN[(size_t)b * (C + 1) + c];
            }
            const double beta = 1.0 - pow(1.0 - alpha, den);
            w += beta * (num / den - w);
            W64[(size_t)k * C + c] = w;
        }
        W32[(size_t)k * C + c] = (float)w;
    }
}

cudaError_t launch_som_apply(double *W64, float *W32, const double *SN, int xdim, int ydim, int C,
                             double sigma, double alpha, cudaStream_t stream)
{
    const double inv2s2 = 1.0 / (2.0 * sigma * sigma);
    som_apply_kernel<<<xdim * ydim, 128, 0, stream>>>(W64, W32, SN, xdim, ydim, C, inv2s2, alpha);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
