// som_kernels.cu -- the CUDA-core kernels of the Pixie SOM path:
//   bmu_exact_kernel     fp64 replica of the reference BMU loop (fix-up of sentinel rows, forced
//                        exact mode, and shapes the tensor-core kernel does not take)
//   bmu_dist_kernel      exact fp64 distance to the assigned node (map_data_to_nodes()[1])
//   cluster_sums_kernel  per-node channel sums + counts for a label array (batch-SOM statistics
//                        and the SOM-cluster channel averages of pixel_cluster_utils.py:369-404)
//   reduce_partials / som_apply  the batch-SOM update
#include <float.h>

#include "common.cuh"
#include "som_update.cuh"

namespace pixie {

// ------------------------------------------------------------------------------------------------
// exact BMU: identical operation sequence to oracle/pixie_oracle.c nearest_node()
// (cluster_helpers.py:152-157 semantics): fp64 subtract, multiply, add in channel order (no FMA
// contraction), sqrt, strict '<', nodes in index order.  A warp takes 32 consecutive rows; every
// row that needs work is then handled by the WHOLE warp -- lane l evaluates nodes l, l+32, ... and
// keeps its first minimum; the lexicographic (distance, index) minimum over lanes is exactly the
// first minimum of the sequential loop.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bmu_exact_kernel(const float *__restrict__ X, int64_t n, int C, int64_t ldX,
                 const float *__restrict__ W, int K, int32_t *__restrict__ labels,
                 int64_t tile_first, int64_t tile_stride, int64_t ntiles, int compact_labels,
                 const int *__restrict__ fixup_count, double *__restrict__ SN_add)
{
    const bool fix_only = fixup_count != nullptr;
    if (fix_only && *fixup_count == 0) return;  // nothing was flagged: the common case
    const int lane = threadIdx.x & 31;
    for (int64_t j = blockIdx.x; j < ntiles; j += gridDim.x) {
        const int64_t tile = tile_first + j * tile_stride;
        const int64_t row = tile * kTile + threadIdx.x;
        const int64_t lidx = compact_labels ? j * kTile + threadIdx.x : row;
        bool todo = row < n;
        if (todo && fix_only) todo = labels[lidx] == kLabelFixup;
        unsigned work = __ballot_sync(0xffffffffu, todo);
        while (work) {
            const int src = __ffs(work) - 1;
            work &= work - 1;
            const int64_t r = row - lane + src;
            const float *x = X + (size_t)r * ldX;
            int minid = -1;
            double mindist = DBL_MAX;
            for (int k = lane; k < K; k += 32) {
                const float *w = W + (size_t)k * C;
                double acc = 0.0;
                for (int c = 0; c < C; ++c) {
                    const double tmp = __dsub_rn((double)__ldg(x + c), (double)__ldg(w + c));
                    acc = __dadd_rn(acc, __dmul_rn(tmp, tmp));
                }
                const double d = __dsqrt_rn(acc);
                if (d < mindist) {
                    mindist = d;
                    minid = k;
                }
            }
            // lexicographic (distance, index) minimum; lanes without a finite candidate carry
            // (DBL_MAX, INT_MAX) and never win against a real one
            int bid = minid < 0 ? 0x7fffffff : minid;
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, mindist, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bid, o);
                if (od < mindist || (od == mindist && oi < bid)) {
                    mindist = od;
                    bid = oi;
                }
            }
            if (lane == src) labels[lidx] = (bid == 0x7fffffff) ? 0 : bid + 1;
            if (SN_add != nullptr && bid != 0x7fffffff) {
                // train mode: the tensor-core kernel left this row out of the fused sums
                double *dst = SN_add + (size_t)bid * (C + 1);
                for (int c = lane; c < C; c += 32) atomicAdd(dst + c, (double)__ldg(x + c));
                if (lane == 0) atomicAdd(dst + C, 1.0);
            }
        }
    }
}

cudaError_t launch_bmu_exact(const float *X, int64_t n, int C, int64_t ldX, const float *W, int K,
                             int32_t *labels, int64_t tile_first, int64_t tile_stride,
                             int64_t ntiles, int compact_labels, const int *fixup_count_or_null,
                             double *SN_add_or_null, cudaStream_t stream)
{
    if (ntiles <= 0) return cudaSuccess;
    int64_t grid = ntiles < 148 * 16 ? ntiles : 148 * 16;
    bmu_exact_kernel<<<(unsigned)grid, 128, 0, stream>>>(X, n, C, ldX, W, K, labels, tile_first,
                                                        tile_stride, ntiles, compact_labels,
                                                        fixup_count_or_null, SN_add_or_null);
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
// cluster sums: SN[k] = [sum of the rows labelled k+1 | their count].
//
// Each CTA walks its tiles.  Per tile the 128 rows are staged in shared memory (coalesced), the
// labels are counting-sorted with STABLE ranks (warp match + per-warp counts, so the order inside a
// node's segment is the row order whatever the scheduling), then each warp sums whole segments with
// lanes = channels and adds the segment sum into the CTA's fp32 accumulator -- in shared memory when
// K x (C+1) fits (a segment is owned by one warp and tiles are separated by a CTA barrier: no
// atomics, fixed order), else in the CTA's private slice of a global scratch.  The CTA's totals go
// to partials[cta]; the LAST CTA to finish (atomic ticket) folds all partials into SN in fp64 in
// CTA order, so the result is run-to-run deterministic and no second launch is needed.
// partials layout: [nparts][K][C+1] fp32 (last column = count).
// ------------------------------------------------------------------------------------------------
constexpr int kSumThreads = 256;
constexpr int kSumWarps = kSumThreads / 32;

template <bool kSmemAcc>
__global__ void __launch_bounds__(kSumThreads)
cluster_sums_kernel(const float *__restrict__ X, int64_t n, int C, int64_t ldX,
                    const int32_t *__restrict__ labels, int compact_labels, int K,
                    int64_t tile_first, int64_t tile_stride, int64_t ntiles,
                    float *__restrict__ partials, double *__restrict__ SN, unsigned int *sync)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int len = K * (C + 1);
    // layout: [tile rows 128 x Cp fp32][acc (optional) len fp32][ints]
    const int Cp = (C + 3) & ~3;
    float *s_tile = reinterpret_cast<float *>(s_raw);
    float *s_acc = s_tile + kTile * Cp;
    int *s_int = reinterpret_cast<int *>(s_acc + (kSmemAcc ? len : 0));
    int *s_wcnt = s_int;                 // [4][K] rows per (row-warp, node) in this tile
    int *s_start = s_wcnt + 4 * K;       // [K+1] segment starts
    int *s_lab = s_start + K + 1;        // [128] label-1 per row (-1 = skip)
    int *s_rank = s_lab + kTile;         // [128] stable rank of the row within its node
    int *s_perm = s_rank + kTile;        // [128] rows sorted by node
    int *s_seg = s_perm + kTile;         // [128] list of non-empty nodes
    __shared__ int s_nseg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *mine = partials + (size_t)blockIdx.x * len;
    float *acc = kSmemAcc ? s_acc : mine;
    for (int i = tid; i < len; i += kSumThreads) acc[i] = 0.f;
    __syncthreads();

    for (int64_t j = blockIdx.x; j < ntiles; j += gridDim.x) {
        const int64_t tile = tile_first + j * tile_stride;
        const int64_t row0 = tile * kTile;
        // stage the tile (rows past n are never referenced)
        {
            const int vec_per_row = Cp >> 2;
            const bool vec_ok = ((ldX & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
            if (vec_ok) {
                for (int i = tid; i < kTile * vec_per_row; i += kSumThreads) {
                    const int r = i / vec_per_row, v = i - r * vec_per_row;
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row0 + r < n) {
                        const float *src = X + (size_t)(row0 + r) * ldX + 4 * v;
                        if (4 * v + 3 < C) {
                            val = __ldg(reinterpret_cast<const float4 *>(src));
                        } else {
                            if (4 * v + 0 < C) val.x = __ldg(src + 0);
                            if (4 * v + 1 < C) val.y = __ldg(src + 1);
                            if (4 * v + 2 < C) val.z = __ldg(src + 2);
                        }
                    }
                    reinterpret_cast<float4 *>(s_tile)[i] = val;
                }
            } else {
                for (int i = tid; i < kTile * Cp; i += kSumThreads) {
                    const int r = i / Cp, c = i - r * Cp;
                    s_tile[i] = (row0 + r < n && c < C) ? __ldg(X + (size_t)(row0 + r) * ldX + c) : 0.f;
                }
            }
        }
        for (int i = tid; i < 4 * K; i += kSumThreads) s_wcnt[i] = 0;
        __syncthreads();
        if (tid < kTile) {
            const int64_t row = row0 + tid;
            int lab = -1;
            if (row < n) {
                lab = labels[compact_labels ? j * kTile + tid : row] - 1;
                if (lab < 0 || lab >= K) lab = -1;
            }
            s_lab[tid] = lab;
            // stable rank inside the warp: earlier lanes with the same label
            const unsigned peers = __match_any_sync(0xffffffffu, lab);
            const int rank_w = __popc(peers & ((1u << lane) - 1u));
            s_rank[tid] = rank_w;
            if (lab >= 0 && rank_w == 0) s_wcnt[warp * K + lab] = __popc(peers);
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan over nodes of the per-node totals, list of non-empty nodes
            int carry = 0, nseg = 0;
            for (int base = 0; base < K; base += 32) {
                const int k = base + lane;
                int c = 0;
                if (k < K) c = s_wcnt[k] + s_wcnt[K + k] + s_wcnt[2 * K + k] + s_wcnt[3 * K + k];
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
        if (tid < kTile) {
            const int lab = s_lab[tid];
            if (lab >= 0) {
                int before = 0;  // rows of the same node in earlier row-warps
                for (int w = 0; w < warp; ++w) before += s_wcnt[w * K + lab];
                s_perm[s_start[lab] + before + s_rank[tid]] = tid;
            }
        }
        __syncthreads();
        const int nseg = s_nseg;
        for (int sg = warp; sg < nseg; sg += kSumWarps) {
            const int k = s_seg[sg];
            const int lo = s_start[k], hi = s_start[k + 1];
            float *dst = acc + (size_t)k * (C + 1);
            for (int c0 = 0; c0 < C; c0 += 32) {
                const int c = c0 + lane;
                if (c < C) {
                    float a = 0.f;
                    for (int q = lo; q < hi; ++q) a += s_tile[s_perm[q] * Cp + c];
                    dst[c] += a;
                }
            }
            if (lane == 0) dst[C] += (float)(hi - lo);
        }
        __syncthreads();
    }
    if (kSmemAcc)
        for (int i = tid; i < len; i += kSumThreads) mine[i] = s_acc[i];
    // ---- grid barrier (all CTAs are co-resident: grid <= SM count, one CTA per SM fits), then
    // every CTA folds its slice of the K x (C+1) table over all partials in CTA order, in fp64:
    // deterministic, parallel, and no second launch.
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(&sync[0], 1u);
        unsigned spins = 0;
        while (atomicAdd(&sync[0], 0u) < gridDim.x) {
            __nanosleep(64);
            if (++spins > (1u << 24)) __trap();  // ~2 s: never on a healthy launch
        }
    }
    __syncthreads();
    __threadfence();
    {
        const int nparts = gridDim.x;
        const int per = (len + nparts - 1) / nparts;
        const int e0 = blockIdx.x * per;
        const int e1 = min(len, e0 + per);
        // 8 threads per element: thread q of the octet sums partials q, q+8, ... (ascending), the
        // octet is then combined in a fixed shuffle tree
        const int oct = tid >> 3, q = tid & 7;
        const int rounds = (per + kSumThreads / 8 - 1) / (kSumThreads / 8);  // uniform trip count
        for (int it = 0; it < rounds; ++it) {
            const int e = e0 + it * (kSumThreads / 8) + oct;
            // all loads of this thread first (independent, in flight together), then the sum in
            // ascending CTA order
            float v[kFoldMax];
#pragma unroll
            for (int u = 0; u < kFoldMax; ++u) {
                const int p = q + 8 * u;
                v[u] = (e < e1 && p < nparts) ? __ldcg(partials + (size_t)p * len + e) : 0.f;
            }
            double a = 0.0;
#pragma unroll
            for (int u = 0; u < kFoldMax; ++u) a += (double)v[u];
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            if (e < e1 && q == 0) SN[e] = a;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // the last CTA to leave re-arms the barrier for the next launch on this stream
        if (atomicAdd(&sync[1], 1u) == gridDim.x - 1) {
            sync[0] = 0u;
            sync[1] = 0u;
            __threadfence();
        }
    }
}

cudaError_t launch_cluster_sums(const float *X, int64_t n, int C, int64_t ldX,
                                const int32_t *labels, int compact_labels, int K,
                                int64_t tile_first, int64_t tile_stride, int64_t ntiles,
                                float *partials, int nparts, double *SN, unsigned int *ticket,
                                cudaStream_t stream)
{
    const int len = K * (C + 1);
    const int Cp = (C + 3) & ~3;
    const size_t ints = (size_t)(4 * K + K + 1 + 4 * kTile) * sizeof(int);
    const size_t tile_bytes = (size_t)kTile * Cp * sizeof(float);
    const size_t smem_acc = tile_bytes + (size_t)len * sizeof(float) + ints;
    const size_t smem_noacc = tile_bytes + ints;
    int grid = nparts;
    if ((int64_t)grid > ntiles) grid = ntiles > 0 ? (int)ntiles : 1;
    cudaError_t e;
    if (smem_acc <= 200 * 1024) {
        e = cudaFuncSetAttribute(cluster_sums_kernel<true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_acc);
        if (e != cudaSuccess) return e;
        void *args[] = {&X, &n, &C, &ldX, &labels, &compact_labels, &K, &tile_first, &tile_stride,
                        &ntiles, &partials, &SN, &ticket};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(&cluster_sums_kernel<true>),
                                        dim3(grid), dim3(kSumThreads), args, smem_acc, stream);
        count_launch();
        return e;
    } else {
        e = cudaFuncSetAttribute(cluster_sums_kernel<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_noacc);
        if (e != cudaSuccess) return e;
        void *args[] = {&X, &n, &C, &ldX, &labels, &compact_labels, &K, &tile_first, &tile_stride,
                        &ntiles, &partials, &SN, &ticket};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(&cluster_sums_kernel<false>),
                                        dim3(grid), dim3(kSumThreads), args, smem_noacc, stream);
        count_launch();
        return e;
    }
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// batch-SOM update (fp64 restatement in oracle/pixie_oracle.c
// oracle_som_batch): som_update_nodes (som_update.cuh), the function the whole-pass training kernel
// runs at the end of every step, on its own grid -- CTA j updates nodes j, j + grid, ...
// ------------------------------------------------------------------------------------------------
constexpr int kApplyThreads = 256;

__global__ void __launch_bounds__(kApplyThreads)
som_apply_kernel(double *__restrict__ W64, float *__restrict__ W32, const double *__restrict__ SN,
                 int xdim, int ydim, int C, double inv2s2, double alpha)
{
    extern __shared__ double s_dyn[];
    som_update_nodes(
        SN, W64, W32, xdim * ydim, C, ydim, inv2s2, alpha, (int)blockIdx.x, (int)gridDim.x,
        (int)threadIdx.x, kApplyThreads, s_dyn, [] { __syncthreads(); },
        [](int, int, float) {}, [](int, int, double, bool) {});
}

cudaError_t launch_som_apply(double *W64, float *W32, const double *SN, int xdim, int ydim, int C,
                             double sigma, double alpha, cudaStream_t stream)
{
    const double inv2s2 = 1.0 / (2.0 * sigma * sigma);
    const int K = xdim * ydim;
    const size_t smem = som_update_scratch_bytes(C, K);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(som_apply_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    som_apply_kernel<<<K < 148 ? K : 148, kApplyThreads, smem, stream>>>(W64, W32, SN, xdim, ydim, C,
                                                                         inv2s2, alpha);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
