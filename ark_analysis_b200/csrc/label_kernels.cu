// label_kernels.cu -- consumers of the BMU label array.
//
//   label_histogram_kernel : counts[cell, cluster] += 1 for every pixel -- the per-cell histogram
//       of pixel cluster labels that create_c2pc_data builds with groupby(['label', cluster]).size()
//       + pivot (/root/reference/src/ark/phenotyping/cell_cluster_utils.py:119-132).  Integer
//       counts: bit-exact whatever the order.
//   scatter_labels_kernel  : img[row_index * W + column_index] = id_map[cluster] (int16), the
//       cluster mask of generate_pixel_cluster_mask (/root/reference/src/ark/utils/data_utils.py:
//       523-551).
//
// Both are HBM-bound streaming kernels (8 resp. 12 bytes read per pixel, 128-bit coalesced
// loads).  The histogram aggregates in two levels before it touches global memory: equal keys of
// a thread's four consecutive pixels are merged in registers (pixels are in image order, so
// neighbours mostly share cell AND cluster), then equal keys inside the warp are merged with
// match.any and only the group leader issues the atomic (to an L2-resident table).
#include "common.cuh"

namespace pixie {

namespace {

__device__ __forceinline__ void warp_aggregated_add(int32_t *counts, long long key, int cnt)
{
    // key < 0: nothing to add for this lane (it still takes part in the vote)
    const unsigned peers = __match_any_sync(0xffffffffu, key);  // called by all 32 lanes
    if (key < 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    // sum of the group's counts: lanes of a group hold small numbers (<= 4), add them with shuffles
    // over the peer mask (at most 32 lanes: walk the set bits)
    int total = 0;
    for (unsigned m = peers; m; m &= m - 1) total += __shfl_sync(peers, cnt, __ffs(m) - 1);
    if (lane == leader) atomicAdd(counts + key, total);
}

__global__ void __launch_bounds__(256)
label_histogram_kernel(const int32_t *__restrict__ seg, const int32_t *__restrict__ clu, int64_t n,
                       int32_t n_seg, int32_t n_clu, int32_t *__restrict__ counts,
                       unsigned long long *__restrict__ bad)
{
    const int64_t nvec = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long my_bad = 0;
    // the loop bound is rounded up to a whole warp so that every lane reaches the warp votes
    const int64_t nvec_warp = (nvec + 31) / 32 * 32;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec_warp; v += stride) {
        long long key[4] = {-1, -1, -1, -1};
        int cnt[4] = {0, 0, 0, 0};
        if (v < nvec) {
            const int4 s = __ldcs(reinterpret_cast<const int4 *>(seg) + v);
            const int4 c = __ldcs(reinterpret_cast<const int4 *>(clu) + v);
            const int ss[4] = {s.x, s.y, s.z, s.w}, cc[4] = {c.x, c.y, c.z, c.w};
            int last = -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = ss[j] >= 0 && ss[j] < n_seg && cc[j] >= 0 && cc[j] < n_clu;
                if (!ok) {
                    ++my_bad;
                    continue;
                }
                const long long k = (long long)ss[j] * n_clu + cc[j];
                if (last >= 0 && key[last] == k) {
                    ++cnt[last];
                } else {
                    last = j;
                    key[j] = k;
                    cnt[j] = 1;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // skip the vote when no lane of the warp has a run starting at slot j
            if (__any_sync(0xffffffffu, key[j] >= 0)) warp_aggregated_add(counts, key[j], cnt[j]);
        }
    }
    // tail (n % 4 pixels)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (nvec << 2) + threadIdx.x;
        const int s = seg[i], c = clu[i];
        if (s >= 0 && s < n_seg && c >= 0 && c < n_clu)
            atomicAdd(counts + (long long)s * n_clu + c, 1);
        else
            ++my_bad;
    }
    if (bad != nullptr && my_bad) atomicAdd(bad, my_bad);
}

// pass 1 of the duplicate-safe scatter: winner[pixel] = highest row number that targets it
__global__ void __launch_bounds__(256)
scatter_winner_kernel(const int32_t *__restrict__ row_index, const int32_t *__restrict__ col_index,
                      int64_t n, int32_t H, int32_t W, int32_t *__restrict__ winner)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = __ldcs(row_index + i), c = __ldcs(col_index + i);
        if (r >= 0 && r < H && c >= 0 && c < W) atomicMax(winner + (int64_t)r * W + c, (int32_t)i);
    }
}

__global__ void __launch_bounds__(256)
scatter_labels_kernel(const int32_t *__restrict__ row_index, const int32_t *__restrict__ col_index,
                      const int32_t *__restrict__ clu, int64_t n, const int16_t *__restrict__ id_map,
                      int32_t map_len, int32_t H, int32_t W, int16_t *__restrict__ img,
                      const int32_t *__restrict__ winner, unsigned long long *__restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long my_bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = __ldcs(row_index + i), c = __ldcs(col_index + i), k = __ldcs(clu + i);
        const bool ok = r >= 0 && r < H && c >= 0 && c < W && (id_map == nullptr || (k >= 0 && k < map_len));
        if (!ok) {
            ++my_bad;
            continue;
        }
        if (winner != nullptr && winner[(int64_t)r * W + c] != (int32_t)i) continue;  // a later row wins
        img[(int64_t)r * W + c] = id_map ? id_map[k] : (int16_t)k;
    }
    if (bad != nullptr && my_bad) atomicAdd(bad, my_bad);
}

}  // namespace

cudaError_t launch_label_histogram(const int32_t *seg, const int32_t *clu, int64_t n, int32_t n_seg,
                                   int32_t n_clu, int32_t *counts, unsigned long long *bad,
                                   int num_sms, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)num_sms * 8) blocks = (int64_t)num_sms * 8;
    label_histogram_kernel<<<(int)blocks, 256, 0, stream>>>(seg, clu, n, n_seg, n_clu, counts, bad);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_scatter_labels(const int32_t *row_index, const int32_t *col_index,
                                  const int32_t *clu, int64_t n, const int16_t *id_map,
                                  int32_t map_len, int32_t H, int32_t W, int16_t *img,
                                  int32_t *winner, unsigned long long *bad, int num_sms,
                                  cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)num_sms * 8) blocks = (int64_t)num_sms * 8;
    if (winner != nullptr) {
        cudaError_t e = cudaMemsetAsync(winner, 0xFF, sizeof(int32_t) * (size_t)H * W, stream);  // -1
        if (e != cudaSuccess) return e;
        scatter_winner_kernel<<<(int)blocks, 256, 0, stream>>>(row_index, col_index, n, H, W, winner);
        count_launch();
    }
    scatter_labels_kernel<<<(int)blocks, 256, 0, stream>>>(row_index, col_index, clu, n, id_map,
                                                           map_len, H, W, img, winner, bad);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
