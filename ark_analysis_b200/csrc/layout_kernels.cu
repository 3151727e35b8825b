// layout_kernels.cu -- the data formats either side of the SOM path.
//
// Pixie keeps every FOV as a Feather (Arrow IPC, uncompressed) file with one float64 column per
// channel (written at /root/reference/src/ark/phenotyping/pixie_preprocessing.py:172-183, read at
// pixel_som_clustering.py:118).  The reference turns that into the SOM input with three host
// passes per FOV: DataFrame.copy() + div by the normalisation row (cluster_helpers.py:244-246),
// the .loc[...] column gather (:151-156) and .astype(float64) (:153,:156) -- and pyFlowSOM then
// transposes to Fortran order.  Here the column buffers go to the device as they are and ONE
// kernel does normalise + cast + transpose:
//
//   X[i, c] = (float)(cols[c][i] / divisor[c])       fp64 IEEE division, then round to fp32
//
// which is bit-identical to casting the reference's normalised float64 table to fp32.
//
// HBM-bound: 8 C bytes read + 4 C bytes written per pixel.  A CTA takes 128-row tiles: its warps
// read the columns (a warp-wide load is 256 contiguous bytes of one column), stage the tile
// transposed in shared memory (row pitch odd -> conflict-free both ways) and write the rows back
// contiguously.
#include "common.cuh"

namespace pixie {

namespace {
constexpr int kColThreads = 256;

__global__ void __launch_bounds__(kColThreads)
columns_to_rows_kernel(const double *__restrict__ cols, int64_t col_stride, int64_t n, int C,
                       const double *__restrict__ divisor, float *__restrict__ X, int64_t ldX,
                       int tile_rows, int pitch)
{
    extern __shared__ float tile[];  // [tile_rows][pitch]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = kColThreads / 32;
    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    // element cursor of the write-back loop: thread t owns flat elements t, t + 256, ... of the
    // tile_rows x C tile; (r, c) advance by a constant step with carry (no division per element)
    const int dr = kColThreads / C, dc = kColThreads % C;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t row0 = t * tile_rows;
        const int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
        // ---- columns -> shared memory (transposed), normalise + cast on the way
        for (int c = warp; c < C; c += nwarp) {
            const double *src = cols + (size_t)c * (size_t)col_stride + row0;
            const double d = divisor ? divisor[c] : 1.0;
            for (int r = lane; r < tile_rows; r += 128) {
                // four independent loads in flight per thread
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rr = r + 32 * u;
                    v[u] = (rr < rows) ? __ldg(src + rr) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rr = r + 32 * u;
                    if (rr < tile_rows)
                        tile[rr * pitch + c] = (float)(divisor ? __ddiv_rn(v[u], d) : v[u]);
                }
            }
        }
        __syncthreads();
        // ---- shared memory -> rows
        {
            int r = threadIdx.x / C, c = threadIdx.x % C;
            float *dst = X + (size_t)row0 * (size_t)ldX;
            while (r < rows) {
                dst[(size_t)r * (size_t)ldX + c] = tile[r * pitch + c];
                c += dc;
                r += dr;
                if (c >= C) {
                    c -= C;
                    ++r;
                }
            }
        }
        __syncthreads();
    }
}
}  // namespace

cudaError_t launch_columns_to_rows(const double *cols, int64_t col_stride, int64_t n, int C,
                                   const double *divisor, float *X, int64_t ldX, int num_sms,
                                   cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int pitch = C | 1;
    int tile_rows = 128;
    while (tile_rows > 32 && (size_t)tile_rows * pitch * sizeof(float) > 96u * 1024u) tile_rows -= 32;
    const size_t smem = (size_t)tile_rows * pitch * sizeof(float);
    if (smem > 96u * 1024u) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(columns_to_rows_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    // enough resident CTAs to cover the memory latency; a multiple of the SM count
    int per_sm = (int)(200u * 1024u / (smem + 1024u));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms * per_sm;
    if (grid > ntiles) grid = ntiles;
    columns_to_rows_kernel<<<(unsigned)grid, kColThreads, smem, stream>>>(cols, col_stride, n, C,
                                                                          divisor, X, ldX,
                                                                          tile_rows, pitch);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
