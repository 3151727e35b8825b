// preprocess_kernels.cu -- Pixie pixel preprocessing on the device:
// the arithmetic of create_fov_pixel_data / preprocess_fov
// (/root/reference/src/ark/phenotyping/pixie_preprocessing.py:18-80, :154-161) and normalize_rows
// (/root/reference/src/ark/phenotyping/pixel_cluster_utils.py:109-142), so that the SOM input
// matrix is born in HBM instead of travelling through a DataFrame and a Feather file:
//
//   x      = float64(img[h, w, c]) / norm_vect[c]                    (pixie_preprocessing.py:154-161)
//   x      = gaussian_filter(x[:, :, c], sigma) for every channel    (:47-49; scipy.ndimage)
//   keep   = sum_c x > pixel_thresh_val  and  any_c x != 0           (:67-72)
//   X[j,:] = x / sum_c x     for the kept pixels, in image order      (:75 -> normalize_rows)
//
// Everything is fp64 and follows the reference's operation ORDER, so the kept set and the values
// are bit-identical to the scipy + pandas route (checked against both in tests/):
//   * scipy's correlate1d, symmetric-kernel branch: tmp = x[l] * w[0]; then for the taps from the
//     farthest to the nearest: tmp += (x[l - j] + x[l + j]) * w[j] -- separately rounded multiply
//     and add (no FMA contraction: __dmul_rn / __dadd_rn), axis 0 first, then axis 1, boundary mode
//     'reflect' (d c b a | a b c d | d c b a);
//   * pandas' DataFrame.sum(axis=1) on the channel block: a plain sequential sum in channel order;
//   * IEEE division by the row sum.
//
// Kernels (all HBM/L2 streaming, one thread per element, consecutive threads on consecutive
// (w, c) addresses):  cast_div -> blur axis 0 -> blur axis 1 -> row sums + keep flags ->
// exclusive scan of the flags (cub::DeviceScan) -> normalise + compact.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace pixie {

namespace {

struct BlurTaps {
    int radius;
    double w[kMaxBlurRadius + 1];  // w[0] = centre tap, w[j] = tap at distance j
};

__device__ __forceinline__ int reflect_index(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
}

template <typename T>
__global__ void __launch_bounds__(256)
cast_div_kernel(const T *__restrict__ img, int64_t total, int C, const double *__restrict__ norm,
                double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const double v = (double)__ldcs(img + i);
        out[i] = norm ? __ddiv_rn(v, norm[i % C]) : v;
    }
}

// one 1-D pass of scipy's symmetric correlate1d along the axis whose element stride is `astride`
// and length `alen`; `pos_div` = elements per step of that axis's index (to recover the index)
__global__ void __launch_bounds__(256)
blur_axis_kernel(const double *__restrict__ in, double *__restrict__ out, int64_t total,
                 int64_t astride, int alen, BlurTaps taps)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int l = (int)((i / astride) % alen);
        const double *base = in + (i - (int64_t)l * astride);
        double tmp = __dmul_rn(base[(int64_t)l * astride], taps.w[0]);
        if (l >= taps.radius && l + taps.radius < alen) {
            for (int j = taps.radius; j >= 1; --j) {
                const double a = base[(int64_t)(l - j) * astride], b = base[(int64_t)(l + j) * astride];
                tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), taps.w[j]));
            }
        } else {
            for (int j = taps.radius; j >= 1; --j) {
                const double a = base[(int64_t)reflect_index(l - j, alen) * astride];
                const double b = base[(int64_t)reflect_index(l + j, alen) * astride];
                tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), taps.w[j]));
            }
        }
        out[i] = tmp;
    }
}

// Axis-0 pass for the default radius (sigma = 2 -> 8 taps either side): the neighbours along the
// image-row axis are W * C elements apart, so the generic kernel re-reads every input 17 times from
// L2 (4.5 GB per 1024 x 1024 x 32 FOV).  Here a thread owns one (w, c) column of a strip of S
// consecutive image rows: S + 2R loads (coalesced across the threads of a warp) for S outputs, the
// window in registers, the same operation order.
template <int R, int S>
__global__ void __launch_bounds__(128)
blur_axis0_strip_kernel(const double *__restrict__ in, double *__restrict__ out, int H, int64_t WC,
                        BlurTaps taps)
{
    const int64_t nstrips = (H + S - 1) / S;
    const int64_t total = nstrips * WC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t col = t % WC;
        const int h0 = (int)(t / WC) * S;
        const double *base = in + col;
        double win[S + 2 * R];
        if (h0 >= R && h0 + S + R <= H) {
#pragma unroll
            for (int k = 0; k < S + 2 * R; ++k) win[k] = base[(int64_t)(h0 - R + k) * WC];
        } else {
#pragma unroll
            for (int k = 0; k < S + 2 * R; ++k)
                win[k] = base[(int64_t)reflect_index(h0 - R + k, H) * WC];
        }
#pragma unroll
        for (int o = 0; o < S; ++o) {
            double tmp = __dmul_rn(win[o + R], taps.w[0]);
#pragma unroll
            for (int j = R; j >= 1; --j)
                tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(win[o + R - j], win[o + R + j]), taps.w[j]));
            if (h0 + o < H) out[(int64_t)(h0 + o) * WC + col] = tmp;
        }
    }
}

__global__ void __launch_bounds__(256)
row_filter_kernel(const double *__restrict__ x, int64_t n, int C, double thresh,
                  double *__restrict__ rowsum, int32_t *__restrict__ flags)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double *r = x + i * C;
        double s = 0.0;
        bool any = false;
        for (int c = 0; c < C; ++c) {
            const double v = r[c];
            s = __dadd_rn(s, v);
            any |= (v != 0.0);
        }
        rowsum[i] = s;
        flags[i] = (s > thresh && any) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
normalize_compact_kernel(const double *__restrict__ x, int64_t n, int C, int W,
                         const double *__restrict__ rowsum, const int32_t *__restrict__ flags,
                         const int32_t *__restrict__ pos, const int32_t *__restrict__ seg,
                         double *__restrict__ X64, float *__restrict__ X32, int64_t ldX32,
                         int32_t *__restrict__ row_index, int32_t *__restrict__ col_index,
                         int32_t *__restrict__ labels_out, int64_t *__restrict__ n_kept)
{
    const int64_t total = n * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / C;
        const int c = (int)(e - i * C);
        if (c == 0 && i == n - 1) *n_kept = (int64_t)pos[i] + flags[i];
        if (!flags[i]) continue;
        const int64_t j = pos[i];
        const double v = __ddiv_rn(x[e], rowsum[i]);
        if (X64) X64[j * C + c] = v;
        if (X32) X32[j * ldX32 + c] = (float)v;
        if (c == 0) {
            row_index[j] = (int32_t)(i / W);
            col_index[j] = (int32_t)(i % W);
            if (labels_out) labels_out[j] = seg ? seg[i] : 0;
        }
    }
}

int grid_for(int64_t work, int num_sms)
{
    int64_t b = (work + 255) / 256;
    if (b < 1) b = 1;
    if (b > (int64_t)num_sms * 16) b = (int64_t)num_sms * 16;
    return (int)b;
}

}  // namespace

size_t preprocess_scan_bytes(int64_t n) { return (size_t)(n / 16 + 65536); }

cudaError_t launch_preprocess(const void *img, int img_is_f64, int H, int W, int C, const double *norm,
                              const double *taps_host, int radius, double thresh,
                              const int32_t *seg, double *blurred, double *tmp, double *rowsum,
                              int32_t *flags, int32_t *pos, void *scan_tmp, size_t scan_bytes,
                              double *X64, float *X32, int64_t ldX32, int32_t *row_index,
                              int32_t *col_index, int32_t *labels_out, int64_t *n_kept,
                              int stop_after_blur, int num_sms, cudaStream_t stream)
{
    const int64_t n = (int64_t)H * W, total = n * C;
    BlurTaps taps;
    taps.radius = radius;
    for (int j = 0; j <= kMaxBlurRadius; ++j) taps.w[j] = (taps_host && j <= radius) ? taps_host[j] : 0.0;
    if (img_is_f64)
        cast_div_kernel<double><<<grid_for(total, num_sms), 256, 0, stream>>>(
            static_cast<const double *>(img), total, C, norm, blurred);
    else
        cast_div_kernel<float><<<grid_for(total, num_sms), 256, 0, stream>>>(
            static_cast<const float *>(img), total, C, norm, blurred);
    count_launch();
    if (radius > 0) {
        // axis 0 (image rows, stride W * C), then axis 1 (stride C): scipy's order
        if (radius == 8) {
            constexpr int S = 16;
            const int64_t work = (int64_t)((H + S - 1) / S) * W * C;
            int64_t b = (work + 127) / 128;
            if (b > (int64_t)num_sms * 32) b = (int64_t)num_sms * 32;
            blur_axis0_strip_kernel<8, S><<<(int)b, 128, 0, stream>>>(blurred, tmp, H,
                                                                    (int64_t)W * C, taps);
        } else {
            blur_axis_kernel<<<grid_for(total, num_sms), 256, 0, stream>>>(blurred, tmp, total,
                                                                          (int64_t)W * C, H, taps);
        }
        blur_axis_kernel<<<grid_for(total, num_sms), 256, 0, stream>>>(tmp, blurred, total,
                                                                      (int64_t)C, W, taps);
        count_launch(2);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || stop_after_blur) return e;
    row_filter_kernel<<<grid_for(n, num_sms), 256, 0, stream>>>(blurred, n, C, thresh, rowsum, flags);
    count_launch();
    size_t need = 0;
    e = cub::DeviceScan::ExclusiveSum(nullptr, need, flags, pos, (int)n, stream);
    if (e != cudaSuccess) return e;
    if (need > scan_bytes) return cudaErrorMemoryAllocation;
    e = cub::DeviceScan::ExclusiveSum(scan_tmp, need, flags, pos, (int)n, stream);
    if (e != cudaSuccess) return e;
    count_launch(2);
    normalize_compact_kernel<<<grid_for(total, num_sms), 256, 0, stream>>>(
        blurred, n, C, W, rowsum, flags, pos, seg, X64, X32, ldX32, row_index, col_index,
        labels_out, n_kept);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
