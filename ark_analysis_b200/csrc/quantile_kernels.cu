// quantile_kernels.cu -- per-channel order statistics of the non-zero entries of the pixel matrix
// (SURVEY.md section 8f, row N3): the device half of
//     fov_full_pixel_data.replace(0, np.nan).quantile(q=0.999, axis=0)
// (/root/reference/src/ark/phenotyping/pixie_preprocessing.py:405-410, :424-427), whose mean over
// the FOVs becomes the normalisation row the SOM divides by.
//
// pandas / numpy compute a 'linear' quantile: with m valid values (non-zero, non-NaN) in a column
// and v = (m - 1) * q, the result interpolates between the order statistics of rank floor(v) and
// floor(v) + 1.  The kernels deliver exactly those two values and m for every channel; the
// interpolation (a handful of fp64 operations on C-element arrays) stays on the host, written with
// numpy's own formula, so the result is bit-identical to the reference's.
//
// Exact selection without sorting: a most-significant-digit radix select on the order-preserving
// 64-bit image of the fp64 values, 16 bits per pass (4 passes), all channels at once:
//   hist pass   : every valid element whose upper digits equal the channel's current prefix adds 1
//                 to hist[c][digit] (global atomics; consecutive lanes = consecutive channels of a
//                 row, so loads are coalesced and a warp's atomics go to different histograms);
//   select pass : one CTA per channel scans the 65536 bins, finds the bin holding the wanted rank,
//                 appends it to the prefix, rebases the rank and clears the histogram;
// the first select pass also learns m (the histogram's total) and the rank, the last one whether the
// selected value is repeated past that rank; one more pass finds the smallest larger value (the
// rank + 1 statistic otherwise).  5 streaming passes over X in total.
#include "common.cuh"

namespace pixie {

namespace {

constexpr int kDigitBits = 16;
constexpr int kBins = 1 << kDigitBits;

// order-preserving map fp64 -> uint64 (negative values included, -0.0 < +0.0 never matters: zeros
// are skipped)
__device__ __forceinline__ unsigned long long ordered_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k)
{
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ bool valid(double v) { return v != 0.0 && v == v; }

struct QState {            // per channel, in the workspace
    unsigned long long prefix;   // digits selected so far (right-aligned)
    unsigned long long rank;     // wanted rank among the elements matching the prefix
    unsigned long long m;        // valid elements in the column
    unsigned long long dup_ok;   // the value of rank r is repeated past rank r (so rank r+1 is it too)
    unsigned long long next_key; // smallest key > the selected one (final pass), ~0 if none
    unsigned long long pad[3];
};

__global__ void __launch_bounds__(256)
quant_hist_kernel(const double *__restrict__ X, int64_t n, int C, int64_t ldX,
                  const QState *__restrict__ st, int pass, unsigned int *__restrict__ hist)
{
    const int shift = 64 - kDigitBits * (pass + 1);
    const int64_t total = n * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / C;
        const int c = (int)(e - i * C);
        const double v = X[i * ldX + c];
        if (!valid(v)) continue;
        const unsigned long long k = ordered_key(v);
        if (pass > 0 && (k >> (shift + kDigitBits)) != st[c].prefix) continue;
        atomicAdd(hist + (size_t)c * kBins + (unsigned)((k >> shift) & (kBins - 1)), 1u);
    }
}

// one CTA per channel: find the bin that holds the wanted rank.  Pass 0 also learns m (the total
// of the first histogram) and derives the rank: floor((m - 1) * q) in fp64, numpy's virtual index
// for the 'linear' method.
__global__ void __launch_bounds__(1024)
quant_select_kernel(QState *st, unsigned int *__restrict__ hist, int pass, double q)
{
    __shared__ unsigned long long s_part[1024];
    const int c = blockIdx.x, tid = threadIdx.x;
    unsigned int *h = hist + (size_t)c * kBins;
    constexpr int per = kBins / 1024;   // 64 consecutive bins per thread
    const unsigned long long prefix0 = st[c].prefix, rank0 = st[c].rank, m0 = st[c].m;
    unsigned long long sum = 0;
    for (int j = 0; j < per; ++j) sum += h[tid * per + j];
    s_part[tid] = sum;
    __syncthreads();
    // inclusive scan of the 1024 partial sums (Hillis-Steele on shared memory)
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long add = tid >= o ? s_part[tid - o] : 0ull;
        __syncthreads();
        s_part[tid] += add;
        __syncthreads();
    }
    const unsigned long long incl = s_part[tid], excl = incl - sum, total = s_part[1023];
    unsigned long long m = m0, rank = rank0;
    if (pass == 0) {
        m = total;
        rank = 0;
        if (m > 0) {
            rank = (unsigned long long)floor(__dmul_rn((double)(m - 1), q));
            if (rank > m - 1) rank = m - 1;
        }
    }
    __syncthreads();   // every thread holds its copy of the state before anyone rewrites it
    if (tid == 0 && pass == 0) {
        st[c].m = m;
        st[c].next_key = ~0ull;
        st[c].dup_ok = 0ull;
    }
    if (m > 0 && rank >= excl && rank < incl) {   // exactly one thread
        unsigned long long run = excl;
        for (int j = 0; j < per; ++j) {
            const unsigned long long hj = h[tid * per + j];
            if (rank < run + hj) {
                st[c].prefix = (prefix0 << kDigitBits) | (unsigned long long)(tid * per + j);
                st[c].rank = rank - run;
                if (pass == 64 / kDigitBits - 1) st[c].dup_ok = (rank - run + 1 < hj) ? 1ull : 0ull;
                break;
            }
            run += hj;
        }
    }
    __syncthreads();
    for (int j = 0; j < per; ++j) h[tid * per + j] = 0u;   // ready for the next pass
}

// the smallest key above the selected value (the rank + 1 statistic when the value is not repeated)
__global__ void __launch_bounds__(256)
quant_next_kernel(const double *__restrict__ X, int64_t n, int C, int64_t ldX, QState *st)
{
    const int64_t total = n * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / C;
        const int c = (int)(e - i * C);
        const double v = X[i * ldX + c];
        if (!valid(v)) continue;
        const unsigned long long k = ordered_key(v);
        if (k > st[c].prefix && k < st[c].next_key)   // racy pre-test only prunes; atomicMin decides
            atomicMin(&st[c].next_key, k);
    }
}

__global__ void quant_finish_kernel(const QState *st, int C, double q, double *lo, double *hi,
                                    int64_t *m_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const unsigned long long m = st[c].m;
    m_out[c] = (int64_t)m;
    if (m == 0) {
        lo[c] = hi[c] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    const double v = __dmul_rn((double)(m - 1), q);
    unsigned long long r = (unsigned long long)floor(v);
    if (r > m - 1) r = m - 1;
    const double a = key_to_double(st[c].prefix);
    lo[c] = a;
    // the statistic of rank r + 1: the same value when it is repeated past rank r (or r is the
    // last rank: numpy clips the upper index), else the next larger one
    hi[c] = (st[c].dup_ok || r + 1 > m - 1) ? a : key_to_double(st[c].next_key);
}

int grid_q(int64_t work, int num_sms)
{
    int64_t b = (work + 255) / 256;
    if (b < 1) b = 1;
    if (b > (int64_t)num_sms * 16) b = (int64_t)num_sms * 16;
    return (int)b;
}

}  // namespace

size_t quantile_workspace_bytes(int C)
{
    return (size_t)C * kBins * sizeof(unsigned int) + (size_t)C * sizeof(QState) + 256;
}

cudaError_t launch_column_quantile(const double *X, int64_t n, int C, int64_t ldX, double q,
                                   double *lo, double *hi, int64_t *m_out, void *workspace,
                                   int num_sms, cudaStream_t stream)
{
    unsigned int *hist = static_cast<unsigned int *>(workspace);
    QState *st = reinterpret_cast<QState *>(static_cast<char *>(workspace) +
                                            (size_t)C * kBins * sizeof(unsigned int));
    cudaError_t e = cudaMemsetAsync(workspace, 0, quantile_workspace_bytes(C), stream);
    if (e != cudaSuccess) return e;
    const int64_t total = n * C;
    const int g = grid_q(total, num_sms), gc = (C + 127) / 128;
    if (n > 0) {
        for (int pass = 0; pass < 64 / kDigitBits; ++pass) {
            quant_hist_kernel<<<g, 256, 0, stream>>>(X, n, C, ldX, st, pass, hist);
            quant_select_kernel<<<C, 1024, 0, stream>>>(st, hist, pass, q);
            count_launch(2);
        }
        quant_next_kernel<<<g, 256, 0, stream>>>(X, n, C, ldX, st);
        count_launch();
    }
    quant_finish_kernel<<<gc, 128, 0, stream>>>(st, C, q, lo, hi, m_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
