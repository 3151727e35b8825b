// som_online.cu -- the reference's OWN training rule on the GPU ("parity mode").
//
// pyFlowSOM.som (call site /root/reference/src/ark/phenotyping/cluster_helpers.py:106-109) trains an
// ONLINE self-organising map: rlen * n strictly sequential single-sample updates, each of which
// depends on the codebook the previous one left behind (FlowSOM C_SOM, restated in
// oracle/pixie_oracle.c oracle_som_online and SURVEY.md Appendix A).  The production path of this
// library trains a batch SOM instead, which shards over GPUs; this kernel
// exists so that the sequential rule itself can run on the device, with the SAME floating-point
// operation sequence as the C restatement -- fp64, subtract / multiply / add separately rounded,
// channel order, sqrt, strict '<', lowest node index first, `threshold -= step` accumulated
// iteration by iteration -- and therefore bit-identical codebooks.  It does not shard: one CTA
// walks the sample sequence ("replicas only").
//
// One iteration = distances of the sample to all K nodes (thread k owns node k: a C-long chain of
// dependent fp64 adds is the floor of the iteration time), a lexicographic (distance, index)
// arg-min over the CTA, and the neighbourhood update with one warp per node row.  The codebook
// lives in shared memory as fp64 for the whole run; the sample rows are prefetched kPrefetch
// iterations ahead with cp.async (their indices are known in advance: the host draws them).
//
// The one deliberate deviation: `change` (sum of |x - w| over a pass, used only by the early-stop
// test between passes) is accumulated per thread and reduced in a fixed tree order, not in the
// reference's (node, channel) order.  It only ever feeds the comparison `change < 1.0`.
#include <float.h>

#include "common.cuh"

namespace pixie {

namespace {
constexpr int kOnThreads = 1024;
constexpr int kPrefetch = 16;   // sample rows in flight
constexpr int kRowRing = kPrefetch + 1;  // + the row the slowest warp may still be reading
constexpr int kIdxRing = 512;   // sample indices staged in shared memory (two halves)

__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src)
{
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(kOnThreads, 1)
som_online_kernel(const float *__restrict__ X, int64_t n, int C, int64_t ldX, double *__restrict__ W,
                  int xdim, int ydim, const int64_t *__restrict__ sample_idx, int64_t niter,
                  int64_t n_per_pass, double a0, double a1, double r0, double r1,
                  long long *__restrict__ iters_done)
{
    extern __shared__ double sm[];
    const int K = xdim * ydim;
    const int P = C | 1;  // odd row pitch: thread k reading w[k][j] is bank-conflict free
    double *w = sm;                                              // [K][P]
    double *red_d = w + (size_t)K * P;                           // [32] per-warp minima
    int *red_i = reinterpret_cast<int *>(red_d + 32);            // [32]
    long long *idx = reinterpret_cast<long long *>(red_i + 32);  // [kIdxRing]
    float *xring = reinterpret_cast<float *>(idx + kIdxRing);    // [kRowRing][C]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kOnThreads / 32;

    for (int e = tid; e < K * C; e += kOnThreads) w[(e / C) * P + (e % C)] = W[e];
    for (int e = tid; e < kIdxRing; e += kOnThreads) idx[e] = e < niter ? sample_idx[e] : 0;
    __syncthreads();

    // prologue of the row pipeline: rows of iterations 0 .. kPrefetch-2
    for (int d = 0; d < kPrefetch - 1; ++d) {
        if (tid < C && d < niter)
            cp_async4(xring + d * C + tid, X + (size_t)idx[d] * (size_t)ldX + tid);
        cp_async_commit();
    }

    double threshold = r0;
    const double threshold_step = (r0 - r1) / (double)niter;
    double chg = 0.0;          // this thread's share of `change`
    bool first_pass = true;    // `change` starts at 1.0, i.e. "not converged"
    int64_t kn = 0;            // k % n_per_pass
    long long it = 0;          // iterations executed == samples consumed
    for (int64_t k = 0; k < niter; ++k, ++it) {
        if (kn == 0) {
            if (!first_pass) {
                // reduce `change` over the CTA (fixed order: lanes, then warps)
                double c = chg;
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                __syncthreads();
                if (lane == 0) red_d[warp] = c;
                __syncthreads();
                double tot = 0.0;
                for (int q = 0; q < nwarp; ++q) tot += red_d[q];
                __syncthreads();
                if (tot < 1.0) k = niter;  // early stop after this iteration (reference quirk)
            }
            first_pass = false;
            chg = 0.0;
        }
        if (++kn == n_per_pass) kn = 0;

        // ---- pipeline: issue the row of iteration it + kPrefetch - 1, refill the index ring
        {
            const long long ahead = it + kPrefetch - 1;
            if (tid < C && ahead < niter)
                cp_async4(xring + (int)(ahead % kRowRing) * C + tid,
                          X + (size_t)idx[ahead % kIdxRing] * (size_t)ldX + tid);
            cp_async_commit();
            // when the cursor enters a half of the ring, the OTHER half is refilled with the indices
            // that follow it (nothing reads that half for the next kIdxRing/2 - kPrefetch iterations)
            if ((it % (kIdxRing / 2)) == 0 && it > 0) {
                const long long base = it + kIdxRing / 2;
                for (int e = tid; e < kIdxRing / 2; e += kOnThreads)
                    if (base + e < niter) idx[(base + e) % kIdxRing] = sample_idx[base + e];
            }
            cp_async_wait<kPrefetch - 1>();  // this thread's part of row `it` has landed
        }
        __syncthreads();  // S1: row `it` complete; previous update of w visible
        const float *x = xring + (int)(it % kRowRing) * C;

        // ---- nearest node: thread k < K owns node k
        double d = DBL_MAX;
        int id = 0x7fffffff;
        if (tid < K) {
            const double *wk = w + (size_t)tid * P;
            double acc = 0.0;
            for (int j = 0; j < C; ++j) {
                const double tmp = __dsub_rn((double)x[j], wk[j]);
                acc = __dadd_rn(acc, __dmul_rn(tmp, tmp));
            }
            const double dk = __dsqrt_rn(acc);
            if (dk < DBL_MAX) {  // NaN never wins; DBL_MAX is the loop's initial value
                d = dk;
                id = tid;
            }
        }
        if (warp * 32 < K) {
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, d, o);
                const int oi = __shfl_xor_sync(0xffffffffu, id, o);
                if (od < d || (od == d && oi < id)) {
                    d = od;
                    id = oi;
                }
            }
            if (lane == 0) {
                red_d[warp] = d;
                red_i[warp] = id;
            }
        }
        __syncthreads();  // S2
        int nearest = 0x7fffffff;
        {
            double best = DBL_MAX;
            const int used = (K + 31) >> 5;
            for (int q = 0; q < used; ++q) {
                const double od = red_d[q];
                const int oi = red_i[q];
                if (od < best || (od == best && oi < nearest)) {
                    best = od;
                    nearest = oi;
                }
            }
            if (nearest == 0x7fffffff) nearest = 0;  // every distance NaN: the oracle uses node 0
        }

        // ---- neighbourhood update, one warp per node row
        if (threshold < 1.0) threshold = 0.5;
        const double alpha = a0 - (a0 - a1) * (double)k / (double)niter;
        const int nx = nearest / ydim, ny = nearest - nx * ydim;
        for (int cd = warp; cd < K; cd += nwarp) {
            const int cx = cd / ydim, cy = cd - cx * ydim;
            const int dx = cx > nx ? cx - nx : nx - cx, dy = cy > ny ? cy - ny : ny - cy;
            if ((double)(dx > dy ? dx : dy) > threshold) continue;
            double *wc = w + (size_t)cd * P;
            for (int j = lane; j < C; j += 32) {
                const double tmp = __dsub_rn((double)x[j], wc[j]);
                chg += fabs(tmp);
                wc[j] = __dadd_rn(wc[j], __dmul_rn(tmp, alpha));
            }
        }
        threshold -= threshold_step;
        // S1 of the next iteration orders these writes before its reads.  The copy issued at the
        // top of iteration it + 1 targets the slot of row it - 1 (ring of kPrefetch + 1 rows), which
        // every thread finished reading before it passed S1 of iteration `it`.
    }
    __syncthreads();
    for (int e = tid; e < K * C; e += kOnThreads) W[e] = w[(e / C) * P + (e % C)];
    if (tid == 0 && iters_done) *iters_done = it;
}
}  // namespace

size_t som_online_smem_bytes(int C, int K)
{
    const size_t P = (size_t)(C | 1);
    return (size_t)K * P * 8 + 32 * 8 + 32 * 4 + (size_t)kIdxRing * 8 + (size_t)kRowRing * C * 4 + 16;
}

cudaError_t launch_som_online(const float *X, int64_t n, int C, int64_t ldX, double *W, int xdim,
                              int ydim, const int64_t *sample_idx, int64_t niter,
                              int64_t n_per_pass, double a0, double a1, double r0, double r1,
                              long long *iters_done, cudaStream_t stream)
{
    const size_t smem = som_online_smem_bytes(C, xdim * ydim);
    if (xdim * ydim > kOnThreads || C > kOnThreads || smem > 227u * 1024u) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(som_online_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    som_online_kernel<<<1, kOnThreads, smem, stream>>>(X, n, C, ldX, W, xdim, ydim, sample_idx,
                                                       niter, n_per_pass, a0, a1, r0, r1,
                                                       iters_done);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pixie
