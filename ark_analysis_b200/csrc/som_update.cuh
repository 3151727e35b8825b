// som_update.cuh -- the batch-SOM codebook update as ONE device function,
// shared by the stand-alone som_apply_kernel (step-by-step path, pixie_som_apply_f64) and by the
// whole-pass training kernel (bmu_tc_kernel.cuh), so the two paths are bit-identical by
// construction: the arithmetic below does not depend on how many threads run it.
//
//   h[b]      = n_b == 0 ? 0 : exp(-d(k, b)^2 / (2 sigma^2))        d = Chebyshev grid distance
//   dot(k, c) = sum_b h[b] * SN[b][c],  c = 0..C  (c == C: den_k = sum_b h[b] n_b)
//               eight slices b = s, s + 8, ... (ascending), combined ((0+1)+(2+3))+((4+5)+(6+7))
//   beta      = 1 - (1 - alpha)^den_k;   w_kc += beta * (dot(k, c) / den_k - w_kc)     (den_k > 0)
//
// Restated in fp64 by oracle/pixie_oracle.c oracle_som_batch (which sums over b in one sequence:
// the two agree to ~1e-16 relative, the parity tolerance is 1e-4).
#pragma once
#include <stdint.h>

namespace pixie {

constexpr int kUpdNodes = 4;   // nodes a CTA updates per sweep over SN
constexpr int kUpdSlices = 8;  // b-slices of the neighbourhood sum

__host__ __device__ inline uint32_t som_update_scratch_bytes(int C, int K)
{
    // s_h [kUpdNodes][K] + s_part [kUpdNodes][kUpdSlices][C + 1], doubles
    return (uint32_t)(kUpdNodes * K + kUpdNodes * kUpdSlices * (C + 1)) * 8u;
}

// Called by `nthr` threads (a multiple of 32; tid = 0 .. nthr - 1) of one CTA that can meet at
// `sync()`.  Updates nodes kfirst, kfirst + kstride, ... < K.  `elem(k, c, w32)` is called once per
// (node, channel) by the thread that owns it, `node_done(k, lane, nrm2, neg)` by every lane of the
// warp that finished node k (nrm2 = ||fp32(w_k)||^2 summed in fp64: lane-strided partial sums, then
// the xor-shuffle tree 16, 8, 4, 2, 1 -- the order codebook_prep_kernel uses).
template <class Sync, class Elem, class NodeDone>
__device__ __forceinline__ void som_update_nodes(const double *SN, double *W64, float *W32, int K,
                                                 int C, int ydim, double inv2s2, double alpha,
                                                 int kfirst, int kstride, int tid, int nthr,
                                                 double *scratch, Sync sync, Elem elem,
                                                 NodeDone node_done)
{
    const int ld = C + 1;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    double *s_h = scratch;                                  // [kUpdNodes][K]
    double *s_part = scratch + (size_t)kUpdNodes * K;       // [kUpdNodes][kUpdSlices][ld]
    const int ncb = (ld + 31) >> 5;
    for (int k0 = kfirst; k0 < K; k0 += kstride * kUpdNodes) {
        int nk = 0;
        while (nk < kUpdNodes && k0 + nk * kstride < K) ++nk;
        // ---- neighbourhood weights of this sweep's nodes.  The node counts come from L2 (~1 us a
        // round trip): a thread fetches those of all its (node, b) pairs before the first exp().
        for (int i0 = tid; i0 < nk * K; i0 += 4 * nthr) {
            double cnt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nthr;
                cnt[u] = i < nk * K ? __ldcg(SN + (size_t)(i % K) * ld + C) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nthr;
                if (i < nk * K) {
                    const int j = i / K, b = i - j * K;
                    const int k = k0 + j * kstride;
                    const int dx = abs(k / ydim - b / ydim), dy = abs(k % ydim - b % ydim);
                    const double d = (double)(dx > dy ? dx : dy);
                    s_h[j * K + b] = cnt[u] == 0.0 ? 0.0 : exp(-d * d * inv2s2);  // empty nodes are skipped
                }
            }
        }
        sync();
        // ---- sliced dot products: warp task = (slice s, 32-column block cb), lanes = columns.
        // Sixteen rows of SN are in flight per lane (this phase is nothing but L2 latency: with
        // four, a 20 x 20 map took 13 round trips per task, 31 us per step); the adds stay in
        // ascending b order.
        constexpr int kInFlight = 16;
        for (int task = warp; task < kUpdSlices * ncb; task += nwarps) {
            const int s = task % kUpdSlices, cb = task / kUpdSlices;
            const int c = cb * 32 + lane;
            double acc[kUpdNodes];
#pragma unroll
            for (int j = 0; j < kUpdNodes; ++j) acc[j] = 0.0;
            if (c < ld) {
                const double *col = SN + c;
                for (int b0 = s; b0 < K; b0 += kInFlight * kUpdSlices) {
                    double v[kInFlight];
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) {
                        const int b = b0 + u * kUpdSlices;
                        v[u] = b < K ? __ldcg(col + (size_t)b * ld) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) {
                        const int b = b0 + u * kUpdSlices;
                        if (b < K) {
#pragma unroll
                            for (int j = 0; j < kUpdNodes; ++j)
                                if (j < nk) acc[j] = fma(s_h[j * K + b], v[u], acc[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kUpdNodes; ++j)
                    if (j < nk) s_part[(j * kUpdSlices + s) * ld + c] = acc[j];
            }
        }
        sync();
        // ---- node j of the sweep is finished by warp j
        if (warp < nk) {
            const int j = warp, k = k0 + j * kstride;
            const double *pp = s_part + (size_t)j * kUpdSlices * ld;
            auto total = [&](int c) {
                return ((pp[0 * ld + c] + pp[1 * ld + c]) + (pp[2 * ld + c] + pp[3 * ld + c])) +
                       ((pp[4 * ld + c] + pp[5 * ld + c]) + (pp[6 * ld + c] + pp[7 * ld + c]));
            };
            const double den = total(C);
            const double beta = den > 0.0 ? 1.0 - pow(1.0 - alpha, den) : 0.0;
            double nrm2 = 0.0;
            bool neg = false;
            for (int c = lane; c < C; c += 32) {
                double w = W64[(size_t)k * C + c];
                if (den > 0.0) {
                    w += beta * (total(c) / den - w);
                    W64[(size_t)k * C + c] = w;
                }
                const float wf = (float)w;
                W32[(size_t)k * C + c] = wf;
                elem(k, c, wf);
                nrm2 += (double)wf * (double)wf;
                if (__float_as_int(wf) < 0) neg = true;
            }
            for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
            neg = __any_sync(0xffffffffu, neg);
            node_done(k, lane, nrm2, neg);
        }
        sync();  // scratch is reused by the next sweep
    }
}

}  // namespace pixie
