/*
 * pixie_oracle.c -- CPU oracle for the Pixie SOM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under ark_analysis_b200/ imports, links or executes it.
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in pyFlowSOM 0.1.16 (Cython + C), a
 * third-party dependency that is not vendored in /root/reference (pyproject.toml:48,
 * uv.lock:3005-3013) and is not installable in this sandbox.  The reference's own tests hold no
 * golden vectors for SOM weights or BMU labels, so this restatement is
 * anchored on (i) the reference's call sites, src/ark/phenotyping/cluster_helpers.py:106-109
 * (som) and :152-157 (map_data_to_nodes), (ii) the invariants its tests pin (labels in 1..K,
 * same-seed determinism, shapes) and (iii) the published FlowSOM som.c algorithm (C_SOM /
 * C_mapDataToCodes) that pyFlowSOM wraps, restated in SURVEY.md Appendix A.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march=native, no -ffast-math), i.e. what a generic
 * manylinux wheel of the reference dependency is built with.  See oracle/Makefile.
 *
 * Layout convention: all matrices are row-major (C order) here.  pyFlowSOM transposes to
 * Fortran order internally; the arithmetic per (row, node) pair is the same sequence of
 * operations, so the result is identical.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PIXIE_TILE 128 /* mini-batch interleave unit of the batch SOM */

/* --------------------------------------------------------------------------------------------
 * a1: map_data_to_nodes (pyFlowSOM C_mapDataToCodes), call site cluster_helpers.py:152-157.
 *   for each row i: minid = -1, mindist = DBL_MAX
 *     for each node cd in index order: d = sqrt(sum_j (x_ij - w_cd,j)^2)  (sequential in j, fp64,
 *     separate multiply and add); if (d < mindist) { mindist = d; minid = cd; }
 *   label = minid + 1 (1-indexed; a row holding NaN keeps minid = -1 -> label 0), dist = mindist.
 * ------------------------------------------------------------------------------------------ */
static inline int nearest_node(const double *nodes, int K, int C, const double *x, double *dist_out)
{
    int minid = -1;
    double mindist = DBL_MAX;
    for (int cd = 0; cd < K; ++cd) {
        const double *w = nodes + (size_t)cd * C;
        double acc = 0.0;
        for (int j = 0; j < C; ++j) {
            double tmp = x[j] - w[j];
            acc += tmp * tmp;
        }
        double d = sqrt(acc);
        if (d < mindist) {
            mindist = d;
            minid = cd;
        }
    }
    if (dist_out) *dist_out = mindist;
    return minid;
}

void oracle_map_data_to_nodes(const double *nodes, int K, const double *data, int64_t m, int C,
                              int32_t *labels, double *dists)
{
    for (int64_t i = 0; i < m; ++i) {
        double d;
        int id = nearest_node(nodes, K, C, data + (size_t)i * C, &d);
        labels[i] = id + 1;
        if (dists) dists[i] = d;
    }
}

/* Same arithmetic, fp32 inputs promoted to fp64 on the fly (the parity protocol of DESIGN.md:
 * both sides consume the same fp32-representable values).  ld = row pitch in floats. */
void oracle_map_data_to_nodes_f32(const float *nodes, int K, const float *data, int64_t m, int C,
                                  int64_t ld, int32_t *labels, double *dists)
{
    double *wd = (double *)malloc((size_t)K * C * sizeof(double));
    double *xd = (double *)malloc((size_t)C * sizeof(double));
    for (size_t t = 0; t < (size_t)K * C; ++t) wd[t] = (double)nodes[t];
    for (int64_t i = 0; i < m; ++i) {
        const float *x = data + (size_t)i * ld;
        for (int j = 0; j < C; ++j) xd[j] = (double)x[j];
        double d;
        int id = nearest_node(wd, K, C, xd, &d);
        labels[i] = id + 1;
        if (dists) dists[i] = d;
    }
    free(xd);
    free(wd);
}

/* Row-range variant used by the multi-threaded timing driver (bench.py --impl reference): the
 * reference's only parallelism is FOV-level process parallelism (pixel_som_clustering.py:257),
 * i.e. independent row ranges; each thread runs the scalar loop above on its own range. */
struct mt_job {
    const double *nodes, *data;
    int K, C;
    int64_t lo, hi;
    int32_t *labels;
    double *dists;
};

static void *mt_worker(void *arg)
{
    struct mt_job *j = (struct mt_job *)arg;
    oracle_map_data_to_nodes(j->nodes, j->K, j->data + (size_t)j->lo * j->C, j->hi - j->lo, j->C,
                             j->labels + j->lo, j->dists ? j->dists + j->lo : NULL);
    return NULL;
}

void oracle_map_data_to_nodes_mt(const double *nodes, int K, const double *data, int64_t m, int C,
                                 int32_t *labels, double *dists, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    struct mt_job *jobs = (struct mt_job *)malloc((size_t)nthreads * sizeof(struct mt_job));
    for (int t = 0; t < nthreads; ++t) {
        struct mt_job j = {nodes, data, K, C, m * t / nthreads, m * (t + 1) / nthreads, labels,
                           dists};
        jobs[t] = j;
        pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
}

/* --------------------------------------------------------------------------------------------
 * Chebyshev grid distance between SOM nodes (pyFlowSOM `nhbrdist`), node k <-> grid point
 * (k / ydim, k % ydim), i.e. [(x, y) for x in range(xdim) for y in range(ydim)].
 * ------------------------------------------------------------------------------------------ */
void oracle_grid_chebyshev(int xdim, int ydim, double *D)
{
    int K = xdim * ydim;
    for (int a = 0; a < K; ++a)
        for (int b = 0; b < K; ++b) {
            int dx = abs(a / ydim - b / ydim), dy = abs(a % ydim - b % ydim);
            D[(size_t)a * K + b] = (double)(dx > dy ? dx : dy);
        }
}

/* --------------------------------------------------------------------------------------------
 * a2: som (pyFlowSOM C_SOM), call site cluster_helpers.py:106-109.  ONLINE, sequential.
 * Restated from SURVEY.md Appendix A -- UNVERIFIED against the binary (RNG details in
 * particular).  `nodes` holds the initial codebook on entry (K rows sampled from the data by the
 * caller) and the trained codebook on exit.  Returns the number of iterations executed.
 * ------------------------------------------------------------------------------------------ */
int64_t oracle_som_online(const double *data, int64_t n, int C, double *nodes, int K,
                          const double *nhbrdist, double alpha0, double alpha1, double radius0,
                          double radius1, int rlen, unsigned int seed)
{
    int64_t niter = (int64_t)rlen * n;
    double threshold = radius0;
    double threshold_step = (radius0 - radius1) / (double)niter;
    double change = 1.0;
    int64_t k;
    srand(seed);
    for (k = 0; k < niter; ++k) {
        if (k % n == 0) {
            if (change < 1.0) k = niter; /* early stop after this iteration */
            change = 0.0;
        }
        int64_t i = (int64_t)((double)n * ((double)rand() / ((double)RAND_MAX + 1.0)));
        const double *x = data + (size_t)i * C;
        int nearest = nearest_node(nodes, K, C, x, NULL);
        if (nearest < 0) nearest = 0;
        if (threshold < 1.0) threshold = 0.5;
        double alpha = alpha0 - (alpha0 - alpha1) * (double)k / (double)niter;
        for (int cd = 0; cd < K; ++cd) {
            if (nhbrdist[(size_t)cd * K + nearest] > threshold) continue;
            double *w = nodes + (size_t)cd * C;
            for (int j = 0; j < C; ++j) {
                double tmp = x[j] - w[j];
                change += fabs(tmp);
                w[j] += tmp * alpha;
            }
        }
        threshold -= threshold_step;
    }
    return k;
}

/* --------------------------------------------------------------------------------------------
 * Batch SOM -- the fp64 restatement of the algorithm the B200 path runs (BASELINE.json
 * north_star: per-node delta aggregation + one allreduce per step).  This is NOT the reference's
 * online algorithm; it is the oracle for the "weights within 1e-4 relative" claim.  Specified in
 * DESIGN.md section 4:
 *   T = rlen * B steps.  Step t uses mini-batch m = t % B = rows i with (i / 128) % B == m.
 *   W32 = fp32(W64); b_i = BMU(x_i; W32) by the a1 rule; S_b += x_i, n_b += 1 (fp64).
 *   r_t = r0 - (r0 - r1) t / T, r_eff = r_t < 1 ? 0.5 : r_t, sigma = r_eff / 2,
 *   H[k,b] = exp(-D[k,b]^2 / (2 sigma^2)), alpha_t = a0 - (a0 - a1) t / T,
 *   num_k = sum_b H[k,b] S_b, den_k = sum_b H[k,b] n_b,
 *   den_k > 0: W64_k += (1 - (1 - alpha_t)^den_k) * (num_k / den_k - W64_k).
 * X is fp32 (the device matrix), ld = row pitch in floats.  W64 in/out.
 * ------------------------------------------------------------------------------------------ */
void oracle_som_batch(const float *X, int64_t n, int C, int64_t ld, double *W64, int xdim,
                      int ydim, int rlen, int B, double alpha0, double alpha1, double radius0,
                      double radius1)
{
    int K = xdim * ydim;
    int64_t T = (int64_t)rlen * B;
    int64_t ntiles = (n + PIXIE_TILE - 1) / PIXIE_TILE;
    double *D = (double *)malloc((size_t)K * K * sizeof(double));
    double *Wr = (double *)malloc((size_t)K * C * sizeof(double));
    double *S = (double *)malloc((size_t)K * C * sizeof(double));
    double *cnt = (double *)malloc((size_t)K * sizeof(double));
    double *xd = (double *)malloc((size_t)C * sizeof(double));
    double *num = (double *)malloc((size_t)C * sizeof(double));
    oracle_grid_chebyshev(xdim, ydim, D);
    for (int64_t t = 0; t < T; ++t) {
        int64_t m = t % B;
        for (size_t q = 0; q < (size_t)K * C; ++q) Wr[q] = (double)(float)W64[q];
        memset(S, 0, (size_t)K * C * sizeof(double));
        memset(cnt, 0, (size_t)K * sizeof(double));
        for (int64_t tile = m; tile < ntiles; tile += B) {
            int64_t lo = tile * PIXIE_TILE, hi = lo + PIXIE_TILE;
            if (hi > n) hi = n;
            for (int64_t i = lo; i < hi; ++i) {
                const float *x = X + (size_t)i * ld;
                for (int j = 0; j < C; ++j) xd[j] = (double)x[j];
                int b = nearest_node(Wr, K, C, xd, NULL);
                if (b < 0) continue;
                for (int j = 0; j < C; ++j) S[(size_t)b * C + j] += xd[j];
                cnt[b] += 1.0;
            }
        }
        double frac = (double)t / (double)T;
        double r = radius0 - (radius0 - radius1) * frac;
        double r_eff = r < 1.0 ? 0.5 : r;
        double sigma = 0.5 * r_eff;
        double inv2s2 = 1.0 / (2.0 * sigma * sigma);
        double alpha = alpha0 - (alpha0 - alpha1) * frac;
        for (int k = 0; k < K; ++k) {
            double den = 0.0;
            for (int j = 0; j < C; ++j) num[j] = 0.0;
            for (int b = 0; b < K; ++b) {
                if (cnt[b] == 0.0) continue;
                double d = D[(size_t)k * K + b];
                double h = exp(-d * d * inv2s2);
                den += h * cnt[b];
                for (int j = 0; j < C; ++j) num[j] += h * S[(size_t)b * C + j];
            }
            if (den > 0.0) {
                double beta = 1.0 - pow(1.0 - alpha, den);
                double *w = W64 + (size_t)k * C;
                for (int j = 0; j < C; ++j) w[j] += beta * (num[j] / den - w[j]);
            }
        }
    }
    free(num);
    free(xd);
    free(cnt);
    free(S);
    free(Wr);
    free(D);
}

/* Per-node sums and counts for a label array -- the N1 row (compute_pixel_cluster_channel_avg,
 * pixel_cluster_utils.py:369-404): sum of channel values and pixel count per SOM cluster. */
void oracle_cluster_sums_f32(const float *X, int64_t n, int C, int64_t ld, const int32_t *labels,
                             int K, double *S, double *cnt)
{
    memset(S, 0, (size_t)K * C * sizeof(double));
    memset(cnt, 0, (size_t)K * sizeof(double));
    for (int64_t i = 0; i < n; ++i) {
        int b = labels[i] - 1;
        if (b < 0 || b >= K) continue;
        const float *x = X + (size_t)i * ld;
        for (int j = 0; j < C; ++j) S[(size_t)b * C + j] += (double)x[j];
        cnt[b] += 1.0;
    }
}
