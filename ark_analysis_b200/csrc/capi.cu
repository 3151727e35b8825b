// capi.cu -- the extern "C" surface declared in include/pixie_b200.h.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace pixie;

namespace pixie {
#ifdef PIXIE_PROFILE
static unsigned long long *g_trace = nullptr;   // device: 2 x 2^15 words
static unsigned int *g_trace_count = nullptr;   // device
#endif
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace pixie

namespace {

#define PX_CUDA(call)                          \
    do {                                       \
        cudaError_t e__ = (call);              \
        if (e__ != cudaSuccess) {              \
            set_last_cuda_error(e__, #call);   \
            return PIXIE_ERR_CUDA;             \
        }                                      \
    } while (0)

thread_local char g_last_error[512] = "";

void set_last_cuda_error(cudaError_t e, const char *what)
{
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__global__ void set_u64_kernel(unsigned long long *dst, unsigned long long v) { *dst = v; }

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// X [n x C] fp32, row pitch ldX: box = 32 channels x 128 rows, SWIZZLE_128B, OOB reads as zero.
bool make_x_tensor_map(CUtensorMap *tm, const float *X, int64_t n, int C, int64_t ldX)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)n};
    cuuint64_t gstride[1] = {(cuuint64_t)ldX * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)kTile};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(X), gdim, gstride,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled failed: %d", (int)r);
        return false;
    }
    return true;
}

// plan.tail8: the last 8 channels of the K-steps as 32-byte rows (box 8 x 128, SWIZZLE_32B)
bool make_x_tail_tensor_map(CUtensorMap *tm, const float *X, int64_t n, int C, int64_t ldX)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)n};
    cuuint64_t gstride[1] = {(cuuint64_t)ldX * sizeof(float)};
    cuuint32_t box[2] = {8u, (cuuint32_t)kTile};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(X), gdim, gstride,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled (tail) failed: %d", (int)r);
        return false;
    }
    return true;
}

int num_sms_current_device()
{
    static int cached[64];
    static bool have[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!have[dev]) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = v;
        have[dev] = true;
    }
    return cached[dev];
}

// CTAs of the cluster-sums kernel: one per SM (its grid barrier needs them co-resident)
int sum_parts()
{
    const int sms = num_sms_current_device();
    return sms < kSumParts ? sms : kSumParts;
}

// Candidate-window scale: 1 in production.  PIXIE_DELTA_SCALE < 1 shrinks the tensor-core
// candidate window below its proven bound -- tests use it to measure how much margin the bound has.
float delta_scale_from_env()
{
    const char *v = getenv("PIXIE_DELTA_SCALE");
    if (!v || !v[0]) return 1.0f;
    const float f = (float)atof(v);
    return (f > 0.f && f <= 64.f) ? f : 1.0f;
}

// ---- workspace carve-up
struct Workspace {
    float *wimg;
    CodebookAux *aux;
    float *partials;  // group tables of the fused sums: [kSumParts x NG <= 4][K][part_pitch(C)]
    float *parts;     // CTA parts: [kSumParts][K][part_pitch(C)]
    int32_t *labels_scratch;
    size_t total;
};

Workspace carve(void *base, int64_t n_visit, int C, int K)
{
    Workspace w{};
    size_t off = 0;
    const size_t wimg_max = (size_t)5 * 512 * 128;  // nblkW <= 5, Ntot <= 512
    w.wimg = reinterpret_cast<float *>((char *)base + off);
    off += align_up(wimg_max, 1024);
    w.aux = reinterpret_cast<CodebookAux *>((char *)base + off);
    off += 256;
    const size_t table = (size_t)K * part_pitch(C) * sizeof(float);
    w.partials = reinterpret_cast<float *>((char *)base + off);
    off += align_up((size_t)kSumParts * 4 * table, 256);
    w.parts = reinterpret_cast<float *>((char *)base + off);
    off += align_up((size_t)kSumParts * table, 256);
    w.labels_scratch = reinterpret_cast<int32_t *>((char *)base + off);
    const int64_t tiles = (n_visit + kTile - 1) / kTile + 1;
    off += align_up((size_t)tiles * kTile * sizeof(int32_t), 256);
    w.total = off;
    return w;
}

// the group tables are accumulated into in place: they start a launch at zero (every step leaves
// them so, but the workspace may be new or may have served another shape)
cudaError_t clear_group_tables(const TcPlan &plan, const Workspace &ws, cudaStream_t stream)
{
    if (!plan.tab_global) return cudaSuccess;
    return cudaMemsetAsync(ws.partials, 0,
                           (size_t)kSumParts * plan.NG * plan.K * part_pitch(plan.C) * sizeof(float),
                           stream);
}

// BMU over the tiles {tile_first + j * tile_stride}, j < ntiles.  When SN != null the per-node sums
// and counts of the visited rows are produced as well: fused into the tensor-core kernel when the
// accumulators fit in shared memory, else by the cluster-sums kernel behind it.
int bmu_tiles(const float *X, int64_t n, int C, int64_t ldX, const float *W, int K,
              int32_t *labels, int compact, int64_t tile_first, int64_t tile_stride,
              int64_t ntiles, const Workspace &ws, uint32_t flags, unsigned long long *stats,
              double *SN, cudaStream_t stream)
{
    if (ntiles <= 0) {
        if (SN) PX_CUDA(cudaMemsetAsync(SN, 0, sizeof(double) * (size_t)K * (C + 1), stream));
        return PIXIE_OK;
    }
    TcPlan plan = make_tc_plan(C, K, SN != nullptr);
    bool fused = SN != nullptr && plan.ok;
    if (!plan.ok && SN != nullptr) plan = make_tc_plan(C, K, false);
    if (!fused) {
        // the default Pixie shape has a faster kernel of its own (split tf32 operands)
        const TcPlan x3 = make_x3_plan(C, K);
        if (x3.ok) plan = x3;
    }
    const bool aligned = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && (ldX % 4 == 0) &&
                         ldX >= C && n < ((int64_t)1 << 31) - kTile;
    bool use_tc = plan.ok && aligned && !(flags & PIXIE_FLAG_FORCE_EXACT);
    CUtensorMap tm;
    if (use_tc && !make_x_tensor_map(&tm, X, n, C, ldX)) use_tc = false;
    CUtensorMap tm_tail = tm;
    if (use_tc && plan.tail8 && !make_x_tail_tensor_map(&tm_tail, X, n, C, ldX)) use_tc = false;
    // control block: norms, flags, fix-up counter, grid-barrier words
    PX_CUDA(cudaMemsetAsync(ws.aux, 0, sizeof(CodebookAux), stream));
    if (!use_tc) {
        if (flags & PIXIE_FLAG_FORCE_TC) return PIXIE_ERR_UNSUPPORTED;
        PX_CUDA(launch_bmu_exact(X, n, C, ldX, W, K, labels, tile_first, tile_stride, ntiles,
                                 compact, nullptr, nullptr, stream));
        if (stats) {
            set_u64_kernel<<<1, 1, 0, stream>>>(stats + PIXIE_STAT_KERNEL, 2ull);
            count_launch();
        }
        fused = false;
    } else {
        PX_CUDA(launch_codebook_prep(W, K, C, plan, ws.wimg, ws.aux, stream));
        TcParams p{};
        p.n = n;
        p.tile_first = tile_first;
        p.tile_stride = tile_stride;
        p.ntiles = ntiles;
        p.wimg = ws.wimg;
        p.labels = labels;
        p.compact_labels = compact;
        p.stats = stats;
        p.ctl = ws.aux;
        p.partials = fused ? ws.partials : nullptr;
        p.parts = fused ? ws.parts : nullptr;
        p.SN = fused ? SN : nullptr;
        p.delta_scale = delta_scale_from_env();
        p.dbg_flags = getenv("PIXIE_DBG_FLAGS") ? atoi(getenv("PIXIE_DBG_FLAGS")) : 0;
        p.plan = plan;
        p.tm_tail = tm_tail;
        if (fused) PX_CUDA(clear_group_tables(plan, ws, stream));
        if (plan.x3)
            PX_CUDA(launch_bmu_x3(tm, p, sum_parts(), stream));
        else
            PX_CUDA(launch_bmu_tc(tm, p, sum_parts(), stream));
        // rows the tensor-core kernel could not settle (NaN/Inf rows, degenerate codebooks): exact
        // kernel; returns immediately when the counter is zero.  In fused mode it also adds those
        // rows to SN.
        if (!fused)  // the fused (train-mode) kernel resolves those rows itself
            PX_CUDA(launch_bmu_exact(X, n, C, ldX, W, K, labels, tile_first, tile_stride, ntiles,
                                     compact, &ws.aux->fixup_count, nullptr, stream));
        if (stats) {
            set_u64_kernel<<<1, 1, 0, stream>>>(stats + PIXIE_STAT_KERNEL, 1ull);
            count_launch();
        }
    }
    if (SN != nullptr && !fused)
        PX_CUDA(launch_cluster_sums(X, n, C, ldX, labels, compact, K, tile_first, tile_stride,
                                    ntiles, ws.partials, sum_parts(), SN, ws.aux->sums_sync,
                                    stream));
    return PIXIE_OK;
}

bool bad_shape(int64_t n, int C, int64_t ldX, int K)
{
    return n < 0 || C < 1 || K < 1 || ldX < C || C > 4096 || K > 65535;
}

}  // namespace

extern "C" {

int pixie_version(void) { return 100; }

const char *pixie_error_string(int code)
{
    switch (code) {
        case PIXIE_OK: return "ok";
        case PIXIE_ERR_INVALID_ARG: return "invalid argument";
        case PIXIE_ERR_WORKSPACE: return "workspace too small";
        case PIXIE_ERR_CUDA: return g_last_error[0] ? g_last_error : "CUDA error";
        case PIXIE_ERR_UNSUPPORTED: return "shape not supported by the tensor-core kernel";
        case PIXIE_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown error";
    }
}

unsigned long long pixie_kernel_launches(void)
{
    return pixie::g_launches.load(std::memory_order_relaxed);
}

/* diagnostic builds (make prof): copies the event trace of the last launches to the host and
 * resets it; returns the number of events (0 in production builds) */
int pixie_debug_trace(unsigned long long *out_host, int max_events)
{
#ifdef PIXIE_PROFILE
    // 32 warps x 1024 slots of two words; unused slots are zero
    // ... followed by 2 x 256 CTA arrival times (max_events must leave room: 2^15 + 256)
    if (!pixie::g_trace || max_events < (1 << 15) + 256) return 0;
    if (cudaMemcpy(out_host, pixie::g_trace, (size_t)(1u << 15) * 16 + 4096, cudaMemcpyDeviceToHost) !=
        cudaSuccess)
        return -1;
    return 1 << 15;
#else
    (void)out_host;
    (void)max_events;
    return 0;
#endif
}

int pixie_plan_describe(int32_t C, int32_t K, int32_t train, int32_t *out_host)
{
    if (!out_host || C < 1 || K < 1) return PIXIE_ERR_INVALID_ARG;
    TcPlan plan = make_tc_plan(C, K, train != 0);
    if (!train) {
        const TcPlan x3 = make_x3_plan(C, K);
        if (x3.ok) plan = x3;
    }
    const int32_t v[16] = {plan.ok ? 1 : 0, plan.x3, plan.SL, plan.spc, plan.NCH, plan.NG,
                           plan.nstage, (int32_t)plan.stage_bytes, (int32_t)plan.smem_bytes,
                           (int32_t)plan.wimg_bytes, plan.tmem_cols, plan.tail8, plan.tab_global,
                           plan.Nmma, plan.Ntot, plan.ksteps};
    for (int i = 0; i < 16; ++i) out_host[i] = plan.ok ? v[i] : 0;
    return PIXIE_OK;
}

int pixie_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return PIXIE_ERR_NO_DEVICE;
    return n;
}

size_t pixie_workspace_bytes(int64_t n, int32_t C, int32_t K)
{
    if (n < 0 || C < 1 || K < 1) return 0;
    return carve(nullptr, n, C, K).total;
}

int pixie_bmu_f32(const float *X, int64_t n, int32_t C, int64_t ldX, const float *W, int32_t K,
                  int32_t *labels, double *SN_or_null, void *workspace, size_t ws_bytes,
                  uint32_t flags, unsigned long long *stats_or_null, void *stream)
{
    if (bad_shape(n, C, ldX, K) || !W || !labels || (n > 0 && !X)) return PIXIE_ERR_INVALID_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // assignment does not use the label scratch: size the check for 0 visited rows
    Workspace ws = carve(workspace, 0, C, K);
    if (!workspace || ws_bytes < ws.total) return PIXIE_ERR_WORKSPACE;
    const int64_t ntiles = (n + kTile - 1) / kTile;
    return bmu_tiles(X, n, C, ldX, W, K, labels, 0, 0, 1, ntiles, ws, flags, stats_or_null,
                     SN_or_null, st);
}

int pixie_bmu_dist_f64(const float *X, int64_t n, int32_t C, int64_t ldX, const float *W, int32_t K,
                       const int32_t *labels, double *dists, void *stream)
{
    if (bad_shape(n, C, ldX, K) || !W || !labels || !dists || (n > 0 && !X))
        return PIXIE_ERR_INVALID_ARG;
    PX_CUDA(launch_bmu_dist(X, n, C, ldX, W, K, labels, dists,
                            reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

int pixie_cluster_sums_f32(const float *X, int64_t n, int32_t C, int64_t ldX, const int32_t *labels,
                           int32_t K, double *SN, void *workspace, size_t ws_bytes, void *stream)
{
    if (bad_shape(n, C, ldX, K) || !labels || !SN || (n > 0 && !X)) return PIXIE_ERR_INVALID_ARG;
    Workspace ws = carve(workspace, 0, C, K);
    if (!workspace || ws_bytes < ws.total) return PIXIE_ERR_WORKSPACE;
    PX_CUDA(cudaMemsetAsync(ws.aux->sums_sync, 0, 2 * sizeof(unsigned int),
                            reinterpret_cast<cudaStream_t>(stream)));
    PX_CUDA(launch_cluster_sums(X, n, C, ldX, labels, 0, K, 0, 1, (n + kTile - 1) / kTile,
                                ws.partials, sum_parts(), SN, ws.aux->sums_sync,
                                reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

int pixie_columns_to_rows_f32(const double *cols, int64_t col_stride, int64_t n, int32_t C,
                              const double *divisor_or_null, float *X, int64_t ldX, void *stream)
{
    if (n < 0 || C < 1 || C > 4096 || ldX < C || col_stride < n || !X || (n > 0 && !cols))
        return PIXIE_ERR_INVALID_ARG;
    PX_CUDA(launch_columns_to_rows(cols, col_stride, n, C, divisor_or_null, X, ldX,
                                   num_sms_current_device(),
                                   reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

namespace {
struct PreWs {
    double *tmp, *rowsum;
    int32_t *flags, *pos;
    void *scan;
    size_t scan_bytes, total;
};
PreWs carve_pre(void *base, int64_t n, int C)
{
    PreWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *p = (char *)base + off;
        off += align_up(bytes, 256);
        return p;
    };
    w.tmp = (double *)take(sizeof(double) * (size_t)n * C);
    w.rowsum = (double *)take(sizeof(double) * (size_t)n);
    w.flags = (int32_t *)take(sizeof(int32_t) * (size_t)n);
    w.pos = (int32_t *)take(sizeof(int32_t) * (size_t)n);
    w.scan_bytes = pixie::preprocess_scan_bytes(n);
    w.scan = take(w.scan_bytes);
    w.total = off;
    return w;
}
}  // namespace

size_t pixie_preprocess_workspace_bytes(int32_t H, int32_t W, int32_t C)
{
    if (H < 1 || W < 1 || C < 1) return 0;
    return carve_pre(nullptr, (int64_t)H * W, C).total;
}

int pixie_preprocess_fov_f64(const void *img, int32_t H, int32_t W, int32_t C,
                             const double *norm_vect_or_null, const double *taps_host,
                             int32_t radius, double pixel_thresh_val,
                             const int32_t *seg_labels_or_null, double *blurred, double *X64_or_null,
                             float *X32_or_null, int64_t ldX32, int32_t *row_index,
                             int32_t *column_index, int32_t *labels_out_or_null, int64_t *n_kept,
                             void *workspace, size_t ws_bytes, uint32_t flags, void *stream)
{
    if (H < 1 || W < 1 || C < 1 || (int64_t)H * W >= ((int64_t)1 << 31) || !img || !blurred ||
        radius < 0 || radius > pixie::kMaxBlurRadius || (radius > 0 && !taps_host))
        return PIXIE_ERR_INVALID_ARG;
    const int blur_only = (flags & PIXIE_PREPROCESS_BLUR_ONLY) ? 1 : 0;
    if (!blur_only && (!row_index || !column_index || !n_kept || (X32_or_null && ldX32 < C)))
        return PIXIE_ERR_INVALID_ARG;
    PreWs ws = carve_pre(workspace, (int64_t)H * W, C);
    if (!workspace || ws_bytes < ws.total) return PIXIE_ERR_WORKSPACE;
    PX_CUDA(pixie::launch_preprocess(img, (flags & PIXIE_PREPROCESS_IMG_F64) ? 1 : 0, H, W, C, norm_vect_or_null, taps_host, radius,
                                     pixel_thresh_val, seg_labels_or_null, blurred, ws.tmp,
                                     ws.rowsum, ws.flags, ws.pos, ws.scan, ws.scan_bytes,
                                     X64_or_null, X32_or_null, ldX32, row_index, column_index,
                                     labels_out_or_null, n_kept, blur_only,
                                     num_sms_current_device(),
                                     reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

size_t pixie_column_quantile_workspace_bytes(int32_t C)
{
    return C < 1 ? 0 : pixie::quantile_workspace_bytes(C);
}

int pixie_column_quantile_f64(const double *X, int64_t n, int32_t C, int64_t ldX, double q,
                              double *lo, double *hi, int64_t *m, void *workspace,
                              size_t ws_bytes, void *stream)
{
    if (n < 0 || C < 1 || ldX < C || !(q >= 0.0 && q <= 1.0) || !lo || !hi || !m || (n > 0 && !X) ||
        n >= ((int64_t)1 << 32) || n * (int64_t)C >= ((int64_t)1 << 40))
        return PIXIE_ERR_INVALID_ARG;
    if (!workspace || ws_bytes < pixie::quantile_workspace_bytes(C)) return PIXIE_ERR_WORKSPACE;
    PX_CUDA(pixie::launch_column_quantile(X, n, C, ldX, q, lo, hi, m, workspace,
                                          num_sms_current_device(),
                                          reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

int pixie_label_histogram_i32(const int32_t *seg_labels, const int32_t *clusters, int64_t n,
                              int32_t n_seg, int32_t n_clusters, int32_t *counts,
                              unsigned long long *out_of_range_or_null, void *stream)
{
    if (n < 0 || n_seg < 1 || n_clusters < 1 || !counts || (n > 0 && (!seg_labels || !clusters)) ||
        ((reinterpret_cast<uintptr_t>(seg_labels) | reinterpret_cast<uintptr_t>(clusters)) & 15u))
        return PIXIE_ERR_INVALID_ARG;
    PX_CUDA(launch_label_histogram(seg_labels, clusters, n, n_seg, n_clusters, counts,
                                   out_of_range_or_null, num_sms_current_device(),
                                   reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

int pixie_scatter_labels_i16(const int32_t *row_index, const int32_t *column_index,
                             const int32_t *clusters, int64_t n, const int16_t *id_map_or_null,
                             int32_t map_len, int32_t H, int32_t W, int16_t *img,
                             int32_t *winner_ws_or_null, unsigned long long *out_of_range_or_null,
                             void *stream)
{
    if (n < 0 || n > INT32_MAX || H < 1 || W < 1 || !img || (n > 0 && (!row_index || !column_index || !clusters)) ||
        (id_map_or_null && map_len < 1))
        return PIXIE_ERR_INVALID_ARG;
    PX_CUDA(launch_scatter_labels(row_index, column_index, clusters, n, id_map_or_null, map_len, H,
                                  W, img, winner_ws_or_null, out_of_range_or_null,
                                  num_sms_current_device(),
                                  reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

int pixie_som_online_f64(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64,
                         int32_t xdim, int32_t ydim, const int64_t *sample_idx, int64_t niter,
                         double alpha0, double alpha1, double radius0, double radius1,
                         long long *iters_done_or_null, void *stream)
{
    if (xdim < 1 || ydim < 1 || !W64 || !sample_idx || niter < 1 || n < 1 || !X)
        return PIXIE_ERR_INVALID_ARG;
    const int K = xdim * ydim;
    if (bad_shape(n, C, ldX, K)) return PIXIE_ERR_INVALID_ARG;
    if (K > 1024 || C > 1024 || som_online_smem_bytes(C, K) > 227u * 1024u)
        return PIXIE_ERR_UNSUPPORTED;
    PX_CUDA(launch_som_online(X, n, C, ldX, W64, xdim, ydim, sample_idx, niter, n, alpha0, alpha1,
                              radius0, radius1, iters_done_or_null,
                              reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

int pixie_libc_sample_indices(uint32_t seed, int64_t n, int64_t count, int64_t *out_host)
{
    if (n < 1 || count < 0 || (count > 0 && !out_host)) return PIXIE_ERR_INVALID_ARG;
    srand(seed);
    for (int64_t k = 0; k < count; ++k)
        out_host[k] = (int64_t)((double)n * ((double)rand() / ((double)RAND_MAX + 1.0)));
    return PIXIE_OK;
}

int pixie_som_accum_f32(const float *X, int64_t n, int32_t C, int64_t ldX, const float *W32,
                        int32_t K, int64_t tile_first, int64_t tile_stride, double *SN,
                        void *workspace, size_t ws_bytes, uint32_t flags,
                        unsigned long long *stats_or_null, void *stream)
{
    if (bad_shape(n, C, ldX, K) || !W32 || !SN || (n > 0 && !X) || tile_first < 0 ||
        tile_stride < 1)
        return PIXIE_ERR_INVALID_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t tiles_total = (n + kTile - 1) / kTile;
    const int64_t ntiles =
        tile_first < tiles_total ? (tiles_total - tile_first + tile_stride - 1) / tile_stride : 0;
    Workspace ws = carve(workspace, ntiles * kTile, C, K);
    if (!workspace || ws_bytes < ws.total) return PIXIE_ERR_WORKSPACE;
    return bmu_tiles(X, n, C, ldX, W32, K, ws.labels_scratch, 1, tile_first, tile_stride, ntiles,
                     ws, flags, stats_or_null, SN, st);
}

int pixie_som_apply_f64(double *W64, float *W32, const double *SN, int32_t xdim, int32_t ydim,
                        int32_t C, double sigma, double alpha, void *stream)
{
    if (!W64 || !W32 || !SN || xdim < 1 || ydim < 1 || C < 1 || !(sigma > 0.0))
        return PIXIE_ERR_INVALID_ARG;
    PX_CUDA(launch_som_apply(W64, W32, SN, xdim, ydim, C, sigma, alpha,
                             reinterpret_cast<cudaStream_t>(stream)));
    return PIXIE_OK;
}

// The whole training run as one persistent launch.  PIXIE_ERR_UNSUPPORTED when the shape has no
// plan with room for the fused accumulators (callers fall back to the step-by-step path).
static int launch_whole_pass(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64,
                             float *W32, double *SN, int32_t xdim, int32_t ydim, int32_t rlen,
                             int32_t batches_per_pass, double alpha0, double alpha1, double radius0,
                             double radius1, int64_t tile_offset, int32_t world, int32_t rank,
                             const uint64_t *peer_bufs, uint32_t flag_base, void *workspace,
                             size_t ws_bytes, uint32_t flags, cudaStream_t st)
{
    const int K = xdim * ydim;
    TcPlan plan = make_tc_plan(C, K, true);
    const bool aligned = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && (ldX % 4 == 0) &&
                         n < ((int64_t)1 << 31) - kTile;
    Workspace ws = carve(workspace, 0, C, K);
    if (!workspace || ws_bytes < ws.total) return PIXIE_ERR_WORKSPACE;
    CUtensorMap tm;
    if (!plan.ok || !aligned || (flags & PIXIE_FLAG_FORCE_EXACT) || world < 1 || world > 8)
        return PIXIE_ERR_UNSUPPORTED;
    // a rank without rows (fewer tiles than ranks) still runs the kernel -- it folds, exchanges and
    // updates like the others; its tensor map points at one tile of scratch that is never loaded
    const bool empty = n == 0;
    if (!make_x_tensor_map(&tm, empty ? ws.wimg : X, empty ? kTile : n, C,
                           empty ? (int64_t)(C + 3) / 4 * 4 : ldX))
        return PIXIE_ERR_UNSUPPORTED;
    CUtensorMap tm_tail = tm;
    if (plan.tail8 && !make_x_tail_tensor_map(&tm_tail, empty ? ws.wimg : X, empty ? kTile : n, C,
                                              empty ? (int64_t)(C + 3) / 4 * 4 : ldX))
        return PIXIE_ERR_UNSUPPORTED;
    if (world > 1) {
        const int len = K * (C + 1);
        const int grid = sum_parts();
        if ((size_t)((len + grid - 1) / grid) * sizeof(double) > plan.pairs_bytes)
            return PIXIE_ERR_UNSUPPORTED;
    }
    const int64_t T = (int64_t)rlen * batches_per_pass;
    PX_CUDA(cudaMemsetAsync(ws.aux, 0, sizeof(CodebookAux), st));
    PX_CUDA(cudaMemsetAsync(SN, 0, sizeof(double) * (size_t)K * (C + 1), st));
    int rc = pixie_som_apply_f64(W64, W32, SN, xdim, ydim, C, 1.0, 0.0, st);  // W32 = fp32(W64)
    if (rc != PIXIE_OK) return rc;
    PX_CUDA(launch_codebook_prep(W32, K, C, plan, ws.wimg, ws.aux, st));
    TcParams p{};
    p.n = n;
    p.tiles_total = (n + kTile - 1) / kTile;
    p.ntiles = (p.tiles_total + batches_per_pass - 1) / batches_per_pass;  // sizes the grid
    if (world > 1) p.ntiles = sum_parts();  // every rank runs the full grid (equal barrier counts)
    p.wimg = ws.wimg;
    p.wimg_rw = ws.wimg;
    p.labels = nullptr;
    p.compact_labels = 0;
    p.stats = nullptr;
    p.ctl = ws.aux;
    p.partials = ws.partials;
    p.parts = ws.parts;
    p.SN = SN;
    p.nsteps = (int)T;
    p.apply = 1;
    p.B = batches_per_pass;
    p.t0 = 0;
    p.T = (int)T;
    p.xdim = xdim;
    p.ydim = ydim;
    p.tile_offset = tile_offset;
    p.a0 = alpha0;
    p.a1 = alpha1;
    p.r0 = radius0;
    p.r1 = radius1;
    p.W64 = W64;
    p.W32 = W32;
    p.world = world;
    p.rank = rank;
    p.flag_base = flag_base;
    p.delta_scale = delta_scale_from_env();
    p.dbg_flags = getenv("PIXIE_DBG_FLAGS") ? atoi(getenv("PIXIE_DBG_FLAGS")) : 0;
    p.dbg_step0 = getenv("PIXIE_TRACE_STEP") ? atoi(getenv("PIXIE_TRACE_STEP")) : 1 << 30;
#ifdef PIXIE_PROFILE
    if (!pixie::g_trace) {
        cudaMalloc(&pixie::g_trace, (size_t)2 * (1u << 15) * 8 + 4096);
        cudaMalloc(&pixie::g_trace_count, 4);
        cudaMemset(pixie::g_trace_count, 0, 4);
    }
    cudaMemsetAsync(pixie::g_trace, 0, (size_t)2 * (1u << 15) * 8 + 4096, st);
    p.trace = pixie::g_trace;
    p.trace_count = pixie::g_trace_count;
#endif
    for (int r = 0; r < 8; ++r)
        p.peer_buf[r] = (world > 1 && r < world) ? reinterpret_cast<double *>(peer_bufs[r]) : nullptr;
    p.plan = plan;
    p.tm_tail = tm_tail;
    PX_CUDA(clear_group_tables(plan, ws, st));
    PX_CUDA(launch_bmu_tc(tm, p, sum_parts(), st));
    return PIXIE_OK;
}

// Enqueues the T = rlen * B accumulate + apply steps on `st` (no host synchronisation).
static int enqueue_train_steps(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64,
                               float *W32, double *SN, int32_t xdim, int32_t ydim, int32_t rlen,
                               int32_t batches_per_pass, double alpha0, double alpha1,
                               double radius0, double radius1, void *workspace, size_t ws_bytes,
                               uint32_t flags, cudaStream_t st)
{
    const int K = xdim * ydim;
    const int64_t T = (int64_t)rlen * batches_per_pass;
    // W32 = fp32(W64) for the first step: an apply with nothing accumulated (SN = 0) only casts
    PX_CUDA(cudaMemsetAsync(SN, 0, sizeof(double) * (size_t)K * (C + 1), st));
    int rc = pixie_som_apply_f64(W64, W32, SN, xdim, ydim, C, 1.0, 0.0, st);
    if (rc != PIXIE_OK) return rc;
    for (int64_t t = 0; t < T; ++t) {
        const int64_t m = t % batches_per_pass;
        rc = pixie_som_accum_f32(X, n, C, ldX, W32, K, m, batches_per_pass, SN, workspace, ws_bytes,
                                 flags, nullptr, st);
        if (rc != PIXIE_OK) return rc;
        const double frac = (double)t / (double)T;
        const double r = radius0 - (radius0 - radius1) * frac;
        const double r_eff = r < 1.0 ? 0.5 : r;
        const double alpha = alpha0 - (alpha0 - alpha1) * frac;
        rc = pixie_som_apply_f64(W64, W32, SN, xdim, ydim, C, 0.5 * r_eff, alpha, st);
        if (rc != PIXIE_OK) return rc;
    }
    return PIXIE_OK;
}

namespace {
// The training pass is ~7 short launches per step; on the host that is launch-bound.  The whole
// pass is therefore captured once into a CUDA graph (on a private stream: the caller's may be the
// legacy default stream, which cannot be captured) and replayed on the caller's stream.  Executable
// graphs are cached on the full argument tuple.
struct TrainKey {
    const void *X, *W64, *W32, *SN, *ws;
    int64_t n, ldX;
    size_t ws_bytes;
    int32_t C, xdim, ydim, rlen, B;
    double a0, a1, r0, r1;
    uint32_t flags;
    int device;
    bool operator==(const TrainKey &o) const { return memcmp(this, &o, sizeof(TrainKey)) == 0; }
};
struct TrainGraph {
    TrainKey key;
    cudaGraphExec_t exec;
    unsigned long long launches;  // kernels one replay launches
    uint64_t stamp;
};
std::mutex g_graph_mutex;
std::vector<TrainGraph> g_graphs;
uint64_t g_graph_clock = 0;
cudaStream_t g_capture_stream[64] = {};
}  // namespace

int pixie_som_train_f32(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64, float *W32,
                        double *SN, int32_t xdim, int32_t ydim, int32_t rlen,
                        int32_t batches_per_pass, double alpha0, double alpha1, double radius0,
                        double radius1, void *workspace, size_t ws_bytes, uint32_t flags,
                        void *stream)
{
    if (xdim < 1 || ydim < 1 || rlen < 1 || batches_per_pass < 1 || !W64 || !W32 || !SN)
        return PIXIE_ERR_INVALID_ARG;
    const int K = xdim * ydim;
    if (bad_shape(n, C, ldX, K) || (n > 0 && !X)) return PIXIE_ERR_INVALID_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread)
        cudaStreamIsCapturing(st, &cap);
    const char *env = getenv("PIXIE_DISABLE_GRAPH");
    int dev = 0;
    PX_CUDA(cudaGetDevice(&dev));
    if ((env && env[0] == '1') || cap != cudaStreamCaptureStatusNone || dev < 0 || dev >= 64)
        return enqueue_train_steps(X, n, C, ldX, W64, W32, SN, xdim, ydim, rlen, batches_per_pass,
                                   alpha0, alpha1, radius0, radius1, workspace, ws_bytes, flags, st);

    // ---- whole-pass kernel: every step of the run inside ONE persistent launch (BMU + fused sums,
    // grid barrier, fold, batch update, codebook image rewrite) when the accumulators fit
    {
        const char *env2 = getenv("PIXIE_DISABLE_PERSISTENT");
        if (!(env2 && env2[0] == '1')) {
            int rc = launch_whole_pass(X, n, C, ldX, W64, W32, SN, xdim, ydim, rlen,
                                       batches_per_pass, alpha0, alpha1, radius0, radius1, 0, 1, 0,
                                       nullptr, 0u, workspace, ws_bytes, flags, st);
            if (rc != PIXIE_ERR_UNSUPPORTED) return rc;
        }
    }

    TrainKey key;
    memset(&key, 0, sizeof(key));
    key.X = X; key.W64 = W64; key.W32 = W32; key.SN = SN; key.ws = workspace;
    key.n = n; key.ldX = ldX; key.ws_bytes = ws_bytes;
    key.C = C; key.xdim = xdim; key.ydim = ydim; key.rlen = rlen; key.B = batches_per_pass;
    key.a0 = alpha0; key.a1 = alpha1; key.r0 = radius0; key.r1 = radius1;
    key.flags = flags; key.device = dev;

    std::lock_guard<std::mutex> lock(g_graph_mutex);
    for (auto &g : g_graphs)
        if (g.key == key) {
            g.stamp = ++g_graph_clock;
            PX_CUDA(cudaGraphLaunch(g.exec, st));
            count_launch((int)g.launches);
            return PIXIE_OK;
        }
    if (!g_capture_stream[dev])
        PX_CUDA(cudaStreamCreateWithFlags(&g_capture_stream[dev], cudaStreamNonBlocking));
    cudaStream_t cs = g_capture_stream[dev];
    const unsigned long long before = pixie::g_launches.load();
    PX_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
    int rc = enqueue_train_steps(X, n, C, ldX, W64, W32, SN, xdim, ydim, rlen, batches_per_pass,
                                 alpha0, alpha1, radius0, radius1, workspace, ws_bytes, flags, cs);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    const unsigned long long per_replay = pixie::g_launches.load() - before;
    pixie::g_launches.fetch_sub(per_replay);  // nothing ran during capture
    if (rc != PIXIE_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess) {
        set_last_cuda_error(e, "cudaStreamEndCapture");
        return PIXIE_ERR_CUDA;
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        set_last_cuda_error(e, "cudaGraphInstantiate");
        return PIXIE_ERR_CUDA;
    }
    if (g_graphs.size() >= 8) {  // evict the least recently used
        size_t victim = 0;
        for (size_t i = 1; i < g_graphs.size(); ++i)
            if (g_graphs[i].stamp < g_graphs[victim].stamp) victim = i;
        cudaGraphExecDestroy(g_graphs[victim].exec);
        g_graphs.erase(g_graphs.begin() + victim);
    }
    g_graphs.push_back(TrainGraph{key, exec, per_replay, ++g_graph_clock});
    PX_CUDA(cudaGraphLaunch(exec, st));
    count_launch((int)per_replay);
    return PIXIE_OK;
}

size_t pixie_peer_buffer_bytes(int32_t C, int32_t K)
{
    if (C < 1 || K < 1) return 0;
    // cells uint32[4] [2][8][K (C+1)]  (exchange_slice, bmu_tc_kernel.cuh)
    return (size_t)2 * 8 * K * (C + 1) * 16 + 256;
}

int pixie_som_train_peers_supported(int32_t C, int32_t K, int64_t ldX, int32_t x_aligned16)
{
    if (C < 1 || K < 1 || ldX < C || (ldX % 4) != 0 || !x_aligned16) return 0;
    const TcPlan plan = make_tc_plan(C, K, true);
    if (!plan.ok) return 0;
    const int len = K * (C + 1), grid = sum_parts();
    return (size_t)((len + grid - 1) / grid) * sizeof(double) <= plan.pairs_bytes ? 1 : 0;
}

int pixie_som_train_peers_f32(const float *X, int64_t n, int32_t C, int64_t ldX, double *W64,
                              float *W32, double *SN, int32_t xdim, int32_t ydim, int32_t rlen,
                              int32_t batches_per_pass, double alpha0, double alpha1,
                              double radius0, double radius1, int64_t tile_offset, int32_t world,
                              int32_t rank, const uint64_t *peer_bufs, uint32_t flag_base,
                              void *workspace, size_t ws_bytes, uint32_t flags, void *stream)
{
    if (xdim < 1 || ydim < 1 || rlen < 1 || batches_per_pass < 1 || !W64 || !W32 || !SN ||
        world < 2 || world > 8 || rank < 0 || rank >= world || !peer_bufs || tile_offset < 0)
        return PIXIE_ERR_INVALID_ARG;
    const int K = xdim * ydim;
    if (bad_shape(n, C, ldX, K) || (n > 0 && !X)) return PIXIE_ERR_INVALID_ARG;
    for (int r = 0; r < world; ++r)
        if (!peer_bufs[r]) return PIXIE_ERR_INVALID_ARG;
    return launch_whole_pass(X, n, C, ldX, W64, W32, SN, xdim, ydim, rlen, batches_per_pass, alpha0,
                             alpha1, radius0, radius1, tile_offset, world, rank, peer_bufs,
                             flag_base, workspace, ws_bytes, flags,
                             reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// host-buffer entry points (H2D / kernel / D2H pipelined over two streams)
// ------------------------------------------------------------------------------------------------
namespace {

// dst[r * ld + c] = (float)src[r * C + c]  (round to nearest), r < rows, c < C
__global__ void f64_to_f32_kernel(const double *__restrict__ src, float *__restrict__ dst,
                                  int64_t rows, int C, int64_t ld)
{
    const int64_t count = rows * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        dst[r * ld + c] = (float)src[i];
    }
}

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int p) : prev(p) {}
    ~DeviceGuard() { cudaSetDevice(prev); }
};

struct HostCtx {
    int device = -1;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t w_ready = nullptr;
    void *dX[2] = {nullptr, nullptr};     // fp32 chunk
    void *dX64[2] = {nullptr, nullptr};   // fp64 staging chunk (f64 entry point only)
    int32_t *dLab[2] = {nullptr, nullptr};
    double *dDist[2] = {nullptr, nullptr};
    void *dWs[2] = {nullptr, nullptr};
    float *dW = nullptr;
    double *dW64 = nullptr;
    size_t capX[2] = {0, 0}, capX64[2] = {0, 0}, capRows[2] = {0, 0}, capDist[2] = {0, 0};
    size_t capWs[2] = {0, 0}, capW = 0, capW64 = 0;
};

std::mutex g_host_mutex;
std::vector<HostCtx> g_host_ctx;

template <typename T>
cudaError_t ensure(T **ptr, size_t *cap, size_t need)
{
    if (*cap >= need) return cudaSuccess;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(ptr), need);
    if (e == cudaSuccess) *cap = need;
    return e;
}

template <typename TIn>
int map_data_host(const TIn *nodes, int32_t K, const TIn *data, int64_t n, int32_t C,
                  int32_t *labels, double *dists, int32_t device, int64_t chunk_rows)
{
    if (K < 1 || C < 1 || n < 0 || !nodes || !labels || (n > 0 && !data))
        return PIXIE_ERR_INVALID_ARG;
    if (n == 0) return PIXIE_OK;
    constexpr bool kIsF64 = sizeof(TIn) == 8;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return PIXIE_ERR_NO_DEVICE;
    int prev = 0;
    PX_CUDA(cudaGetDevice(&prev));
    if (device < 0) device = prev;
    if (device >= ndev) return PIXIE_ERR_INVALID_ARG;
    if (chunk_rows <= 0) chunk_rows = 1 << 20;
    if (chunk_rows > n) chunk_rows = n;
    chunk_rows = (chunk_rows + kTile - 1) / kTile * kTile;

    std::lock_guard<std::mutex> lock(g_host_mutex);
    DeviceGuard guard(prev);
    PX_CUDA(cudaSetDevice(device));
    HostCtx *ctx = nullptr;
    for (auto &c : g_host_ctx)
        if (c.device == device) ctx = &c;
    if (!ctx) {
        g_host_ctx.emplace_back();
        ctx = &g_host_ctx.back();
        ctx->device = device;
        for (int s = 0; s < 2; ++s)
            PX_CUDA(cudaStreamCreateWithFlags(&ctx->stream[s], cudaStreamNonBlocking));
        PX_CUDA(cudaEventCreateWithFlags(&ctx->w_ready, cudaEventDisableTiming));
    }
    const int64_t ld = (C + 3) / 4 * 4;  // device row pitch: TMA needs 16-byte multiples
    const size_t ws_need = pixie_workspace_bytes(0, C, K);
    for (int s = 0; s < 2; ++s) {
        PX_CUDA(ensure(reinterpret_cast<char **>(&ctx->dX[s]), &ctx->capX[s],
                       (size_t)chunk_rows * ld * sizeof(float)));
        if (kIsF64)
            PX_CUDA(ensure(reinterpret_cast<char **>(&ctx->dX64[s]), &ctx->capX64[s],
                           (size_t)chunk_rows * C * sizeof(double)));
        PX_CUDA(ensure(&ctx->dLab[s], &ctx->capRows[s], (size_t)chunk_rows * sizeof(int32_t)));
        if (dists)
            PX_CUDA(ensure(&ctx->dDist[s], &ctx->capDist[s], (size_t)chunk_rows * sizeof(double)));
        PX_CUDA(ensure(reinterpret_cast<char **>(&ctx->dWs[s]), &ctx->capWs[s], ws_need));
    }
    PX_CUDA(ensure(&ctx->dW, &ctx->capW, (size_t)K * C * sizeof(float)));

    // codebook -> device (fp32)
    cudaStream_t s0 = ctx->stream[0];
    if (kIsF64) {
        PX_CUDA(ensure(&ctx->dW64, &ctx->capW64, (size_t)K * C * sizeof(double)));
        PX_CUDA(cudaMemcpyAsync(ctx->dW64, nodes, (size_t)K * C * sizeof(double),
                                cudaMemcpyHostToDevice, s0));
        f64_to_f32_kernel<<<64, 256, 0, s0>>>(ctx->dW64, ctx->dW, (int64_t)K, C, (int64_t)C);
        count_launch();
    } else {
        PX_CUDA(cudaMemcpyAsync(ctx->dW, nodes, (size_t)K * C * sizeof(float),
                                cudaMemcpyHostToDevice, s0));
    }
    PX_CUDA(cudaEventRecord(ctx->w_ready, s0));
    PX_CUDA(cudaStreamWaitEvent(ctx->stream[1], ctx->w_ready, 0));

    int rc = PIXIE_OK;
    int64_t ci = 0;
    for (int64_t r0 = 0; r0 < n && rc == PIXIE_OK; r0 += chunk_rows, ++ci) {
        const int s = (int)(ci & 1);
        cudaStream_t st = ctx->stream[s];
        const int64_t rows = (n - r0) < chunk_rows ? (n - r0) : chunk_rows;
        float *dX = reinterpret_cast<float *>(ctx->dX[s]);
        if (kIsF64) {
            PX_CUDA(cudaMemcpyAsync(ctx->dX64[s], data + (size_t)r0 * C,
                                    (size_t)rows * C * sizeof(double), cudaMemcpyHostToDevice, st));
            f64_to_f32_kernel<<<148 * 8, 256, 0, st>>>(
                reinterpret_cast<const double *>(ctx->dX64[s]), dX, rows, C, ld);
            count_launch();
        } else if (ld == C) {
            PX_CUDA(cudaMemcpyAsync(dX, data + (size_t)r0 * C, (size_t)rows * C * sizeof(float),
                                    cudaMemcpyHostToDevice, st));
        } else {
            PX_CUDA(cudaMemcpy2DAsync(dX, (size_t)ld * sizeof(float), data + (size_t)r0 * C,
                                      (size_t)C * sizeof(float), (size_t)C * sizeof(float),
                                      (size_t)rows, cudaMemcpyHostToDevice, st));
        }
        rc = pixie_bmu_f32(dX, rows, C, ld, ctx->dW, K, ctx->dLab[s], nullptr, ctx->dWs[s],
                           ws_need, PIXIE_FLAG_AUTO, nullptr, st);
        if (rc != PIXIE_OK) break;
        PX_CUDA(cudaMemcpyAsync(labels + r0, ctx->dLab[s], (size_t)rows * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
        if (dists) {
            rc = pixie_bmu_dist_f64(dX, rows, C, ld, ctx->dW, K, ctx->dLab[s], ctx->dDist[s], st);
            if (rc != PIXIE_OK) break;
            PX_CUDA(cudaMemcpyAsync(dists + r0, ctx->dDist[s], (size_t)rows * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
        }
    }
    cudaError_t e0 = cudaStreamSynchronize(ctx->stream[0]);
    cudaError_t e1 = cudaStreamSynchronize(ctx->stream[1]);
    if (rc != PIXIE_OK) return rc;
    if (e0 != cudaSuccess) {
        set_last_cuda_error(e0, "stream 0");
        return PIXIE_ERR_CUDA;
    }
    if (e1 != cudaSuccess) {
        set_last_cuda_error(e1, "stream 1");
        return PIXIE_ERR_CUDA;
    }
    return PIXIE_OK;
}

}  // namespace

extern "C" {

int pixie_map_data_to_nodes_host_f32(const float *nodes, int32_t K, const float *data, int64_t n,
                                     int32_t C, int32_t *labels, double *dists_or_null,
                                     int32_t device, int64_t chunk_rows)
{
    return map_data_host<float>(nodes, K, data, n, C, labels, dists_or_null, device, chunk_rows);
}

int pixie_map_data_to_nodes_host_f64(const double *nodes, int32_t K, const double *data, int64_t n,
                                     int32_t C, int32_t *labels, double *dists_or_null,
                                     int32_t device, int64_t chunk_rows)
{
    return map_data_host<double>(nodes, K, data, n, C, labels, dists_or_null, device, chunk_rows);
}

}  // extern "C"
