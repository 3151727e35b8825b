"""ctypes binding of libpixie_b200.so (C ABI: include/pixie_b200.h).

The library is built in tree (``ark_analysis_b200/_lib/libpixie_b200.so``) by ``build()`` /
``__graft_entry__.build()``.  There is NO fallback: if the library is missing, or a call returns an
error, this module raises -- the product path never routes through a CPU implementation.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libpixie_b200.so")
# experiments only: load another build of the same library (e.g. the `make prof` diagnostic build)
_LIB_OVERRIDE = os.environ.get("PIXIE_LIB_PATH")
CSRC = os.path.join(_HERE, "csrc")

TILE = 128
FLAG_AUTO, FLAG_FORCE_EXACT, FLAG_FORCE_TC = 0, 1, 2
NSTATS = 8
STAT_ROWS_FLAGGED, STAT_PAIRS, STAT_ROWS_FP64, STAT_ROWS_FIXUP, STAT_KERNEL = 0, 1, 2, 3, 4

# every symbol include/pixie_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "pixie_version", "pixie_error_string", "pixie_kernel_launches", "pixie_device_count", "pixie_debug_trace", "pixie_plan_describe", "pixie_workspace_bytes",
    "pixie_bmu_f32", "pixie_bmu_dist_f64", "pixie_cluster_sums_f32", "pixie_columns_to_rows_f32", "pixie_som_online_f64", "pixie_libc_sample_indices",
    "pixie_som_accum_f32", "pixie_som_apply_f64",
    "pixie_som_train_f32", "pixie_peer_buffer_bytes", "pixie_som_train_peers_supported",
    "pixie_som_train_peers_f32",
    "pixie_map_data_to_nodes_host_f32", "pixie_map_data_to_nodes_host_f64",
    "pixie_label_histogram_i32", "pixie_scatter_labels_i16",
    "pixie_preprocess_workspace_bytes", "pixie_preprocess_fov_f64",
    "pixie_column_quantile_workspace_bytes", "pixie_column_quantile_f64",
]

_lib = None
_lock = threading.Lock()


class PixieError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "pixie_b200.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        cmd = ["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", CSRC] + (["-B"] if force else [])
        out = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout)
            print(out.stderr)
        if out.returncode != 0:
            raise PixieError("nvcc build of libpixie_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded library; raises PixieError loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise PixieError(
                f"{LIB_PATH} is missing: the CUDA extension was not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback for the Pixie SOM path.")
        L = ctypes.CDLL(_LIB_OVERRIDE or LIB_PATH)
        c = ctypes
        vp, i32, i64, u32, dbl, sz = c.c_void_p, c.c_int32, c.c_int64, c.c_uint32, c.c_double, \
            c.c_size_t
        L.pixie_version.restype = c.c_int
        L.pixie_error_string.restype = c.c_char_p
        L.pixie_error_string.argtypes = [c.c_int]
        L.pixie_device_count.restype = c.c_int
        if hasattr(L, "pixie_plan_describe"):
            L.pixie_plan_describe.restype = c.c_int
            L.pixie_plan_describe.argtypes = [i32, i32, i32, vp]
        if hasattr(L, "pixie_debug_trace"):  # absent from older builds loaded via PIXIE_LIB_PATH
            L.pixie_debug_trace.restype = c.c_int
            L.pixie_debug_trace.argtypes = [vp, c.c_int]
        L.pixie_kernel_launches.restype = c.c_ulonglong
        L.pixie_workspace_bytes.restype = sz
        L.pixie_workspace_bytes.argtypes = [i64, i32, i32]
        L.pixie_bmu_f32.argtypes = [vp, i64, i32, i64, vp, i32, vp, vp, vp, sz, u32, vp, vp]
        L.pixie_bmu_dist_f64.argtypes = [vp, i64, i32, i64, vp, i32, vp, vp, vp]
        L.pixie_cluster_sums_f32.argtypes = [vp, i64, i32, i64, vp, i32, vp, vp, sz, vp]
        L.pixie_columns_to_rows_f32.argtypes = [vp, i64, i64, i32, vp, vp, i64, vp]
        L.pixie_som_online_f64.argtypes = [vp, i64, i32, i64, vp, i32, i32, vp, i64, dbl, dbl, dbl, dbl,
                                           vp, vp]
        L.pixie_libc_sample_indices.argtypes = [u32, i64, i64, vp]
        L.pixie_som_accum_f32.argtypes = [vp, i64, i32, i64, vp, i32, i64, i64, vp, vp, sz, u32,
                                          vp, vp]
        L.pixie_som_apply_f64.argtypes = [vp, vp, vp, i32, i32, i32, dbl, dbl, vp]
        L.pixie_som_train_f32.argtypes = [vp, i64, i32, i64, vp, vp, vp, i32, i32, i32, i32, dbl,
                                          dbl, dbl, dbl, vp, sz, u32, vp]
        L.pixie_peer_buffer_bytes.restype = sz
        L.pixie_peer_buffer_bytes.argtypes = [i32, i32]
        if hasattr(L, "pixie_som_train_peers_supported"):
            L.pixie_som_train_peers_supported.restype = c.c_int
            L.pixie_som_train_peers_supported.argtypes = [i32, i32, i64, i32]
        L.pixie_som_train_peers_f32.argtypes = [vp, i64, i32, i64, vp, vp, vp, i32, i32, i32, i32, dbl,
                                                dbl, dbl, dbl, i64, i32, i32, vp, u32, vp, sz, u32, vp]
        L.pixie_som_train_peers_f32.restype = c.c_int
        L.pixie_map_data_to_nodes_host_f32.argtypes = [vp, i32, vp, i64, i32, vp, vp, i32, i64]
        L.pixie_map_data_to_nodes_host_f64.argtypes = [vp, i32, vp, i64, i32, vp, vp, i32, i64]
        L.pixie_preprocess_workspace_bytes.restype = sz
        L.pixie_preprocess_workspace_bytes.argtypes = [i32, i32, i32]
        L.pixie_preprocess_fov_f64.argtypes = [vp, i32, i32, i32, vp, vp, i32, dbl, vp, vp, vp, vp,
                                               i64, vp, vp, vp, vp, vp, sz, u32, vp]
        L.pixie_preprocess_fov_f64.restype = c.c_int
        L.pixie_column_quantile_workspace_bytes.restype = sz
        L.pixie_column_quantile_workspace_bytes.argtypes = [i32]
        L.pixie_column_quantile_f64.argtypes = [vp, i64, i32, i64, dbl, vp, vp, vp, vp, sz, vp]
        L.pixie_column_quantile_f64.restype = c.c_int
        L.pixie_label_histogram_i32.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
        L.pixie_scatter_labels_i16.argtypes = [vp, vp, vp, i64, vp, i32, i32, i32, vp, vp, vp, vp]
        for name in ("pixie_label_histogram_i32", "pixie_scatter_labels_i16", "pixie_bmu_f32", "pixie_bmu_dist_f64", "pixie_cluster_sums_f32", "pixie_columns_to_rows_f32", "pixie_som_online_f64", "pixie_libc_sample_indices",
    "pixie_som_accum_f32",
                     "pixie_som_apply_f64", "pixie_som_train_f32",
                     "pixie_map_data_to_nodes_host_f32", "pixie_map_data_to_nodes_host_f64"):
            getattr(L, name).restype = c.c_int
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().pixie_error_string(rc).decode()
        raise PixieError(f"{what} failed: {msg} (code {rc})")
