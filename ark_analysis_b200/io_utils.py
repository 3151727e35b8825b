"""The handful of file helpers the Pixie SOM drivers use.

The reference takes these from third-party packages that are not part of its tree (``alpineer``
``io_utils`` / ``misc_utils`` and ``feather``; call sites listed in SURVEY.md Appendix C).  They are
re-provided here with the behaviour those call sites and the reference's tests rely on.
"""
import os
import pathlib

import pandas as pd
import pyarrow.feather as _paf
from pyarrow.lib import ArrowInvalid  # noqa: F401  (re-exported: the corruption path catches it)


# ------------------------------------------------------------------------------------------------
# feather (reference: `import feather`; read at cluster_helpers.py:78, :205, :213, written at :116)
# ------------------------------------------------------------------------------------------------
def read_dataframe(path, columns=None) -> pd.DataFrame:
    """Read a Feather (Arrow IPC) file.  Raises ``pyarrow.lib.ArrowInvalid`` or ``OSError`` on a
    corrupted file, which is what pixel_som_clustering.py:120 catches."""
    return _paf.read_feather(str(path), columns=columns)


def write_dataframe(df: pd.DataFrame, path, compression="uncompressed"):
    _paf.write_feather(df, str(path), compression=compression)


def read_table(path, columns=None):
    """Arrow table view of a Feather file (zero-copy column buffers for the device upload)."""
    return _paf.read_table(str(path), columns=columns)


def write_table(table, path, compression="uncompressed"):
    """Write an Arrow table as Feather.  Stale pandas metadata (it may name columns that were
    dropped or lack the ones appended since) is removed: the Arrow types carry everything
    ``read_dataframe`` needs."""
    _paf.write_feather(table.replace_schema_metadata(None), str(path), compression=compression)


# ------------------------------------------------------------------------------------------------
# alpineer.io_utils
# ------------------------------------------------------------------------------------------------
def validate_paths(paths):
    """Raise FileNotFoundError if any path (one path or a list) does not exist
    (pinned by tests/phenotyping/pixel_som_clustering_test.py:96-100 of the reference)."""
    if isinstance(paths, (str, pathlib.Path)):
        paths = [paths]
    for p in paths:
        if not os.path.exists(p):
            raise FileNotFoundError(
                f"A bad path, {p}, was provided: the file or directory does not exist.")


def list_files(dir_name, substrs=None, exact_match=False, ignore_hidden=True):
    """Names (not paths) of the files in ``dir_name`` whose name contains one of ``substrs``.
    Sorted, so that every run and every rank sees the same order (the reference returns directory
    order; the divergence is deliberate and noted in DESIGN.md)."""
    names = [f for f in os.listdir(dir_name) if os.path.isfile(os.path.join(dir_name, f))]
    if ignore_hidden:
        names = [f for f in names if not f.startswith(".")]
    if substrs is not None:
        if isinstance(substrs, str):
            substrs = [substrs]
        if exact_match:
            names = [f for f in names if any(os.path.splitext(f)[0] == s for s in substrs)]
        else:
            names = [f for f in names if any(s in f for s in substrs)]
    return sorted(names)


def remove_file_extensions(files):
    if files is None:
        return None
    return [os.path.splitext(f)[0] for f in files]


# ------------------------------------------------------------------------------------------------
# alpineer.misc_utils
# ------------------------------------------------------------------------------------------------
def _as_list(v):
    if isinstance(v, (str, bytes)) or not hasattr(v, "__iter__"):
        return [v]
    return list(v)


def verify_in_list(warn=False, **kwargs):
    """``verify_in_list(a=..., b=...)``: ValueError unless every element of the first keyword
    argument is in the second (call sites: pixel_som_clustering.py:70, :75; cluster_helpers.py:140)."""
    if len(kwargs) != 2:
        raise ValueError("verify_in_list expects exactly two keyword arguments")
    (name_a, a), (name_b, b) = kwargs.items()
    a, b = _as_list(a), _as_list(b)
    pool = set(b)
    missing = [x for x in a if x not in pool]
    if missing:
        shown = ", ".join(str(x) for x in missing[:10])
        msg = (f"Not all values given in list {name_a} were found in list {name_b}.\n"
               f" Invalid values (first 10): {shown}")
        if warn:
            import warnings
            warnings.warn(msg)
            return False
        raise ValueError(msg)
    return True


def verify_same_elements(enforce_order=False, warn=False, **kwargs):
    """ValueError if the two keyword collections differ as sets; with ``enforce_order`` also if the
    order differs (call site: pixel_som_clustering.py:206-217)."""
    if len(kwargs) != 2:
        raise ValueError("verify_same_elements expects exactly two keyword arguments")
    (name_a, a), (name_b, b) = kwargs.items()
    a, b = _as_list(a), _as_list(b)
    if set(a) != set(b):
        only_a = [x for x in a if x not in set(b)][:10]
        only_b = [x for x in b if x not in set(a)][:10]
        msg = (f"Lists {name_a} and {name_b} are not the same: only in {name_a}: {only_a}; "
               f"only in {name_b}: {only_b}")
        if warn:
            import warnings
            warnings.warn(msg)
            return False
        raise ValueError(msg)
    if enforce_order and a != b:
        first = next(i for i, (x, y) in enumerate(zip(a, b)) if x != y)
        msg = (f"Lists {name_a} and {name_b} ordered differently: first mismatch at index {first} "
               f"({a[first]} vs {b[first]})")
        if warn:
            import warnings
            warnings.warn(msg)
            return False
        raise ValueError(msg)
    return True


def make_blank_file(folder, name):
    """test helper (alpineer.test_utils._make_blank_file)"""
    pathlib.Path(os.path.join(folder, name)).touch()
