"""pytest plugin (`-p ref_plugin`) for running the REFERENCE's own test files, unchanged, from
baseline/_ref/ref_tests (scripts/stage_reference.py) in an environment the reference was not
written for.  It changes nothing in those files; it only restores the environment they assume:

* module stand-ins for feather / natsort / alpineer / skimage.io on sys.path
  (ark_analysis_b200/compat/shims; a really installed package wins);
* `pyFlowSOM`: PIXIE_REF_BACKEND=b200 -> the B200 operators (compat shim), =oracle -> the CPU
  oracle (tests/ref_shims_cpu; lets the CPU suite pin the oracle on the reference's assertions);
* PIXIE_REF_MODULES=reference -> `ark.phenotyping.*` are the UNMODIFIED reference modules
  (baseline/_ref/ark); =repo -> they are this repository's modules (compat.install(force=True));
* the reference's pytest addopts include `--randomly-seed=24` (pyproject.toml:106-121):
  pytest-randomly is not installed, so `random` / `numpy.random` are re-seeded with 24 before every
  test, which is what that plugin does;
* pandas: the reference pins pandas < 2.  pandas 3 (installed here) infers Arrow string columns
  (breaks `df[df2.columns.values] = ...`, cluster_helpers.py:244) and refuses to upcast an int64
  column on `df.loc[rows, col] = 'fov0'` (the tests' own fixtures, cluster_helpers_test.py:164).
  `future.infer_string` is switched off and the legacy upcast is restored for `.loc` assignment.
"""
import os
import random
import sys

import numpy as np
import pandas as pd
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BACKEND = os.environ.get("PIXIE_REF_BACKEND", "b200")
MODULES = os.environ.get("PIXIE_REF_MODULES", "reference")

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if BACKEND == "oracle":
    sys.path.insert(0, os.path.join(HERE, "ref_shims_cpu"))
from ark_analysis_b200 import compat  # noqa: E402

sys.path.append(compat.shim_path())
if MODULES == "reference":
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
else:
    compat.install(force=True)
    # The reference's cluster_helpers.py also holds the consensus (meta) clustering classes, which
    # are outside the SOM path and not mirrored here; its test file imports them at module level.
    # They are taken from the UNMODIFIED reference module, loaded under a private name.
    import importlib.util
    from ark_analysis_b200 import cluster_helpers as _ours
    _spec = importlib.util.spec_from_file_location(
        "_reference_cluster_helpers",
        os.path.join(ROOT, "baseline", "_ref", "ark", "phenotyping", "cluster_helpers.py"))
    _refmod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_refmod)
    for _name in ("PixieConsensusCluster", "verify_unique_meta_clusters"):
        if not hasattr(_ours, _name):
            setattr(_ours, _name, getattr(_refmod, _name))
    if BACKEND == "oracle":
        import pyFlowSOM as _cpu  # tests/ref_shims_cpu
        from ark_analysis_b200 import som as _som
        _som.som = lambda data, xdim=10, ydim=10, rlen=10, alpha_range=(0.05, 0.01), seed=None, \
            **kw: _cpu.som(data, xdim, ydim, rlen, alpha_range, seed=seed)
        _som.map_data_to_nodes = lambda nodes, newdata, **kw: _cpu.map_data_to_nodes(nodes, newdata)

try:
    pd.set_option("future.infer_string", False)
except Exception:  # noqa: BLE001 -- option absent in older pandas: nothing to switch off
    pass

_loc_setitem = pd.core.indexing._LocIndexer.__setitem__


def _legacy_loc_setitem(self, key, value):
    try:
        return _loc_setitem(self, key, value)
    except TypeError as exc:
        if "Invalid value" not in str(exc) or not isinstance(key, tuple) or len(key) != 2:
            raise
        col = key[1]
        if not isinstance(col, str) or col not in self.obj.columns:
            raise
        self.obj[col] = self.obj[col].astype(object)  # what pandas < 2 did silently
        return _loc_setitem(self, key, value)


pd.core.indexing._LocIndexer.__setitem__ = _legacy_loc_setitem


@pytest.fixture(autouse=True)
def _randomly_seed_24():
    random.seed(24)
    np.random.seed(24)
    yield
