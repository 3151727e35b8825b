"""Pins for the day the real pyFlowSOM is importable (it is not in this image: SURVEY.md section
8c).  Skipped without it.  With it, the oracle's two restatements and the device's online mode are
compared with the package itself on the same seeded inputs -- the a1 / a2 rows of SURVEY.md
section 8 stop being "parity unpinned" when these pass."""
import numpy as np
import pytest

pyFlowSOM = pytest.importorskip("pyFlowSOM")
if getattr(pyFlowSOM, "__doc__", "") and "B200" in (pyFlowSOM.__doc__ or ""):
    pytest.skip("pyFlowSOM here is this repository's stand-in, not the package",
                allow_module_level=True)
if "ark_analysis_b200" in getattr(getattr(pyFlowSOM, "som", None), "__module__", ""):
    pytest.skip("pyFlowSOM here is this repository's stand-in, not the package",
                allow_module_level=True)

import oracle  # noqa: E402


def _data(n=3000, C=12, seed=0):
    return np.random.default_rng(seed).random((n, C))


def test_oracle_map_data_to_nodes_equals_pyflowsom():
    X, W = _data(), _data(100, 12, 1)
    lab, dist = pyFlowSOM.map_data_to_nodes(W, X)
    olab, odist = oracle.map_data_to_nodes(W, X)
    np.testing.assert_array_equal(np.asarray(lab).ravel(), olab)
    np.testing.assert_array_equal(np.asarray(dist).ravel(), odist)


@pytest.mark.parametrize("rlen", [1, 3])
def test_oracle_online_som_equals_pyflowsom(rlen):
    X = _data(2000, 8, 2)
    want = np.asarray(pyFlowSOM.som(X, xdim=6, ydim=5, rlen=rlen, alpha_range=(0.05, 0.01), seed=42))
    got = oracle.som_online(X, 6, 5, rlen=rlen, alpha_range=(0.05, 0.01), seed=42)
    np.testing.assert_array_equal(want.reshape(30, 8), got)


@pytest.mark.gpu
def test_device_online_mode_equals_pyflowsom():
    from ark_analysis_b200 import som as S
    X = _data(2000, 8, 3).astype(np.float32).astype(np.float64)  # fp32-representable inputs
    want = np.asarray(pyFlowSOM.som(X, xdim=6, ydim=5, rlen=1, alpha_range=(0.05, 0.01), seed=7))
    got = S.som(X, xdim=6, ydim=5, rlen=1, alpha_range=(0.05, 0.01), seed=7, algorithm="online")
    np.testing.assert_array_equal(want.reshape(30, 8), got)
