"""Stage the reference package and its three hot-path test files under baseline/_ref/ (git-ignored,
NOT gpurun-ignored: it travels to the GPU box like a built .so, and never enters the history).

`pip install --target baseline/_ref /root/reference` does not work in this image: the reference
builds with hatchling, which is neither installed nor in /opt/wheelhouse.  The package is pure
Python, so what a wheel install would put there is exactly `src/ark/`; this script copies that
directory and, beside it, the reference's own tests for the Pixie SOM path:
tests/phenotyping/{cluster_helpers,pixel_som_clustering,cell_som_clustering}_test.py and the root
conftest.py.  Nothing is modified.  tests/test_reference_suite*.py run those files unchanged.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PIXIE_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
TESTS = ["cluster_helpers_test.py", "pixel_som_clustering_test.py", "cell_som_clustering_test.py"]


def stage(verbose=True):
    """Returns DST, or None when the reference tree is not on this machine."""
    if not os.path.isdir(os.path.join(REF, "src", "ark")):
        return None
    os.makedirs(DST, exist_ok=True)
    pkg = os.path.join(DST, "ark")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    shutil.copytree(os.path.join(REF, "src", "ark"), pkg,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    tdir = os.path.join(DST, "ref_tests", "phenotyping")
    os.makedirs(tdir, exist_ok=True)
    shutil.copy(os.path.join(REF, "conftest.py"), os.path.join(DST, "ref_tests", "conftest.py"))
    for t in TESTS:
        shutil.copy(os.path.join(REF, "tests", "phenotyping", t), os.path.join(tdir, t))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{REF}\n")
    if verbose:
        print("staged the reference under", DST)
    return DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
