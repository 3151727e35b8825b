"""The reference's OWN tests for the Pixie SOM path, run UNCHANGED:
tests/phenotyping/{cluster_helpers,pixel_som_clustering,cell_som_clustering}_test.py, staged by
scripts/stage_reference.py under baseline/_ref/ (git-ignored; travels to the GPU box).  Each run is
a pytest subprocess with tests/ref_plugin.py, which only restores the environment those files
assume (module stand-ins, the reference's `--randomly-seed=24`, pandas < 2 semantics).

  CPU  : the unmodified reference modules over a `pyFlowSOM` served by the ORACLE -- pins the
         oracle (and the legacy-RNG initialisation) on every assertion the reference makes;
  GPU a: the unmodified reference modules over `pyFlowSOM` = the B200 operators;
  GPU b: the same test files against this repository's `ark.phenotyping.*` modules.
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")
FILES = ["cluster_helpers_test.py", "pixel_som_clustering_test.py", "cell_som_clustering_test.py"]

needs_ref = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF, "ref_tests", "phenotyping")),
    reason="baseline/_ref is not staged (scripts/stage_reference.py needs /root/reference)")


def run_reference_tests(backend, modules, files=FILES, extra=()):
    env = dict(os.environ)
    env["PIXIE_REF_BACKEND"] = backend
    env["PIXIE_REF_MODULES"] = modules
    env["PYTHONPATH"] = HERE + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "pytest", "-o", "addopts=", "-q", "-p", "no:cacheprovider",
           "-p", "ref_plugin", "--rootdir", os.path.join(REF, "ref_tests"), "--tb=short",
           *extra] + [os.path.join(REF, "ref_tests", "phenotyping", f) for f in files]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=ROOT, timeout=3000)
    tail = "\n".join((out.stdout + out.stderr).splitlines()[-60:])
    return out.returncode, tail


@needs_ref
def test_reference_tests_pass_unchanged_over_the_oracle():
    rc, tail = run_reference_tests("oracle", "reference")
    print(tail)
    assert rc == 0, tail
    assert " passed" in tail and "failed" not in tail.splitlines()[-1]


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("modules", ["reference", "repo"])
def test_reference_tests_pass_unchanged_on_the_b200_path(modules):
    rc, tail = run_reference_tests("b200", modules)
    print(tail)
    assert rc == 0, tail
    assert " passed" in tail and "failed" not in tail.splitlines()[-1]
