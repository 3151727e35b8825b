"""The kernel planner (csrc/bmu_tc.cu make_tc_plan / make_x3_plan) through pixie_plan_describe: pure
host logic, no GPU.  Which kernel, how many epilogue groups / pipeline stages, which layouts -- for
BASELINE.json's shapes and over a sweep of (C, K) -- and the invariants every plan must keep
(shared memory and TMEM within what a CTA can have, stage counts the barrier scheme allows)."""
import ctypes

import numpy as np
import pytest

from ark_analysis_b200 import _native

KEYS = ["ok", "x3", "SL", "spc", "NCH", "NG", "nstage", "stage_bytes", "smem_bytes", "wimg_bytes",
        "tmem_cols", "tail8", "tab_global", "Nmma", "Ntot", "ksteps"]
SMEM_MAX = 227 * 1024


def plan(C, K, train=False):
    out = np.zeros(16, np.int32)
    rc = _native.lib().pixie_plan_describe(C, K, int(train), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return dict(zip(KEYS, (int(v) for v in out)))


def test_baseline_shapes():
    # cfg1: 16 channels, 10x10 -> the split-operand kernel, four groups, eight stages
    p = plan(16, 100)
    assert p["ok"] and p["x3"] == 1 and p["NG"] == 4 and p["nstage"] == 8 and p["SL"] * p["spc"] >= 100
    # cfg2: 32 channels -> the plain kernel (13 MMAs per tile would make the tensor pipe the bound)
    p = plan(32, 100)
    assert p["ok"] and p["x3"] == 0 and p["NG"] == 4 and p["nstage"] == 8 and p["tail8"] == 0
    # cfg3: 40 channels, 20x20 -> two chunks per tile, two groups; the tail8 layout buys six stages
    p = plan(40, 400)
    assert p["ok"] and p["NCH"] == 2 and p["NG"] == 2 and p["tail8"] == 1
    assert p["stage_bytes"] == 16384 + 4096 and p["nstage"] == 6
    # cfg4: 100 features -> tail8 again (C8 = 104)
    p = plan(100, 100)
    assert p["ok"] and p["tail8"] == 1 and p["stage_bytes"] == 3 * 16384 + 4096
    # training plans: every BASELINE shape has one (one persistent launch), none uses tail8
    for C, K, tabg in ((32, 100, 0), (40, 400, 1), (100, 100, 1), (16, 100, 0)):
        p = plan(C, K, train=True)
        assert p["ok"] and p["x3"] == 0 and p["tail8"] == 0 and p["tab_global"] == tabg, (C, K, p)


def test_split_operand_kernel_is_only_planned_where_it_fits_and_pays():
    for C in range(1, 40):
        for K in (1, 36, 64, 96, 100, 104, 105, 144):
            p = plan(C, K)
            assert p["ok"]
            assert p["x3"] == int(C <= 24 and K <= 104), (C, K)
            if p["x3"]:
                assert p["NG"] == 4 and p["NCH"] == 1 and p["nstage"] == 8 and p["tail8"] == 0
                assert p["smem_bytes"] <= 226 * 1024  # all there is beside the reserved KiB


@pytest.mark.parametrize("train", [False, True])
def test_plan_invariants_over_a_sweep(train):
    for C in list(range(1, 130, 3)) + [32, 40, 64, 72, 100, 104, 128]:
        for K in (1, 25, 64, 100, 128, 144, 256, 400, 512):
            p = plan(C, K, train)
            if not p["ok"]:
                # no room: codebook image (32-channel blocks of 128-byte rows) + one X stage per
                # group (two groups) + pair lists come close to what a CTA can have -- such shapes
                # run the exact kernel (none of BASELINE.json's does)
                nblk = (C + 31) // 32
                assert train or nblk * (K * 128 + 2 * 16384) + 20 * 1024 > 0.9 * SMEM_MAX, (C, K)
                continue
            chunk = p["SL"] * p["spc"]
            assert chunk * p["NCH"] >= K
            assert p["Nmma"] % 16 == 0 and chunk <= p["Nmma"] <= 256
            assert p["ksteps"] == (C + 7) // 8
            assert p["smem_bytes"] <= SMEM_MAX
            nbuf = 2 if p["NCH"] == 2 else p["NG"]
            assert nbuf * p["Nmma"] <= p["tmem_cols"] <= 512
            assert p["NG"] in (2, 4) and p["NCH"] in (1, 2)
            # stage it % nstage must always belong to group it % NG
            assert p["nstage"] >= p["NG"] and p["nstage"] % p["NG"] == 0 and p["nstage"] <= 8
            nblk = (C + 31) // 32
            if p["tail8"]:
                assert not train and ((C + 7) // 8 * 8) % 32 == 8 and nblk >= 2
                assert p["stage_bytes"] == (nblk - 1) * 16384 + 4096
            else:
                assert p["stage_bytes"] == nblk * 16384


def test_bad_arguments():
    out = np.zeros(16, np.int32)
    L = _native.lib()
    assert L.pixie_plan_describe(0, 100, 0, out.ctypes.data_as(ctypes.c_void_p)) != 0
    assert L.pixie_plan_describe(32, 100, 0, None) != 0
    assert plan(200, 100)["ok"] == 0 and plan(32, 600)["ok"] == 0  # outside the tensor-core range
