"""Host build of the marching-cubes replay core (surfd_b200/csrc/mc_core.h) against the reference's outputs.
The same source is what the replay kernel runs on the GPU; here it is compiled with g++ for logic checks only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, GOLDEN
from fields import analytic_field, noise_field, MC_CASES


@pytest.fixture(scope="module")
def host_mc():
    so = os.path.join(ROOT, "tests", "native", "_mc_host.so")
    src = os.path.join(ROOT, "tests", "native", "mc_host.cpp")
    core = os.path.join(ROOT, "surfd_b200", "csrc", "mc_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "surfd_b200", "csrc"),
                               src, "-o", so])
    lib = ctypes.CDLL(so)

    def run(udf, g, fn="mc_host_run_r"):
        """fn: mc_host_run = scalar replay; mc_host_run_r = the warp-cooperative variant the kernel runs (lanes emulated)"""
        N = udf.shape[0]
        cap_v, cap_f = 600_000, 3_600_000
        v = np.empty((cap_v, 3), np.float32); f = np.empty(cap_f, np.int32)
        nv, nf = ctypes.c_int64(), ctypes.c_int64()
        st = (ctypes.c_int64 * 5)()
        P = ctypes.c_void_p
        rc = getattr(lib, fn)(P(udf.ctypes.data), P(g.ctypes.data), N, P(v.ctypes.data), ctypes.c_int64(cap_v), P(f.ctypes.data),
                             ctypes.c_int64(cap_f), ctypes.byref(nv), ctypes.byref(nf), st)
        return rc, v[:nv.value].copy(), f[:nf.value].copy(), list(st)
    return run


@pytest.mark.parametrize("case", MC_CASES, ids=lambda c: f"{c[0]}-{c[1]}-{c[2]}")
def test_core_matches_reference_golden_bit_exact(host_mc, case):
    kind, N, noise = case
    g = np.load(os.path.join(GOLDEN, "mc_fields.npz"))
    udf, grads = analytic_field(kind, N, noise, seed=N)
    key = f"{kind}_{N}_{noise}"
    for fn in ("mc_host_run", "mc_host_run_r"):
        rc, v, f, st = host_mc(udf, grads, fn)
        assert rc == 0
        assert np.array_equal(v, g[key + "_v"])      # vertex positions and numbering, bit for bit
        assert np.array_equal(f, g[key + "_f"])      # face indices and order


def test_core_matches_live_reference_on_fresh_fields(host_mc, ref_mc):
    if ref_mc is None:
        pytest.skip("oracle/_ref not built (reference absent)")
    for kind, N, noise, seed in [("sphere", 96, 0.3, 1), ("two", 72, 1.0, 2), ("hemi", 80, 0.3, 3)]:
        udf, grads = analytic_field(kind, N, noise, seed=seed)
        rv, rf = ref_mc(udf, grads)
        rc, v, f, st = host_mc(udf, grads)
        assert rc == 0 and np.array_equal(v, rv) and np.array_equal(f, rf), (kind, N, noise)


def test_core_exact_zero_udf_extension_rule(host_mc, ref_mc):
    if ref_mc is None:
        pytest.skip("oracle/_ref not built (reference absent)")
    for kind, N, noise, zf in [("sphere", 48, 0.0, 0.4), ("torus", 64, 0.3, 0.15), ("hemi", 56, 1.0, 0.4)]:
        udf, grads = analytic_field(kind, N, noise, seed=7)
        udf = udf.copy(); udf[udf < zf * 2 / (N - 1)] = 0.0
        rv, rf = ref_mc(udf, grads)
        for fn in ("mc_host_run", "mc_host_run_r"):
            rc, v, f, st = host_mc(udf, grads, fn)
            assert rc == 0 and np.array_equal(v, rv) and np.array_equal(f, rf), (kind, fn)


def test_empty_field_reports_empty_surface(host_mc):
    udf = np.full((16, 16, 16), 0.1, np.float32)
    g = np.zeros((16, 16, 16, 3), np.float32)
    rc, v, f, st = host_mc(udf, g)
    assert rc == 1 and len(v) == 0 and len(f) == 0


def test_core_dense_noise_field_matches_live_reference(host_mc, ref_mc):
    if ref_mc is None:
        pytest.skip("oracle/_ref not built (reference absent)")
    for N, seed in [(20, 1), (28, 2)]:
        udf, grads = noise_field(N, seed)
        rv, rf = ref_mc(udf, grads)
        for fn in ("mc_host_run", "mc_host_run_r"):
            rc, v, f, st = host_mc(udf, grads, fn)
            assert rc == 0 and np.array_equal(v, rv) and np.array_equal(f, rf), (N, fn, len(v), len(rv), st)
