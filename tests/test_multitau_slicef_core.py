"""The per-row routines of the CUDA kernel k_multitau_slicef (float-valued rows), run on the CPU.

xpcs-eigen_b200/csrc/multitau_slicef_core.h (on top of multitau_slice_core.h in its frame-only build) is compiled
twice: by nvcc into the kernel, and here by g++ into a small harness (tests/host_mt/mt_slicef_host.cpp) that walks
the 32 lanes of a slice one after the other with the kernel's own split of the work and its order of adding the
pieces up.  The result must agree with the oracle's multiTau2 (reference corr.cpp:315-431) within the 1e-5 relative
tolerance of BASELINE.json's north_star -- G2, IP and IF at every level, with and without the stale-tail behaviour
 -- and the pattern of exact zeros (the pairs the reference's search loses) must be identical.  No GPU
involved: this pins the arithmetic of the kernel before it reaches one; tests/test_gpu_dense.py and
tests/test_gpu_parity.py then check the kernel itself."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

import multitau_model as mm  # noqa: E402
from test_multitau_slice_core import make_rows  # noqa: E402

RTOL = 1e-5   # north_star tolerance for floating point
TYPICAL = 2e-6  # what rows of up to ~1000 events actually reach (fp64 sums against the reference's fp32 chains)


@pytest.fixture(scope="module")
def hostf(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_mtf") / "libmt_slicef_host.so")
    src = os.path.join(ROOT, "tests", "host_mt", "mt_slicef_host.cpp")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-Wno-unknown-pragmas", "-o", out, src])
    lib = C.CDLL(out)
    lib.mt_slicef_host.restype = C.c_int
    return lib


def run_slicef(lib, rows_f, rows_v, F, dpl, compat, ld_factor=4, np_=6, nps=2, nd=6, nio=2, nwarps=16, len_cap_factor=1):
    lev, tau = O.delay_schedule(F, dpl)
    nl, first, count, lo = mm.build_sched(np.asarray(lev), np.asarray(tau))
    T = len(lev)
    lastl = max([l for l in range(nl) if count[l] > 0 and l >= 1] + [0])
    cnt_last = count[lastl] if lastl >= 1 else 0
    assert lo[0] == 1
    n = np.array([len(f) for f in rows_f] + [0] * (32 - len(rows_f)), np.int32)
    ln = int(n.max())
    frames = np.zeros((max(ln, 1), 32), np.uint32)
    values = np.zeros((max(ln, 1), 32), np.float32)
    for r, (f, v) in enumerate(zip(rows_f, rows_v)):
        frames[: len(f), r] = np.asarray(f, np.uint32)
        values[: len(f), r] = np.asarray(v, np.float32)
    G2 = np.zeros((T, 32), np.float32)
    IP = np.zeros((T, 32), np.float32)
    IF = np.zeros((T, 32), np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    ld_min = nl   # as the launcher plans: first dense level of the longest slice of the job, the cap one level beyond
    for l in range(1, nl):
        if (F >> l) <= ld_factor * max(ln * len_cap_factor, 1):
            ld_min = l
            break
    ld_cap = min(nl, ld_min + 1)
    rc = lib.mt_slicef_host(dpl, int(compat), F, nl, T, count[0], lastl, cnt_last, p(n, C.c_int), p(frames, C.c_uint32),
                            p(values, C.c_float), ln, ld_factor, ld_cap, np_, nps, nd, nio, nwarps,
                            p(G2, C.c_float), p(IP, C.c_float), p(IF, C.c_float))
    assert rc == 0
    return G2, IP, IF


def run_oracle(rows_f, rows_v, F, dpl, compat):
    P = 32
    ptr = np.zeros(P + 1, np.int64)
    for r, f in enumerate(rows_f):
        ptr[r + 1] = len(f)
    ptr = np.cumsum(ptr)
    ptr[len(rows_f) + 1:] = ptr[len(rows_f)]
    t = np.concatenate([np.asarray(f, np.int32) for f in rows_f] + [np.zeros(0, np.int32)])
    v = np.concatenate([np.asarray(c, np.float32) for c in rows_v] + [np.zeros(0, np.float32)])
    return O.multitau(P, F, dpl, O.Rows(ptr, t.astype(np.int32), v.astype(np.float32)), compat=compat)


def float_values(rng, rows_c):
    """photon counts times a per-row flat-field factor times a little per-event noise (dark subtraction)"""
    out = []
    for c in rows_c:
        g = rng.uniform(0.8, 1.25)
        out.append((np.asarray(c, np.float64) * g * (1.0 + 0.1 * rng.random(len(c)))).astype(np.float32))
    return out


def check(lib, rows_f, rows_v, F, dpl, compat, rtol=RTOL, **kw):
    G2, IP, IF = run_slicef(lib, rows_f, rows_v, F, dpl, compat, **kw)
    rG2, rIP, rIF = run_oracle(rows_f, rows_v, F, dpl, compat)
    worst = 0.0
    for name, a, b in (("IP", IP, rIP), ("IF", IF, rIF), ("G2", G2, rG2)):
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        both_nan = np.isnan(a64) & np.isnan(b64)
        err = np.abs(a64 - b64)
        bad = ~(both_nan | (err <= rtol * np.abs(b64)))
        where = np.argwhere(bad)
        assert where.size == 0, "%s differs at (tau index, row) %s: %r vs %r (F=%d dpl=%d compat=%s %s)" % (
            name, where[0], a[tuple(where[0])], b[tuple(where[0])], F, dpl, compat, kw)
        assert np.array_equal(a == 0.0, b == 0.0), "%s: pattern of exact zeros differs (F=%d compat=%s)" % (name, F, compat)
        nz = b64 != 0
        if nz.any():
            worst = max(worst, float(np.nanmax(err[nz] / np.abs(b64[nz]))))
    return worst, (G2, IP, IF)


@pytest.mark.parametrize("F,dpl,occ,seed", [
    (20000, 8, 0.016, 1),    # bench workload c2: ~326 survivors per row
    (100000, 8, 0.001, 2),   # c3 with a flat-field
    (10000, 8, 0.01, 3),
    (10000, 4, 0.01, 4),
    (1500, 8, 0.2, 5),
    (500, 4, 0.9, 6),
    (33, 8, 0.5, 7),
    (1000000, 8, 0.0001, 8),
])
@pytest.mark.parametrize("compat", [True, False])
def test_uniform_slices(hostf, F, dpl, occ, seed, compat):
    rng = np.random.default_rng(seed)
    rows_f, rows_c = make_rows(rng, F, [occ * rng.uniform(0.5, 1.5) for _ in range(32)])
    worst, _ = check(hostf, rows_f, float_values(rng, rows_c), F, dpl, compat)
    assert worst < TYPICAL, worst


@pytest.mark.parametrize("F,dpl,seed", [(512, 8, 1), (4096, 8, 2), (2500, 4, 3), (33, 8, 4), (100000, 8, 5), (6000, 8, 6)])
@pytest.mark.parametrize("compat", [True, False])
def test_mixed_slices(hostf, F, dpl, seed, compat):
    """rows of every density next to each other, clustered rows (the stale-tail regime), single-event and empty
    rows, events in the dropped tail frames"""
    rng = np.random.default_rng(seed)
    scale = min(1.0, 300.0 / F)
    kinds = [0.002, 0.01, 0.03, 0.08, 0.2, 0.5, 0.9, 1.0, "cluster", 0.0, "one", "tail", "head"]
    kinds = [k if isinstance(k, str) else k * scale for k in kinds]
    kinds = (kinds * 3)[:31]
    rows_f, rows_c = make_rows(rng, F, kinds)
    if F == 512:  # the hand example of SURVEY.md A.4
        rows_f[0] = np.array(list(range(32)) + [400, 440])
        rows_c[0] = np.ones(34, np.int64)
    check(hostf, rows_f, float_values(rng, rows_c), F, dpl, compat)


@pytest.mark.parametrize("ld_factor,np_,nps,nd,nio", [(1, 1, 0, 1, 1), (2, 3, 1, 2, 3), (8, 4, 4, 6, 2), (16, 2, 0, 3, 5), (4, 8, 8, 12, 4)])
def test_work_split_stays_within_tolerance(hostf, ld_factor, np_, nps, nd, nio):
    """any first dense level, any number of pair pieces and dense pieces"""
    rng = np.random.default_rng(21)
    for F, dpl in ((100000, 8), (3000, 4), (700, 8)):
        scale = min(1.0, 300.0 / F)
        kinds = [0.3 * scale, 1.0 * scale, "cluster", 0.05 * scale, "tail", "head", 0.6 * scale, "burst"] * 4
        rows_f, rows_c = make_rows(rng, F, kinds)
        check(hostf, rows_f, float_values(rng, rows_c), F, dpl, True, ld_factor=ld_factor, np_=np_, nps=nps, nd=nd, nio=nio)


@pytest.mark.parametrize("len_cap_factor", [2, 5, 40])
def test_short_slices_of_a_long_job(hostf, len_cap_factor):
    rng = np.random.default_rng(31)
    for F, dpl in ((100000, 8), (3000, 4), (20000, 8)):
        scale = min(1.0, 100.0 / F)
        kinds = [0.3 * scale, 1.0 * scale, "cluster", 0.05 * scale, "tail", "head_small", 0.6 * scale, "burst_small"] * 4
        rows_f, rows_c = make_rows(rng, F, kinds)
        vals = float_values(rng, rows_c)
        check(hostf, rows_f, vals, F, dpl, True, len_cap_factor=len_cap_factor)
        check(hostf, rows_f, vals, F, dpl, False, len_cap_factor=len_cap_factor)


def test_stale_tail_cases_are_hit(hostf):
    """clustered rows must actually lose pairs in compat mode, otherwise the tests above prove nothing about K*"""
    rng = np.random.default_rng(5)
    for F, dpl in ((512, 8), (4096, 8), (20000, 8), (3000, 4)):
        rows_f, rows_c = make_rows(rng, F, ["burst", "head"] * 16)
        vals = float_values(rng, rows_c)
        _, (G2c, _, _) = check(hostf, rows_f, vals, F, dpl, True)
        _, (G2e, _, _) = check(hostf, rows_f, vals, F, dpl, False)
        assert (G2c != G2e).any(axis=0).sum() >= (4 if F < 10000 else 1), "F=%d: hardly any row loses pairs" % F
        check(hostf, rows_f, vals, F, dpl, True, ld_factor=1, np_=3, nps=0, nd=1)


def test_integer_valued_rows_are_exact(hostf):
    """with small integer values every sum is exact in fp32 and fp64 alike: bit for bit the oracle"""
    rng = np.random.default_rng(41)
    for F, dpl in ((20000, 8), (3000, 4)):
        rows_f, rows_c = make_rows(rng, F, [0.01 * rng.uniform(0.5, 1.5) for _ in range(32)])
        vals = [np.asarray(c, np.float32) for c in rows_c]
        for compat in (True, False):
            G2, IP, IF = run_slicef(hostf, rows_f, vals, F, dpl, compat)
            rG2, rIP, rIF = run_oracle(rows_f, vals, F, dpl, compat)
            for a, b in ((G2, rG2), (IP, rIP), (IF, rIF)):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
