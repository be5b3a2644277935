"""The per-row routines of the CUDA kernel k_multitau_slice, run on the CPU.

xpcs-eigen_b200/csrc/multitau_slice_core.h is compiled twice: by nvcc into the kernel, and here by g++
into a small harness (tests/host_mt/mt_slice_host.cpp) that walks the 32 lanes of a slice one after the
other with the kernel's own split of the work.  The result must equal the oracle's multiTau2
(reference corr.cpp:315-431) bit for bit -- G2, IP and IF at every level, with and without the
stale-tail behaviour.  No GPU involved: this pins the arithmetic of the kernel before
it reaches one; tests/test_gpu_multitau_warp.py then checks the kernel itself."""
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


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_mt") / "libmt_slice_host.so")
    src = os.path.join(ROOT, "tests", "host_mt", "mt_slice_host.cpp")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-Wno-unknown-pragmas", "-o", out, src])
    lib = C.CDLL(out)
    lib.mt_slice_host.restype = C.c_int
    return lib


def run_slice(lib, rows_f, rows_c, F, dpl, compat, ld_factor=4, np_=8, nd=4, nio=2, nwarps=8, bins8=True, len_cap_factor=1):
    lev, tau = O.delay_schedule(F, dpl)
    nl, first, count, lo = mm.build_sched(np.asarray(lev), np.asarray(tau))
    T = len(lev)
    lastl = max([l for l in range(nl) if count[l] > 0 and l >= 1] + [0])
    cnt_last = count[lastl] if lastl >= 1 else 0
    assert lo[0] == 1
    n = np.array([len(f) for f in rows_f] + [0] * (32 - len(rows_f)), np.int32)
    ln = int(n.max())
    words = np.zeros((max(ln, 1), 32), np.uint32)
    for r, (f, c) in enumerate(zip(rows_f, rows_c)):
        words[: len(f), r] = (np.asarray(f, np.uint32) << 12) | np.asarray(c, np.uint32)
    G2 = np.zeros((T, 32), np.float32)
    IP = np.zeros((T, 32), np.float32)
    IF = np.zeros((T, 32), np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    # as the launcher plans: first dense level of the longest slice of the job (len_cap), the cap one level beyond it,
    # the first bin array sized for that level
    bins_rows, ld_min = 0, nl
    for l in range(1, nl):
        if (F >> l) <= ld_factor * max(ln * len_cap_factor, 1):
            ld_min = l
            bins_rows = (F >> l) if (count[l] > 0 and bins8) else 0
            break
    ld_cap = min(nl, ld_min + 1)
    rc = lib.mt_slice_host(dpl, int(compat), F, nl, T, count[0], lastl, cnt_last, p(n, C.c_int), p(words, C.c_uint32),
                           ln, ld_factor, ld_cap, np_, nd, nio, nwarps, bins_rows, p(G2, C.c_float), p(IP, C.c_float), p(IF, C.c_float))
    return rc, G2, IP, IF


def run_oracle(rows_f, rows_c, F, dpl, compat):
    P = 32
    ptr = np.zeros(P + 1, np.int64)
    for r, f in enumerate(rows_f):
        ptr[r + 1] = len(f)
    ptr = np.cumsum(ptr)
    ptr[len(rows_f) + 1:] = ptr[len(rows_f)]
    t = np.concatenate([np.asarray(f, np.int32) for f in rows_f] + [np.zeros(0, np.int32)])
    v = np.concatenate([np.asarray(c, np.float32) for c in rows_c] + [np.zeros(0, np.float32)])
    return O.multitau(P, F, dpl, O.Rows(ptr, t.astype(np.int32), v.astype(np.float32)), compat=compat)


def make_rows(rng, F, kinds):
    rows_f, rows_c = [], []
    for kind in kinds:
        if kind == "cluster":
            base = int(rng.integers(0, max(F - 40, 1)))
            f = np.unique(np.concatenate([rng.integers(0, F, 3), base + rng.integers(0, 40, 25)]))
            f = f[f < F]
        elif kind == "one":
            f = np.array([int(rng.integers(0, F))])
        elif kind in ("burst", "burst_small"):   # a dense stretch early in the series and a few late events: the rows that lose pairs
            nb = int(rng.integers(20, 200 if kind == "burst" else 100))
            base = int(rng.integers(0, max(F // 4, 1)))
            f = np.unique(np.concatenate([base + rng.integers(0, max(min(nb * 2, F - base), 1), nb),
                                          rng.integers(0, F, int(rng.integers(1, 6)))]))
            f = f[f < F]
        elif kind == "tail":    # events in the frames the deep levels drop
            f = np.unique(np.concatenate([rng.integers(0, F, 10), np.arange(max(F - 9, 0), F)]))
        elif kind in ("head", "head_small"):
            f = np.unique(np.concatenate([np.arange(min(40, F)), rng.integers(0, F, 5)]))
        else:
            f = np.nonzero(rng.random(F) < kind)[0]
        rows_f.append(f.astype(np.int64))
        rows_c.append(1 + rng.poisson(0.3, f.size))
        if kind in ("burst_small", "head_small"):
            rows_c[-1] = np.ones(f.size, np.int64)
    return rows_f, rows_c


def check(lib, rows_f, rows_c, F, dpl, compat, **kw):
    """both dense paths: the 8-bit bin arrays (taken when every row's counts sum to <= 255) and the on-the-fly walk"""
    rc = None
    for bins8 in (True, False):
        rc8, G2, IP, IF = run_slice(lib, rows_f, rows_c, F, dpl, compat, bins8=bins8, **kw)
        assert rc8 in (0, 2) and (bins8 or rc8 == 0)
        rc = rc8 if rc is None else rc
        check_against_oracle(G2, IP, IF, rows_f, rows_c, F, dpl, compat, dict(kw, bins8=bins8))
    return rc


def check_against_oracle(G2, IP, IF, rows_f, rows_c, F, dpl, compat, kw):
    rG2, rIP, rIF = run_oracle(rows_f, rows_c, F, dpl, compat)
    for name, a, b in (("IP", IP, rIP), ("IF", IF, rIF), ("G2", G2, rG2)):
        bad = np.argwhere(a.view(np.uint32) != b.view(np.uint32))
        assert bad.size == 0, "%s differs at (tau index, row) %s: %r vs %r (F=%d dpl=%d compat=%s %s)" % (
            name, bad[0], a[tuple(bad[0])], b[tuple(bad[0])], F, dpl, compat, kw)


@pytest.mark.parametrize("F,dpl,occ,seed", [
    (100000, 8, 0.001, 1),   # bench workload c3
    (10000, 8, 0.01, 2),     # c1
    (1000000, 8, 0.0001, 3),  # c5
    (10000, 4, 0.01, 4),
    (1500, 8, 0.2, 5),
    (500, 4, 0.9, 6),
    (33, 8, 0.5, 7),
    (20000, 8, 0.02, 8),
])
@pytest.mark.parametrize("compat", [True, False])
def test_uniform_slices(host, F, dpl, occ, seed, compat):
    rng = np.random.default_rng(seed)
    rows_f, rows_c = make_rows(rng, F, [occ * rng.uniform(0.5, 1.5) for _ in range(32)])
    rc = check(host, rows_f, rows_c, F, dpl, compat)
    if F * occ <= 110 and F >= 1000:
        assert rc == 2, "rows of ~100 single photons must take the 8-bit bin arrays"


@pytest.mark.parametrize("F,dpl,seed", [(512, 8, 1), (4096, 8, 2), (2500, 4, 3), (33, 8, 4), (100000, 8, 5), (6000, 8, 6)])
@pytest.mark.parametrize("compat", [True, False])
def test_mixed_slices(host, F, dpl, seed, compat):
    """rows of every density next to each other, clustered rows (the stale-tail regime), single-event and
    empty rows, events in the dropped tail frames"""
    rng = np.random.default_rng(seed)
    scale = min(1.0, 300.0 / F)
    kinds = [0.002, 0.01, 0.03, 0.08, 0.2, 0.5, 0.9, 1.0, "cluster", 0.0, "one", "tail", "head"]
    kinds = [k if isinstance(k, str) else k * scale for k in kinds]
    kinds = (kinds * 3)[:31]
    rows_f, rows_c = make_rows(rng, F, kinds)
    if F == 512:  # the hand example of SURVEY.md A.4
        rows_f[0] = np.array(list(range(32)) + [400, 440])
        rows_c[0] = np.ones(34, np.int64)
    check(host, rows_f, rows_c, F, dpl, compat)


@pytest.mark.parametrize("ld_factor,np_,nd,nio", [(1, 1, 1, 1), (2, 3, 2, 3), (8, 4, 6, 2), (16, 2, 3, 5)])
def test_work_split_does_not_matter(host, ld_factor, np_, nd, nio):
    """any first dense level, any number of pair warps and dense pieces: same bits"""
    rng = np.random.default_rng(21)
    for F, dpl in ((100000, 8), (3000, 4), (700, 8)):
        scale = min(1.0, 300.0 / F)
        kinds = [0.3 * scale, 1.0 * scale, "cluster", 0.05 * scale, "tail", "head", 0.6 * scale, "burst"] * 4
        rows_f, rows_c = make_rows(rng, F, kinds)
        check(host, rows_f, rows_c, F, dpl, True, ld_factor=ld_factor, np_=np_, nd=nd, nio=nio, nwarps=np_ + 2 + nd)


@pytest.mark.parametrize("len_cap_factor", [2, 5, 40])
def test_short_slices_of_a_long_job(host, len_cap_factor):
    """a slice much shorter than the longest one of the job starts its dense levels where the launch allows (ld_cap)"""
    rng = np.random.default_rng(31)
    for F, dpl in ((100000, 8), (3000, 4), (20000, 8)):
        scale = min(1.0, 100.0 / F)
        kinds = [0.3 * scale, 1.0 * scale, "cluster", 0.05 * scale, "tail", "head_small", 0.6 * scale, "burst_small"] * 4
        rows_f, rows_c = make_rows(rng, F, kinds)
        check(host, rows_f, rows_c, F, dpl, True, len_cap_factor=len_cap_factor)
        check(host, rows_f, rows_c, F, dpl, False, len_cap_factor=len_cap_factor)


def test_stale_tail_cases_are_hit(host):
    """clustered rows must actually lose pairs in compat mode, otherwise the tests above prove nothing about K*"""
    rng = np.random.default_rng(5)
    for F, dpl in ((512, 8), (4096, 8), (20000, 8), (3000, 4)):
        rows_f, rows_c = make_rows(rng, F, ["burst", "head"] * 16)
        _, G2c, _, _ = run_slice(host, rows_f, rows_c, F, dpl, True)
        _, G2e, _, _ = run_slice(host, rows_f, rows_c, F, dpl, False)
        assert (G2c != G2e).any(axis=0).sum() >= (4 if F < 10000 else 1), "F=%d: hardly any row loses pairs" % F
        check(host, rows_f, rows_c, F, dpl, True)
        check(host, rows_f, rows_c, F, dpl, True, ld_factor=1, np_=3, nd=1)
        # the same regime with row sums <= 255: the 8-bit bin arrays find their own first stale slots and K*
        rows_f, rows_c = make_rows(rng, F, ["burst_small", "head_small"] * 16)
        _, G2c, _, _ = run_slice(host, rows_f, rows_c, F, dpl, True)
        _, G2e, _, _ = run_slice(host, rows_f, rows_c, F, dpl, False)
        assert (G2c != G2e).any(axis=0).sum() >= (4 if F < 10000 else 1), "F=%d: hardly any row loses pairs" % F
        for ldf in (4, 2, 16):
            assert check(host, rows_f, rows_c, F, dpl, True, ld_factor=ldf) == 2


def test_heavy_rows_are_flagged(host):
    F = 6000
    f = np.sort(np.random.default_rng(9).choice(F, 40, replace=False))
    rc, _, _, _ = run_slice(host, [f], [np.full(40, 2000)], F, 8, True)
    assert rc == 1
