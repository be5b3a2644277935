"""The per-row routines of the online multi-tau kernel k_stream_chunk, run on the CPU.

xpcs-eigen_b200/csrc/multitau_stream_core.h is compiled twice: by nvcc into the kernel, and here by g++ into a small
harness (tests/host_mt/mt_stream_host.cpp) that feeds every row chunk by chunk -- 2^k frames at a time, only the
per-row state surviving between chunks -- with the 32 lanes of every phase run one after the other.  The result must
equal the oracle's multiTau2 in its exact-maths form (reference corr.cpp:315-431 without the stale-tail behaviour of
SURVEY.md A.4, which needs the complete row and is refused in stream mode) bit for bit: G2, IP and IF at every level,
for every chunk length.  No GPU involved; tests/test_gpu_stream.py then checks the kernel itself."""
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


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_mt") / "libmt_stream_host.so")
    src = os.path.join(ROOT, "tests", "host_mt", "mt_stream_host.cpp")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-Wno-unknown-pragmas", "-o", out, src])
    lib = C.CDLL(out)
    lib.mt_stream_host.restype = C.c_int
    return lib


def sched_args(F, dpl):
    lev, tau = O.delay_schedule(F, dpl)
    nl, first, count, lo = mm.build_sched(np.asarray(lev), np.asarray(tau))
    lastl = max([l for l in range(nl) if count[l] > 0 and l >= 1] + [0])
    cnt_last = count[lastl] if lastl >= 1 else 0
    return len(lev), count[0], lastl, cnt_last


def run_stream(lib, rows_f, rows_c, F, dpl, k, ev_num=1):
    T, cnt0, lastl, cnt_last = sched_args(F, dpl)
    R = len(rows_f)
    ptr = np.zeros(R + 1, np.int64)
    ptr[1:] = np.cumsum([len(f) for f in rows_f])
    fr = np.concatenate([np.asarray(f, np.int32) for f in rows_f] + [np.zeros(0, np.int32)]).astype(np.int32)
    ct = np.concatenate([np.asarray(c, np.int32) for c in rows_c] + [np.zeros(0, np.int32)]).astype(np.int32)
    G2 = np.zeros((T, R), np.float32)
    IP = np.zeros((T, R), np.float32)
    IF = np.zeros((T, R), np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.mt_stream_host(dpl, F, T, cnt0, lastl, cnt_last, k, ev_num, R, p(ptr, C.c_int64), p(fr, C.c_int32), p(ct, C.c_int32),
                            p(G2, C.c_float), p(IP, C.c_float), p(IF, C.c_float))
    assert rc == 0
    return G2, IP, IF


def run_oracle(rows_f, rows_c, F, dpl):
    R = len(rows_f)
    ptr = np.zeros(R + 1, np.int64)
    ptr[1:] = np.cumsum([len(f) for f in rows_f])
    t = np.concatenate([np.asarray(f, np.int32) for f in rows_f] + [np.zeros(0, np.int32)])
    v = np.concatenate([np.asarray(c, np.float32) for c in rows_c] + [np.zeros(0, np.float32)])
    return O.multitau(R, F, dpl, O.Rows(ptr, t.astype(np.int32), v.astype(np.float32)), compat=False)


def check(lib, rows_f, rows_c, F, dpl, k):
    """every level by its bins (ev_num 0), the sparse levels by the row's events (1, the product), every level with 32 bins or more by events (1000)"""
    rG2, rIP, rIF = run_oracle(rows_f, rows_c, F, dpl)
    for ev_num in (1, 0, 1000):
        check_one(lib, rows_f, rows_c, F, dpl, k, ev_num, rG2, rIP, rIF)


def check_one(lib, rows_f, rows_c, F, dpl, k, ev_num, rG2, rIP, rIF):
    G2, IP, IF = run_stream(lib, rows_f, rows_c, F, dpl, k, ev_num)
    for name, a, b in (("IP", IP, rIP), ("IF", IF, rIF), ("G2", G2, rG2)):
        bad = np.argwhere(a.view(np.uint32) != b.view(np.uint32))
        assert bad.size == 0, "%s differs at (tau index, row) %s: %r vs %r (F=%d dpl=%d k=%d ev_num=%d)" % (
            name, bad[0], a[tuple(bad[0])], b[tuple(bad[0])], F, dpl, k, ev_num)


KINDS = [0.002, 0.01, 0.03, 0.08, 0.2, 0.5, 0.9, 1.0, "cluster", 0.0, "one", "tail", "head", "burst"]


@pytest.mark.parametrize("F,dpl,k,seed", [
    (512, 8, 6, 1), (512, 8, 8, 2), (4096, 8, 9, 3), (2500, 4, 7, 4), (33, 8, 6, 5), (6000, 8, 10, 6),
    (6001, 8, 8, 7), (9999, 4, 11, 8), (1000, 8, 13, 9), (20000, 8, 11, 10), (777, 4, 6, 11), (16384, 8, 12, 12),
])
def test_mixed_rows(host, F, dpl, k, seed):
    """rows of every density, clustered / single-event / empty rows, events in the frames the deep levels drop;
    chunk lengths from 64 frames to more than the whole series"""
    rng = np.random.default_rng(seed)
    scale = min(1.0, 300.0 / F)
    kinds = [q if isinstance(q, str) else q * scale for q in KINDS]
    rows_f, rows_c = make_rows(rng, F, kinds)
    check(host, rows_f, rows_c, F, dpl, k)


@pytest.mark.parametrize("k", [6, 7, 9, 11, 12])
def test_chunk_length_does_not_matter(host, k):
    rng = np.random.default_rng(77)
    for F, dpl in ((5000, 8), (3001, 4), (1 << 12, 8)):
        rows_f, rows_c = make_rows(rng, F, [0.05, 0.3, 1.0, "cluster", "tail", "head", "burst", 0.0])
        check(host, rows_f, rows_c, F, dpl, k)


@pytest.mark.parametrize("F,dpl,k,occ", [(100000, 8, 11, 0.001), (100000, 8, 12, 0.05), (1000000, 8, 12, 0.0001),
                                         (65536, 4, 10, 0.02)])
def test_bench_shapes(host, F, dpl, k, occ):
    rng = np.random.default_rng(3)
    rows_f, rows_c = make_rows(rng, F, [occ, occ * 2, "tail", "head", 0.0, "one"])
    check(host, rows_f, rows_c, F, dpl, k)


@pytest.mark.parametrize("k", [11, 12])
def test_c5_rows_at_5_percent(host, k):
    """BASELINE configs[4] at its high end: 1M frames, 5 % occupancy -- rows of ~50 000 events, 489 (245) chunks, 17 levels
    of which the top 6 (5) arrive one bin every 1..32 chunks.  G2 numerators reach 2^24 on the deep levels, where the
    reference's fp32 running sums start to round: IP / IF stay bit-exact, G2 within 1e-6."""
    F, dpl = 1000000, 8
    rng = np.random.default_rng(50 + k)
    rows_f, rows_c = make_rows(rng, F, [0.05, 0.05, 0.01, "tail"])
    rG2, rIP, rIF = run_oracle(rows_f, rows_c, F, dpl)
    G2, IP, IF = run_stream(host, rows_f, rows_c, F, dpl, k)
    assert np.array_equal(IP.view(np.uint32), rIP.view(np.uint32))
    assert np.array_equal(IF.view(np.uint32), rIF.view(np.uint32))
    exact = G2.view(np.uint32) == rG2.view(np.uint32)
    assert np.allclose(G2, rG2, rtol=1e-6, atol=0)
    assert exact[:, 2:].all(), "the sparser rows stay below 2^24 and must be bit-exact"
    assert exact.mean() > 0.5


def test_random_jobs(host):
    """seeded fuzz over frame counts (powers of two and their neighbours, tiny series, F < 2 dpl), delays per level,
    chunk lengths and row kinds"""
    rng = np.random.default_rng(2026)
    for it in range(60):
        dpl = int(rng.choice([4, 8]))
        F = int(rng.choice([rng.integers(2, 70), rng.integers(70, 5000), rng.integers(5000, 30000),
                            2 ** int(rng.integers(5, 14)) + int(rng.integers(-3, 4))]))
        k = int(rng.integers(6, 14))
        kinds = [float(rng.choice([0.001, 0.01, 0.1, 0.5, 1.0])) * min(1.0, 500.0 / F) for _ in range(2)] + ["tail", "head", "one", 0.0]
        rows_f, rows_c = make_rows(rng, F, kinds)
        check(host, rows_f, rows_c, F, dpl, k)


def test_bright_rows(host):
    """counts up to the packed word's 4095 in every frame: the 64-bit numerators and 32-bit bins must hold"""
    F, dpl, k = 3000, 8, 9
    rows_f = [np.arange(F), np.arange(0, F, 3)]
    rows_c = [np.full(F, 4095), np.full(len(rows_f[1]), 4000)]
    G2, IP, IF = run_stream(host, rows_f, rows_c, F, dpl, k)
    # beyond 2^24 the reference's fp32 running sums round: compare with exact integers instead
    lev, tau = O.delay_schedule(F, dpl)
    x = np.zeros(F, np.int64)
    x[rows_f[0]] = 4095
    for ti, (l, t) in enumerate(zip(lev, tau)):
        L = F >> l
        xl = x[: L << l].reshape(L, 1 << l).sum(axis=1)
        tp = t >> l
        num = int((xl[: L - tp] * xl[tp:]).sum())
        want = np.float32(np.float32(num) * np.float32(2.0 ** (-2 * l))) / np.float32(L - tp)
        assert G2[ti, 0] == want, (ti, G2[ti, 0], want)
        assert IP[ti, 0] == np.float32(np.float32(int(xl[: L - tp].sum())) * np.float32(2.0 ** -l)) / np.float32(L - tp)
        assert IF[ti, 0] == np.float32(np.float32(int(xl[tp:].sum())) * np.float32(2.0 ** -l)) / np.float32(L - tp)


# ---- the WARP build on the CPU: 32 threads as the lanes of a warp, barriers for __syncwarp(), random delays ----
@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_mt") / "libmt_stream_emu.so")
    src = os.path.join(ROOT, "tests", "host_mt", "mt_stream_emu.cpp")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-Wno-unknown-pragmas", "-o", out, src, "-lpthread"])
    lib = C.CDLL(out)
    lib.mt_stream_emu.restype = C.c_int
    return lib


def run_emu(lib, rows_f, rows_c, F, dpl, k, ev_num=1):
    T, cnt0, lastl, cnt_last = sched_args(F, dpl)
    R = len(rows_f)
    ptr = np.zeros(R + 1, np.int64)
    ptr[1:] = np.cumsum([len(f) for f in rows_f])
    fr = np.concatenate([np.asarray(f, np.int32) for f in rows_f] + [np.zeros(0, np.int32)]).astype(np.int32)
    ct = np.concatenate([np.asarray(c, np.int32) for c in rows_c] + [np.zeros(0, np.int32)]).astype(np.int32)
    G2 = np.zeros((T, R), np.float32)
    IP = np.zeros((T, R), np.float32)
    IF = np.zeros((T, R), np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.mt_stream_emu(dpl, F, T, cnt0, lastl, cnt_last, k, ev_num, R, p(ptr, C.c_int64), p(fr, C.c_int32), p(ct, C.c_int32),
                           p(G2, C.c_float), p(IP, C.c_float), p(IF, C.c_float))
    assert rc == 0
    return G2, IP, IF


def emu_equals_oracle(lib, case, ev_num=1):
    rows_f, rows_c, F, dpl, k, ref = case
    got = run_emu(lib, rows_f, rows_c, F, dpl, k, ev_num)
    return all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(got, ref))


def emu_cases(seed):
    rng = np.random.default_rng(seed)
    cases = []
    for F, dpl, k in ((700, 8, 6), (2100, 8, 9), (1500, 4, 7), (300, 4, 6), (5000, 8, 12)):
        rows_f, rows_c = make_rows(rng, F, [0.3, 1.0, 0.02, "cluster", "tail", "head", 0.0, "one"])
        cases.append((rows_f, rows_c, F, dpl, k, run_oracle(rows_f, rows_c, F, dpl)))
    return cases


def test_warp_build_equals_the_oracle(emu):
    """the device's own control flow -- one lane per thread, 64-bit warp sums, owner-lane stores, barriers where the
    kernel has __syncwarp() -- under random delays of the lanes, several times over"""
    emu.mt_stream_emu_drop(0)
    for case in emu_cases(11):
        for ev_num in (1, 0, 1000):   # the product's choice between the two walks, bins only, events wherever possible
            assert emu_equals_oracle(emu, case, ev_num), "F=%d dpl=%d k=%d ev_num=%d" % (case[2:5] + (ev_num,))


def test_warp_emulation_notices_a_missing_barrier(emu):
    """the check above is only worth something if a missing __syncwarp() shows: drop the barrier of one source line
    of the core at a time (for all lanes alike) and the result must go wrong for most of them (the few that survive
    are barriers followed by another one before anything they order is touched)"""
    cases = emu_cases(12)[:2]
    emu.mt_stream_emu_drop(0)
    assert all(emu_equals_oracle(emu, c) for c in cases)
    lines = (C.c_int * 64)()
    n = emu.mt_stream_emu_sites(lines, 64)
    sites = sorted(lines[i] for i in range(n))
    assert len(sites) >= 8, sites
    caught = []
    try:
        for ln in sites:
            emu.mt_stream_emu_drop(ln)
            if not all(emu_equals_oracle(emu, c) for c in cases for _ in range(2)):
                caught.append(ln)
    finally:
        emu.mt_stream_emu_drop(0)
    assert len(caught) >= 5, "barriers whose removal was noticed: %s of %s" % (caught, sites)
    assert all(emu_equals_oracle(emu, c) for c in cases)


def test_warp_build_under_thread_sanitizer(tmp_path):
    """the same build under ThreadSanitizer where the toolchain has it (it does not see every race of this pattern --
    32 threads, four shadow slots per word -- which is why the two tests above exist)"""
    exe = str(tmp_path / "mt_stream_emu_tsan")
    src = os.path.join(ROOT, "tests", "host_mt", "mt_stream_emu.cpp")
    p = subprocess.run(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fsanitize=thread",
                        "-DST_EMU_MAIN", "-Wno-unknown-pragmas", "-o", exe, src, "-lpthread"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        pytest.skip("no ThreadSanitizer build here: " + p.stdout[-300:])
    F, dpl, k = 700, 8, 6
    rng = np.random.default_rng(5)
    rows_f, rows_c = make_rows(rng, F, [0.3, 1.0, 0.02, "cluster", "tail", "head", 0.0, "one"])
    T, cnt0, lastl, cnt_last = sched_args(F, dpl)
    R = len(rows_f)
    ptr = np.zeros(R + 1, np.int64)
    ptr[1:] = np.cumsum([len(f) for f in rows_f])
    fr = np.concatenate([np.asarray(f, np.int32) for f in rows_f]).astype(np.int32)
    ct = np.concatenate([np.asarray(c, np.int32) for c in rows_c]).astype(np.int32)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([dpl, F, T, cnt0, lastl, cnt_last, k, 1, R], np.int32).tobytes())
        f.write(ptr.tobytes())
        f.write(fr.tobytes())
        f.write(ct.tobytes())
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    p = subprocess.run([exe, fin, fout], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    if p.returncode != 0 and "unexpected memory mapping" in p.stdout:
        pytest.skip("ThreadSanitizer cannot map its shadow here")
    assert p.returncode == 0, "66 = data race reported by ThreadSanitizer:\n" + p.stdout[-3000:]
    out = np.fromfile(fout, np.float32).reshape(3, T, R)
    ref = run_oracle(rows_f, rows_c, F, dpl)
    for a, b in zip(out, ref):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
