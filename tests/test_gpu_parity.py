"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on identical seeded
inputs.  Bars (BASELINE.json north_star): integer photon-count accumulations bit-exact --
here that covers G2, IP, IF at every level, frameSum, pixelSum, partition means and
norm-0-g2; floating-point paths (flat-field, averaging, frame-sum normalisation, stderr)
within 1e-5 relative."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance for floating-point results


def run_gpu(pkg, dq, sq, F, off, idx, val, want_g2=True, **kw):
    c = pkg.Correlator(dq, sq, F, **kw)
    c.push_sparse(idx, val, off)
    sums = c.finish_ingest()
    G = c.multitau(want=want_g2)
    g2, se = c.normalize()
    info = c.info()
    c.close()
    return sums, G, g2, se, info


def run_oracle(O, dq, sq, F, off, idx, val, dpl=8, compat=True, flat=None, stride=1, avg=1, swindow=None,
               normalize_by_framesum=False):
    qm = O.QMap(dq, sq)
    swindow = swindow or max(1, F // 10)
    fo = O.sparse_filter(qm, F, off, idx, val, flat=flat, stride=stride, avg=avg, swindow=swindow)
    sums_raw = dict(frame_sum=fo.frame_sum.copy())
    O.post_scale(qm, F, swindow, fo, normalize_by_framesum=normalize_by_framesum)
    G2, IP, IF = O.multitau(qm.P, F, dpl, fo.rows, compat=compat)
    g2, se = O.normalize(qm, G2, IP, IF)
    W = F // swindow
    sums = dict(pixel_sum=fo.pixel_sum, frame_sum=sums_raw["frame_sum"], part_total=fo.part_total[: qm.S],
                part_partial=fo.part_partial[: W * qm.S].reshape(W, qm.S))
    return sums, (G2, IP, IF), g2, se


def assert_exact(a, b, what):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), "%s: %d of %d entries differ, first at %d: %r vs %r" % (
        what, bad.sum(), a.size, np.argmax(bad), a[np.argmax(bad)], b[np.argmax(bad)])


def assert_close(a, b, what, rtol=RTOL, atol=0.0):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    both_nan = np.isnan(a) & np.isnan(b)
    err = np.abs(a - b)
    bad = ~(both_nan | (err <= atol + rtol * np.abs(b)))
    assert not bad.any(), "%s: %d of %d beyond rtol %g, worst rel %.3g" % (
        what, bad.sum(), a.size, rtol, np.nanmax(err / np.maximum(np.abs(b), 1e-300)))


def compare_exact(got, ref):
    """(sums, (G2, IP, IF), g2, se[, info]) of the GPU against the oracle's: integer path, everything but the
    std-error bit for bit."""
    sums, G, g2, se = got[:4]
    rsums, rG, rg2, rse = ref
    for k in ("pixel_sum", "frame_sum", "part_total", "part_partial"):
        assert_exact(sums[k], rsums[k], k)
    for k, nm in enumerate(("G2", "IP", "IF")):
        assert_exact(G[k], rG[k], nm)
    assert_exact(g2, rg2, "norm-0-g2")
    ok = np.isfinite(rse)
    assert_close(se[ok], rse[ok], "norm-0-stderr")
    if len(got) > 4:
        assert got[4].value_kind == 0


CASES = [  # h, w, F, occupancy, seed, dpl
    (32, 32, 100, 0.05, 1, 8),
    (48, 40, 601, 0.02, 2, 8),     # odd frame count: last frame dropped at level 1
    (64, 64, 1000, 0.01, 3, 8),
    (64, 64, 1023, 0.03, 4, 4),
    (40, 56, 2049, 0.004, 5, 8),
    (96, 96, 4000, 0.002, 6, 8),   # sparse rows: the stale-tail regime of SURVEY.md A.4
    (24, 24, 33, 0.3, 7, 8),       # dense rows, F barely above 2*dpl
    (16, 16, 15, 0.5, 8, 8),       # F < 2*dpl: level 0 only, truncated
]


@pytest.mark.parametrize("h,w,F,occ,seed,dpl", CASES)
def test_sparse_integer_path_bit_exact(pkg, oracle, h, w, F, occ, seed, dpl):
    dq, sq, off, idx, val = make_case(pkg, h, w, F, occ, seed)
    sums, G, g2, se, info = run_gpu(pkg, dq, sq, F, off, idx, val, dpl=dpl, compat=True)
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=True)
    assert info.value_kind == 0
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_exact(G[k], rG[k], name)
    assert_exact(sums["frame_sum"], rs["frame_sum"], "frameSum")
    assert_exact(sums["pixel_sum"], rs["pixel_sum"], "pixelSum")
    assert_exact(sums["part_total"], rs["part_total"], "partition-mean-total")
    assert_exact(sums["part_partial"], rs["part_partial"], "partition-mean-partial")
    assert_exact(g2, rg2, "norm-0-g2")
    assert_close(se, rse, "norm-0-stderr")


@pytest.mark.parametrize("h,w,F,occ,seed,dpl", CASES[:6])
def test_sparse_exact_sums_without_compat(pkg, oracle, h, w, F, occ, seed, dpl):
    """compat off = the mathematically exact pair sums (oracle restricted to the live prefix)."""
    dq, sq, off, idx, val = make_case(pkg, h, w, F, occ, seed)
    _, G, _, _, _ = run_gpu(pkg, dq, sq, F, off, idx, val, dpl=dpl, compat=False)
    _, rG, _, _ = run_oracle(oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=False)
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_exact(G[k], rG[k], name)


def test_stale_tail_fixture(pkg, oracle):
    """The hand example of SURVEY.md A.4, verified against the reference binary: one pixel with
    unit counts at frames 0..31, 400, 440, F=512: the reference drops the (100,110) pair at
    level 2, so G2[tau=40] = 0 while the exact value is (1/4*1/4)/118."""
    F = 512
    dq = np.ones((1, 2), np.int32)
    sq = np.ones((1, 2), np.int32)
    frames = list(range(32)) + [400, 440]
    off = np.zeros(F + 1, np.int64)
    for f in frames:
        off[f + 1:] += 1
    idx = np.zeros(len(frames), np.int32)
    val = np.ones(len(frames), np.int16)
    _, tv = oracle.delay_schedule(F, 8)
    k = list(tv).index(40)
    _, G, _, _, _ = run_gpu(pkg, dq, sq, F, off, idx, val, compat=True)
    assert G[0][k, 0] == 0.0
    assert G[1][k, 0] == np.float32(8.5) / np.float32(118)
    _, Ge, _, _, _ = run_gpu(pkg, dq, sq, F, off, idx, val, compat=False)
    assert Ge[0][k, 0] == np.float32(0.0625) / np.float32(118)
    _, rG, _, _ = run_oracle(oracle, dq, sq, F, off, idx, val, compat=True)
    assert_exact(G[0], rG[0], "G2")


def test_adversarial_rows(pkg, oracle):
    """Single-event rows, a duplicated pixel inside one frame, events in the dropped last frame,
    an all-masked column, unsorted pixel order within a frame, an empty frame range."""
    h, w, F = 8, 8, 257
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=2, static_per_dynamic=2, r_min=0.0, r_max=6.0)
    dq[:, 0] = 0
    rng = np.random.default_rng(5)
    per_frame = []
    for f in range(F):
        if 100 <= f < 120:
            per_frame.append((np.zeros(0, np.int32), np.zeros(0, np.int16)))
            continue
        n = rng.integers(0, 6)
        p = rng.integers(0, h * w, n).astype(np.int32)
        if f % 17 == 0 and n > 0:
            p = np.concatenate([p, p[:1]])  # same pixel twice in one frame
        v = rng.integers(1, 4, p.size).astype(np.int16)
        per_frame.append((p, v))
    per_frame[F - 1] = (np.arange(h * w, dtype=np.int32)[::-1].copy(), np.ones(h * w, np.int16))
    per_frame[3] = (np.array([9], np.int32), np.array([2], np.int16))
    off = np.concatenate([[0], np.cumsum([p.size for p, _ in per_frame])]).astype(np.int64)
    idx = np.concatenate([p for p, _ in per_frame])
    val = np.concatenate([v for _, v in per_frame])
    sums, G, g2, se, _ = run_gpu(pkg, dq, sq, F, off, idx, val)
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val)
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_exact(G[k], rG[k], name)
    assert_exact(sums["frame_sum"], rs["frame_sum"], "frameSum")
    assert_exact(sums["pixel_sum"], rs["pixel_sum"], "pixelSum")
    assert_exact(g2, rg2, "norm-0-g2")
    assert_close(se, rse, "norm-0-stderr")


def test_empty_input(pkg, oracle):
    h, w, F = 8, 8, 64
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=2, static_per_dynamic=2, r_min=0.0, r_max=6.0)
    off = np.zeros(F + 1, np.int64)
    idx = np.zeros(0, np.int32)
    val = np.zeros(0, np.int16)
    sums, G, g2, se, _ = run_gpu(pkg, dq, sq, F, off, idx, val)
    assert not G[0].any() and not G[1].any() and not G[2].any()
    assert not sums["pixel_sum"].any()
    assert np.isnan(g2).all()  # 0/0 in every static bin, as the reference produces


def test_large_counts_fall_back_to_float_path(pkg, oracle):
    """Counts beyond the 12-bit packed field switch the store to float words; still exact here
    because every sum stays below 2^24."""
    dq, sq, off, idx, val = make_case(pkg, 32, 32, 300, 0.03, 21)
    val = (val.astype(np.int32) * 3000).clip(0, 32767).astype(np.int16)
    sums, G, g2, se, info = run_gpu(pkg, dq, sq, 300, off, idx, val)
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, 300, off, idx, val)
    assert info.value_kind == 1
    assert_close(G[0], rG[0], "G2")
    assert_close(G[1], rG[1], "IP")
    assert_close(G[2], rG[2], "IF")
    assert_close(g2, rg2, "norm-0-g2")


@pytest.mark.parametrize("stride,avg", [(1, 1), (2, 1), (1, 2), (2, 2), (1, 3)])
def test_flatfield_stride_average(pkg, oracle, stride, avg):
    h, w, F_raw = 40, 40, 1200
    block = stride * avg if (stride > 1 and avg > 1) else max(stride, avg)
    F = F_raw // block
    dq, sq, off, idx, val = make_case(pkg, h, w, F_raw, 0.02, 31)
    flat = pkg.synth.flatfield(h * w)
    sw = max(1, F // 10)
    sums, G, g2, se, info = run_gpu(pkg, dq, sq, F, off, idx, val, flatfield=flat, stride=stride, avg=avg,
                                    static_window=sw)
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val, flat=flat, stride=stride, avg=avg, swindow=sw)
    assert info.value_kind == 1
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_close(G[k], rG[k], name)
    assert_close(sums["frame_sum"], rs["frame_sum"], "frameSum")
    assert_close(sums["pixel_sum"], rs["pixel_sum"], "pixelSum")
    assert_close(sums["part_total"], rs["part_total"], "partition-mean-total")
    assert_close(sums["part_partial"], rs["part_partial"], "partition-mean-partial")
    assert_close(g2, rg2, "norm-0-g2")
    assert_close(se, rse, "norm-0-stderr")


def test_normalize_by_framesum(pkg, oracle):
    h, w, F = 40, 40, 800
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.03, 41)
    sums, G, g2, se, info = run_gpu(pkg, dq, sq, F, off, idx, val, normalize_by_framesum=True)
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val, normalize_by_framesum=True)
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_close(G[k], rG[k], name)
    assert_close(g2, rg2, "norm-0-g2")
    assert_close(se, rse, "norm-0-stderr")


def test_pushes_in_chunks_equal_one_push(pkg):
    dq, sq, off, idx, val = make_case(pkg, 48, 48, 900, 0.02, 51)
    F = 900
    _, G1, g1, s1, _ = run_gpu(pkg, dq, sq, F, off, idx, val)
    c = pkg.Correlator(dq, sq, F)
    for a in range(0, F, 250):
        b = min(F, a + 250)
        c.push_sparse(idx[off[a]: off[b]], val[off[a]: off[b]], off[a: b + 1] - off[a])
    c.finish_ingest(want=False)
    G2 = c.multitau()
    g2, s2 = c.normalize()
    # a second ingest on the same handle
    c.reset()
    c.push_sparse(idx, val, off)
    c.finish_ingest(want=False)
    G3 = c.multitau()
    c.close()
    for k in range(3):
        assert_exact(G1[k], G2[k], "chunked push")
        assert_exact(G1[k], G3[k], "second ingest")
    assert_exact(g1, g2, "g2 chunked")


def test_shards_sum_to_single_gpu_partials(pkg):
    """Static-bin-aligned sharding: the element-wise sum of the shards' partial
    buffers is the single-shard buffer bit for bit, so g2/stderr do not depend on the GPU
    count.  Both shards run on cuda:0 here; the cross-rank SUM is covered by the gloo test."""
    import torch
    dq, sq, off, idx, val = make_case(pkg, 64, 64, 700, 0.02, 61, n_dynamic=5, static_per_dynamic=4)
    F = 700

    def partials(k, K):
        c = pkg.Correlator(dq, sq, F, shard_index=k, shard_count=K)
        c.push_sparse(idx, val, off)
        c.finish_ingest(want=False)
        c.multitau(want=False)
        ptr, n = c.normalize_partials()
        torch.cuda.synchronize()
        host = pkg.torchio.device_view(ptr, n, "float64", "cuda:0").cpu().numpy().copy()
        g2, se = c.normalize_finish()
        rows = c.info().n_rows
        c.close()
        return host, g2, se, rows

    full, g2, se, rows = partials(0, 1)
    parts = [partials(k, 3) for k in range(3)]
    assert sum(p[3] for p in parts) == rows
    total = parts[0][0] + parts[1][0] + parts[2][0]
    assert_exact(total, full, "sum of shard partials")


@pytest.mark.parametrize("flat", [False, True])
def test_frame_dump_matches_filter_rows(pkg, oracle, flat):
    """xpcs_get_frames (the --frameout dump, main.cpp:276-310) against the oracle's filtered rows: integer
    store exactly, float store (flat-field) within 1e-5."""
    h, w, F, N = 40, 48, 300, 25
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.03, 21)
    ff = None
    if flat:
        ff = (1.0 + 0.05 * np.random.default_rng(5).standard_normal(h * w)).astype(np.float64)
    kw = dict(flatfield=ff) if flat else {}
    c = pkg.Correlator(dq, sq, F, dpl=8, **kw)
    c.push_sparse(idx, val, off)
    c.finish_ingest()
    got = c.frames(N)
    c.close()
    qm = oracle.QMap(dq, sq)
    fo = oracle.sparse_filter(qm, F, off, idx, val, flat=ff, stride=1, avg=1, swindow=F // 10)
    want = np.zeros((N, h * w), np.float32)
    ptr, t, v = fo.rows.row_ptr, fo.rows.t, fo.rows.v
    for p in range(h * w):
        for k in range(int(ptr[p]), int(ptr[p + 1])):
            if t[k] < N:
                want[t[k], p] = v[k]
    if flat:
        assert_close(got, want, "frames_out")
    else:
        assert_exact(got, want, "frames_out")


@pytest.mark.parametrize("flat,F,sw", [(False, 400, 40), (True, 400, 40), (False, 333, 37), (True, 250, 1)])
def test_late_window_partition_sums(pkg, oracle, flat, F, sw):
    """XPCS_COMPAT_LATE_WINDOW (the Rigaku reader's static-window rule, io/rigaku.cpp:190-193): the partial
    partition means of the integer and of the float store against the oracle's restatement; everything else
    is unchanged by the flag."""
    h, w = 32, 40
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.03, 31)
    ff = (1.0 + 0.05 * np.random.default_rng(6).standard_normal(h * w)).astype(np.float64) if flat else None
    kw = dict(flatfield=ff) if flat else {}
    out = {}
    for late in (False, True):
        c = pkg.Correlator(dq, sq, F, dpl=8, static_window=sw, late_window=late, **kw)
        c.push_sparse(idx, val, off)
        out[late] = c.finish_ingest()
        c.close()
    qm = oracle.QMap(dq, sq)
    fo = oracle.sparse_filter(qm, F, off, idx, val, flat=ff, stride=1, avg=1, swindow=sw, late_window=True)
    oracle.post_scale(qm, F, sw, fo)
    W = F // sw
    want = fo.part_partial[: W * qm.S].reshape(W, qm.S)
    chk = assert_close if flat else assert_exact
    chk(out[True]["part_partial"], want, "partition-mean-partial (late windows)")
    for key in ("pixel_sum", "frame_sum", "part_total"):
        assert_exact(out[True][key], out[False][key], key)
    if sw > 1:
        assert np.any(out[True]["part_partial"] != out[False]["part_partial"])


def test_pinned_result_buffers_give_the_same_arrays(pkg):
    """Correlator.pinned_results(): results land in page-locked buffers owned by the library (xpcs_host_alloc) and are
    reused from call to call; the values are those of the default (fresh numpy arrays) path."""
    dq, sq, off, idx, val = make_case(pkg, 40, 32, 600, 0.03, 77)
    out = []
    for pinned in (False, True):
        c = pkg.Correlator(dq, sq, 600)
        c.pinned_results(pinned)
        for _ in range(2):  # the second round reuses the buffers
            c.reset()
            c.push_sparse(idx, val, off)
            sums = c.finish_ingest()
            c.multitau(want=False)
            g2, se = c.normalize()
        out.append({k: np.array(v, copy=True) for k, v in sums.items()} | {"g2": np.array(g2, copy=True), "se": np.array(se, copy=True)})
        c.close()
    for k in out[0]:
        assert np.array_equal(out[0][k], out[1][k], equal_nan=True), k


def test_push_sparse_device_rejects_misaligned_pointers(pkg):
    """A sliced device view (d_idx not 16-byte aligned) must come back as XPCS_E_ARG, not as a misaligned-address fault."""
    dq, sq, off, idx, val = make_case(pkg, 16, 16, 50, 0.05, 78)
    c = pkg.Correlator(dq, sq, 50)
    with pytest.raises(pkg.XpcsError) as e:
        c.push_sparse_device(0x7f0000000004, 0x7f0000100000, 0x7f0000200000, 10, 50)
    assert e.value.code == -1 and "aligned" in str(e.value)
    c.close()


def test_duplicates_inside_a_frame_are_merged(pkg, oracle):
    """A pixel listed twice in one frame is summed by the Filter stage (sparse_filter.cpp:152-158): rows of ~75 events,
    one of them with a duplicate, rows that live in the last tenth of the run only, rows that never fire."""
    dq, sq, off, idx, val = make_case(pkg, 32, 32, 1500, 0.05, 82)
    F = 1500
    fr = np.repeat(np.arange(F), np.diff(off))
    keep = ~(((idx >= 100) & (idx < 140) & (fr < 1350)) | ((idx >= 200) & (idx < 210)))
    cnt = np.bincount(fr[keep], minlength=F)
    idx, val = idx[keep], val[keep]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    a, b = int(off[700]), int(off[701])   # frame 700: its first event once more at the end of the frame
    idx = np.concatenate([idx[:b], idx[a:a + 1], idx[b:]])
    val = np.concatenate([val[:b], np.array([2], np.int16), val[b:]])
    off = off.copy()
    off[701:] += 1
    got = run_gpu(pkg, dq, sq, F, off, idx, val)
    ref = run_oracle(oracle, dq, sq, F, off, idx, val)
    compare_exact(got, ref)
