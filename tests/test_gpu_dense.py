"""Dense (non-sparse IMM) path through the C-ABI against the CPU oracle on seeded frames:
dark image, dense filter (dark subtraction, clamp, lld + sigma*std threshold, flat-field),
store build, float-row multi-tau (slice kernel, warp-per-row kernel and lane-per-row kernel) and
normalisation.  Float results within the 1e-5 relative tolerance of BASELINE.json's
north_star; the set of surviving samples and DarkAvg/DarkStd must match exactly.
Reference: filter/dense_filter.cpp:121-210, data_structure/dark_image.cpp:81-106,
corr.cpp:315-431."""
import numpy as np
import pytest

from test_gpu_parity import assert_close, assert_exact

pytestmark = pytest.mark.gpu
AUX_RTOL = 1e-4  # see compare()


def oracle_dense(O, dq, sq, F, data, flat=None, darks=None, lld=0.0, sigma=0.0, stride=1, avg=1, dpl=8,
                 swindow=None, compat=True):
    qm = O.QMap(dq, sq)
    swindow = swindow or max(1, F // 10)
    dark = None
    if darks is not None:
        dark = O.dark_image(darks, flat if flat is not None else np.ones(qm.P))
    fo = O.dense_filter(qm, F, data, flat=flat, dark=dark, lld=lld, sigma=sigma, stride=stride, avg=avg,
                        swindow=swindow)
    frame_sum = fo.frame_sum.copy()
    O.post_scale(qm, F, swindow, fo)
    G2, IP, IF = O.multitau(qm.P, F, dpl, fo.rows, compat=compat)
    g2, se = O.normalize(qm, G2, IP, IF)
    W = F // swindow
    return dict(frame_sum=frame_sum, pixel_sum=fo.pixel_sum, part_total=fo.part_total[: qm.S],
                part_partial=fo.part_partial[: W * qm.S].reshape(W, qm.S), n=fo.n,
                G=(G2, IP, IF), g2=g2, se=se, dark=dark)


def gpu_dense(pkg, dq, sq, F, data, darks=None, chunks=None, report=False, **kw):
    c = pkg.Correlator(dq, sq, F, **kw)
    if darks is not None:
        c.set_dark(darks)
    if chunks is None:
        c.push_dense(data)
    else:
        a = 0
        for n in chunks:
            c.push_dense(data[a: a + n])
            a += n
        assert a == data.shape[0]
    sums = c.finish_ingest()
    G = c.multitau()
    g2, se = c.normalize()
    info = c.info()
    fb = c.multitau_fallback_slices()
    rep = c.kernel_report() if report else None
    c.close()
    return dict(sums=sums, G=G, g2=g2, se=se, info=info, fallback=fb, report=rep)


def compare(got, ref, what="", check_n=True):
    assert got["info"].value_kind == 1
    if check_n:  # survivors of the filter (one event per surviving sample when stride = avg = 1)
        assert got["info"].events_pushed == ref["n"], "%s: %d survivors, oracle keeps %d" % (
            what, got["info"].events_pushed, ref["n"])
    for k, nm in enumerate(("G2", "IP", "IF")):
        assert_close(got["G"][k], ref["G"][k], what + nm)
    # The Filter's auxiliary sums are sequential fp32 chains of thousands of terms in the
    # reference (frames_sum_, partitions_mean_ += v, dense_filter.cpp:176-190): its own rounding
    # noise is ~sqrt(n) * 6e-8, i.e. up to a few 1e-5 here, while the device sums in fp64.  These
    # four are not among the quantities north_star bounds at 1e-5 (G2/IP/IF, g2); they get 1e-4.
    assert_close(got["sums"]["frame_sum"], ref["frame_sum"], what + "frameSum", rtol=AUX_RTOL)
    assert_close(got["sums"]["pixel_sum"], ref["pixel_sum"], what + "pixelSum", rtol=AUX_RTOL)
    assert_close(got["sums"]["part_total"], ref["part_total"], what + "partition-mean-total", rtol=AUX_RTOL)
    assert_close(got["sums"]["part_partial"], ref["part_partial"], what + "partition-mean-partial", rtol=AUX_RTOL)
    assert_close(got["g2"], ref["g2"], what + "norm-0-g2")
    # at the last delays one bin pair is left and every pixel's G2/(IP*IF) is the same number:
    # the reference's spread is exactly 0 there, a one-ulp difference in IP or IF makes ours ~1e-8
    # (typical values are 1e-2); hence the absolute floor
    ok = np.isfinite(ref["se"])
    assert_close(got["se"][ok], ref["se"][ok], what + "norm-0-stderr", atol=1e-6)


def make_dense(pkg, h, w, F, darks, seed, mu=0.05, flat_sigma=0.05):
    """Dark offset 100 ADU, 20 ADU per photon.  The reference subtracts dark_avg = mean(raw * flat)
    from the un-flat-fielded raw value, so with the default 5 % flat-field spread a
    few percent of the pixels sit above their threshold in nearly every frame: rows of thousands
    of events next to rows of tens, which exercises the long-row paths (warp-per-row finalize,
    lane-per-row multi-tau) alongside the short-row ones."""
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=4, static_per_dynamic=3, r_min=2.0)
    fr = pkg.synth.dense_frames(h * w, F, darks=darks, mu=mu, seed=seed)
    flat = pkg.synth.flatfield(h * w, sigma=flat_sigma, seed=seed + 1)
    return dq, sq, fr[:darks], fr[darks:], flat


@pytest.mark.parametrize("h,w,F,darks,lld,sigma,dpl,seed", [
    (32, 32, 600, 20, 5.0, 3.0, 8, 1),      # bench workload c2 in small: a few percent survive
    (40, 24, 1500, 10, 4.0, 2.0, 8, 2),     # ~14 levels' worth of rows, sparse and dense levels
    (32, 32, 333, 5, 8.0, 0.0, 4, 3),       # odd frame count, dpl 4
    (16, 16, 4000, 8, 6.0, 3.0, 8, 4),      # long rows
])
@pytest.mark.parametrize("scalar", [False, True])
def test_dark_flat_threshold(pkg, oracle, h, w, F, darks, lld, sigma, dpl, seed, scalar):
    dq, sq, dk, data, flat = make_dense(pkg, h, w, F, darks, seed)
    ref = oracle_dense(oracle, dq, sq, F, data, flat=flat, darks=dk, lld=lld, sigma=sigma, dpl=dpl)
    got = gpu_dense(pkg, dq, sq, F, data, darks=dk, flatfield=flat, lld=lld, sigma=sigma, dpl=dpl,
                    scalar_dense=scalar)
    assert 0 < ref["n"] < 0.5 * data.size
    compare(got, ref)


def test_vector_and_scalar_filters_agree_exactly(pkg):
    dq, sq, dk, data, flat = make_dense(pkg, 32, 32, 500, 12, 7)
    a = gpu_dense(pkg, dq, sq, 500, data, darks=dk, flatfield=flat, lld=5.0, sigma=3.0)
    b = gpu_dense(pkg, dq, sq, 500, data, darks=dk, flatfield=flat, lld=5.0, sigma=3.0, scalar_dense=True)
    for k in range(3):
        assert_exact(a["G"][k], b["G"][k], "G2/IP/IF vector vs scalar dense filter")
    assert_exact(a["sums"]["pixel_sum"], b["sums"]["pixel_sum"], "pixelSum")
    assert_exact(a["sums"]["frame_sum"], b["sums"]["frame_sum"], "frameSum")
    assert_exact(a["g2"], b["g2"], "norm-0-g2")


def test_slice_warp_and_lane_multitau_agree_on_float_rows(pkg, oracle, monkeypatch):
    """the three float-row kernels on the same store: lane = row / warps = tasks (k_multitau_slicef, the default),
    warp per row (k_multitau_warpf, XPCS_MTF_KERNEL=warp) and lane per row (k_multitau)"""
    dq, sq, dk, data, flat = make_dense(pkg, 32, 32, 2000, 40, 8, flat_sigma=0.005)  # no hot pixels
    kw = dict(darks=dk, flatfield=flat, lld=5.0, sigma=3.0)
    a = gpu_dense(pkg, dq, sq, 2000, data, report=True, **kw)
    monkeypatch.setenv("XPCS_MTF_KERNEL", "warp")
    w = gpu_dense(pkg, dq, sq, 2000, data, report=True, **kw)
    monkeypatch.delenv("XPCS_MTF_KERNEL")
    b = gpu_dense(pkg, dq, sq, 2000, data, lane_multitau=True, **kw)
    assert a["report"].get("k_multitau_slicef", (0, 0))[1] == 1, sorted(a["report"])
    assert "k_multitau_slicef" not in w["report"] and w["report"].get("k_multitau_warpf", (0, 0))[1] == 1
    assert a["fallback"] == 0, "the float slice kernel (and the warp-per-row kernel behind it) did not take every slice"
    assert w["fallback"] == 0, "the float warp-per-row kernel did not take every slice"
    assert b["fallback"] == -1
    for x, what in ((a, "slice"), (w, "warp")):
        for k, nm in enumerate(("G2", "IP", "IF")):
            assert_close(x["G"][k], b["G"][k], nm + " %s vs lane kernel" % what)
        # the pattern of exact zeros (pairs the reference's stale-tail search loses) must be identical
        assert_exact(x["G"][0] == 0.0, b["G"][0] == 0.0, "G2 zero pattern (%s)" % what)


def test_float_slice_kernel_is_deterministic_and_leaves_hot_rows_to_the_others(pkg, oracle):
    """Two pixels whose flat-field factor puts them above their threshold in every frame (the reference subtracts
    dark_avg = mean(raw * flat) from the un-flat-fielded raw value): rows of 2000 events next to rows of
    ~100.  The slice kernel flags their slices, the kernels behind it take them; two runs give the same bits
    (piece-private sums, added up in a fixed order)."""
    F = 2000
    dq, sq, dk, data, flat = make_dense(pkg, 32, 32, F, 20, 9, flat_sigma=0.005)
    valid = np.flatnonzero((dq.ravel() > 0) & (sq.ravel() > 0))
    flat = flat.copy()
    flat.ravel()[valid[[7, valid.size // 2]]] = 0.8
    kw = dict(darks=dk, flatfield=flat, lld=5.0, sigma=3.0)
    ref = oracle_dense(oracle, dq, sq, F, data, flat=flat, darks=dk, lld=5.0, sigma=3.0)
    a = gpu_dense(pkg, dq, sq, F, data, report=True, **kw)
    b = gpu_dense(pkg, dq, sq, F, data, **kw)
    assert a["report"].get("k_multitau_slicef", (0, 0))[1] == 1, sorted(a["report"])
    assert a["report"].get("k_multitau_warpf", (0, 0))[1] == 1, "the warp-per-row kernel runs behind the slice kernel"
    assert 1 <= a["fallback"] <= 2, "the two hot rows (beyond 1024 events) belong to the lane-per-row kernel: %d" % a["fallback"]
    for k in range(3):
        assert_exact(a["G"][k], b["G"][k], "two runs of the float slice kernel")
    compare(a, ref)


def test_no_darks_threshold_zero(pkg, oracle):
    """Without dark frames the threshold is 0: every positive sample survives (dense rows)."""
    h, w, F = 16, 16, 300
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=3, static_per_dynamic=2, r_min=1.0)
    rng = np.random.default_rng(11)
    data = rng.poisson(0.7, (F, h * w)).astype(np.int16)
    data[rng.random(data.shape) < 0.02] = -3   # negative and zero samples are dropped (v <= 0)
    ref = oracle_dense(oracle, dq, sq, F, data)
    got = gpu_dense(pkg, dq, sq, F, data)
    assert ref["n"] > 0.3 * data.size
    compare(got, ref)


def test_negative_lld_keeps_every_sample(pkg, oracle):
    """lld < 0 with darks: the clamp makes v >= 0 > thresh, so even raw = -32768 survives -- the
    case the int16 bound cannot express (every[] flag, exact path)."""
    h, w, F = 16, 16, 64
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=2, static_per_dynamic=2, r_min=1.0)
    rng = np.random.default_rng(12)
    fr = (100 + 3 * rng.standard_normal((8 + F, h * w))).astype(np.int16)
    fr[8 + 5, :7] = -32768
    fr[8 + 9, 3] = 32767
    ref = oracle_dense(oracle, dq, sq, F, fr[8:], darks=fr[:8], lld=-1.0, sigma=0.0)
    got = gpu_dense(pkg, dq, sq, F, fr[8:], darks=fr[:8], lld=-1.0, sigma=0.0)
    valid = int(((dq > 0) & (sq > 0)).sum())
    assert ref["n"] == valid * F
    compare(got, ref)


def test_extreme_raw_values_and_bounds(pkg, oracle):
    """Samples at the int16 limits and right at the threshold of each pixel."""
    h, w, F = 16, 24, 128
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=3, static_per_dynamic=2, r_min=1.0)
    rng = np.random.default_rng(13)
    fr = (50 + 2 * rng.standard_normal((16 + F, h * w))).astype(np.int16)
    data = fr[16:]
    # sweep raw values around dark + lld for every pixel so that both sides of the bound occur
    for f in range(F):
        data[f] = 50 + (f % 17) - 3
    data[7, ::5] = 32767
    data[8, ::7] = -32768
    flat = pkg.synth.flatfield(h * w, seed=14)
    ref = oracle_dense(oracle, dq, sq, F, data, flat=flat, darks=fr[:16], lld=4.0, sigma=1.5)
    got = gpu_dense(pkg, dq, sq, F, data, darks=fr[:16], flatfield=flat, lld=4.0, sigma=1.5)
    assert 0 < ref["n"] < data.size
    compare(got, ref)


def test_odd_pixel_count_uses_scalar_kernel(pkg, oracle):
    h, w, F = 15, 17, 200   # 255 pixels: not a multiple of 8
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=3, static_per_dynamic=2, r_min=1.0)
    fr = pkg.synth.dense_frames(h * w, F, darks=6, mu=0.05, seed=15)
    ref = oracle_dense(oracle, dq, sq, F, fr[6:], darks=fr[:6], lld=5.0, sigma=3.0)
    got = gpu_dense(pkg, dq, sq, F, fr[6:], darks=fr[:6], lld=5.0, sigma=3.0)
    compare(got, ref)


@pytest.mark.parametrize("stride,avg", [(2, 1), (1, 2), (2, 2)])
def test_stride_and_average(pkg, oracle, stride, avg):
    h, w, F_raw = 24, 24, 480
    block = stride * avg if (stride > 1 and avg > 1) else max(stride, avg)
    F = F_raw // block
    dq, sq, dk, data, flat = make_dense(pkg, h, w, F_raw, 10, 16, mu=0.1)
    ref = oracle_dense(oracle, dq, sq, F, data, flat=flat, darks=dk, lld=5.0, sigma=3.0, stride=stride, avg=avg)
    got = gpu_dense(pkg, dq, sq, F, data, darks=dk, flatfield=flat, lld=5.0, sigma=3.0, stride=stride, avg=avg)
    compare(got, ref, check_n=False)


def test_chunked_pushes_equal_one_push(pkg):
    dq, sq, dk, data, flat = make_dense(pkg, 32, 32, 700, 10, 17)
    a = gpu_dense(pkg, dq, sq, 700, data, darks=dk, flatfield=flat, lld=5.0, sigma=3.0)
    b = gpu_dense(pkg, dq, sq, 700, data, darks=dk, flatfield=flat, lld=5.0, sigma=3.0, chunks=[1, 63, 64, 200, 372])
    for k in range(3):
        assert_exact(a["G"][k], b["G"][k], "chunked dense pushes")
    assert_exact(a["g2"], b["g2"], "norm-0-g2")
    assert_exact(a["sums"]["pixel_sum"], b["sums"]["pixel_sum"], "pixelSum")


def test_float_sparse_rows_long_and_short(pkg, oracle):
    """Flat-fielded sparse input at the shapes of the bench workloads: the float warp-per-row
    kernel against the oracle (1e-5) with and without the stale-tail behaviour."""
    from conftest import make_case
    from test_gpu_parity import run_gpu, run_oracle
    for (h, w, F, occ, seed, dpl) in [(48, 48, 20000, 0.015, 21, 8), (64, 64, 100000, 0.001, 22, 8),
                                      (32, 32, 1500, 0.2, 23, 8), (32, 32, 500, 0.9, 24, 4)]:
        dq, sq, off, idx, val = make_case(pkg, h, w, F, occ, seed)
        flat = pkg.synth.flatfield(h * w, seed=seed)
        for compat in (True, False):
            sums, G, g2, se, info = run_gpu(pkg, dq, sq, F, off, idx, val, dpl=dpl, compat=compat, flatfield=flat)
            rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=compat, flat=flat)
            assert info.value_kind == 1
            for k, nm in enumerate(("G2", "IP", "IF")):
                assert_close(G[k], rG[k], "%s F=%d compat=%d" % (nm, F, compat))
            assert_close(g2, rg2, "norm-0-g2 F=%d" % F)
