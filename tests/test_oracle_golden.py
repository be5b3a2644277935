"""Pins the CPU oracle (oracle/xpcs_oracle.c) to the reference itself: every fixture under
tests/golden/ was written by the UNMODIFIED reference binary (oracle/_ref/corr_ref, see
tests/golden/make_golden.py); the restatement must reproduce every result dataset bit for bit
(sparse integer, flat-field / stride / average, frame-sum normalisation, dense + dark +
threshold, two-time incl. sg / C / g2full / g2partials)."""
import numpy as np
import pytest

import golden_util as G


def _filter(O, c):
    qm = O.QMap(c.dq, c.sq)
    dark = None
    if c.kind in ("sparse", "twotime"):
        off, idx, val = c.inp["off"], c.inp["idx"], c.inp["val"]
        if c.fmt == "ufxc":  # the reader's restatement must deliver the stored events from the raw words
            h, w = c.dq.shape
            o2, i2, v2 = O.ufxc_frames(c.inp["words"], h, w, c.F_raw)
            assert np.array_equal(o2, off) and np.array_equal(i2, idx) and np.array_equal(v2, val)
        if c.fmt == "rigaku":  # the reader's restatement must deliver the stored events from the raw words
            h, w = c.dq.shape
            off, idx, val = O.rigaku_frames(c.inp["words"], h, w, 0, c.F, qm.mask, stride=c.stride, avg=c.avg)
            assert np.array_equal(off, c.inp["off"]) and np.array_equal(idx, c.inp["idx"]) and np.array_equal(val, c.inp["val"])
        fo = O.sparse_filter(qm, c.F, off, idx, val, flat=c.flat, stride=c.stride, avg=c.avg, swindow=c.swindow,
                             late_window=c.late_window)
    else:
        fr = c.inp["frames"]
        if c.darks:
            dark = O.dark_image(fr[: c.darks], c.flat)
        lld, sigma = c.inp["thresh"] if "thresh" in c.inp else (0.0, 0.0)
        fo = O.dense_filter(qm, c.F, fr[c.darks:], flat=c.flat, dark=dark, lld=float(lld), sigma=float(sigma),
                            swindow=c.swindow)
    frame_sum = fo.frame_sum.copy()
    O.post_scale(qm, c.F, c.swindow, fo, normalize_by_framesum=bool(c.norm))
    return qm, fo, frame_sum, dark


def _exact(a, b, what):
    assert np.asarray(a).size == np.asarray(b).size, what
    assert G.n_diff(a, b) == 0, "%s: %d entries differ from the reference" % (what, G.n_diff(a, b))


@pytest.mark.parametrize("name", G.names("sparse") + G.names("dense"))
def test_oracle_multitau_path_matches_reference(oracle, name):
    O, c = oracle, G.Case(name)
    qm, fo, frame_sum, dark = _filter(O, c)
    if dark is not None:
        _exact(dark[0], c.ref["DarkAvg"], "DarkAvg")
        _exact(dark[1], c.ref["DarkStd"], "DarkStd")
    _exact(fo.pixel_sum, c.ref["pixelSum"], "pixelSum")
    _exact(frame_sum, c.ref["frameSum"], "frameSum")
    _exact(fo.part_total[: qm.S], c.ref["partition-mean-total"], "partition-mean-total")
    _exact(fo.part_partial[: (c.F // c.swindow) * qm.S], c.ref["partition-mean-partial"], "partition-mean-partial")
    _, tv = O.delay_schedule(c.F, c.dpl)
    _exact(tv.astype(np.float32), c.ref["tau"], "tau")
    G2, IP, IF = O.multitau(qm.P, c.F, c.dpl, fo.rows, compat=True)
    _exact(G2, c.ref["G2"], "G2")
    _exact(IP, c.ref["IP"], "IP")
    _exact(IF, c.ref["IF"], "IF")
    g2, se = O.normalize(qm, G2, IP, IF)
    _exact(g2, c.ref["norm-0-g2"], "norm-0-g2")
    _exact(se, c.ref["norm-0-stderr"], "norm-0-stderr")


def test_fixtures_exercise_the_stale_tail_quirk(oracle):
    """The exact pair sums (compat off) must differ from the reference in a few entries of these
    fixtures -- otherwise the fixtures would not pin the lower_bound behaviour."""
    hit = 0
    for name in ("sparse_staletail_32x32", "staletail_hand_example", "sparse_odd_dpl4"):
        c = G.Case(name)
        qm, fo, _, _ = _filter(oracle, c)
        Ge, IPe, IFe = oracle.multitau(qm.P, c.F, c.dpl, fo.rows, compat=False)
        d = G.n_diff(Ge, c.ref["G2"])
        assert d > 0, name
        assert G.n_diff(IPe, c.ref["IP"]) == 0 and G.n_diff(IFe, c.ref["IF"]) == 0
        low = (Ge != c.ref["G2"])
        assert (Ge[low] > c.ref["G2"][low]).all(), "the quirk only ever drops pairs"
        hit += d
    assert hit >= 3


@pytest.mark.parametrize("name", G.names("twotime"))
def test_oracle_twotime_matches_reference(oracle, name):
    O, c = oracle, G.Case(name)
    qm, fo, _, _ = _filter(O, c)
    r = O.twotime(qm, c.F, fo.rows, c.inp["qbins"], int(c.inp["wsize"]), method=c.method,
                  average=str(c.inp["filt"]) == "Average")
    _exact(r["sg"], c.ref["sg"], "sg")
    if "framethreading" in name:
        # corr --frame_threading (twotimeFrameThreading, corr.cpp:574-779): the same C in another summation order;
        # the restatement follows the q-bin variant (corr.cpp:781-924), so this fixture pins "equal within 1e-5"
        for q in r["bins"]:
            err, nanmis = G.rel_err(r["C"][q], c.ref["C2T_all/g2_%05d" % q])
            assert nanmis == 0 and err <= 1e-5, (q, err)
        for k in ("g2full", "g2partials"):
            err, nanmis = G.rel_err(r[k], c.ref[k])
            assert nanmis == 0 and err <= 1e-5, (k, err)
        return
    for q in r["bins"]:
        _exact(r["C"][q], c.ref["C2T_all/g2_%05d" % q], "C2T_all/g2_%05d" % q)
    _exact(r["g2full"], c.ref["g2full"], "g2full")
    _exact(r["g2partials"], c.ref["g2partials"], "g2partials")


def test_reference_binary_reproduces_fixture_when_present(pkg):
    """In the build container (where /root/reference exists and oracle/_ref/corr_ref was built)
    re-run the reference on one fixture's inputs: the committed fixture is what it produces."""
    from oracle import refdrv
    if not refdrv.available():
        pytest.skip("oracle/_ref/corr_ref not built here")
    c = G.Case("sparse_int_24x24")
    res, _ = refdrv.run_case(pkg.synth, c.dq, c.sq, c.F_raw, sparse=(c.inp["off"], c.inp["idx"], c.inp["val"]),
                             g2out=True, dpl=c.dpl, static_window=c.swindow)
    for k in ("G2", "IP", "IF", "norm-0-g2", "norm-0-stderr", "pixelSum", "frameSum"):
        _exact(res[k], c.ref[k], k)
