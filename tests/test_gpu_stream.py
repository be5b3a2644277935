"""Online multi-tau (xpcs_stream_*: SURVEY.md 8 row f-1).  The frames reach the device in chunks of 2^k frames, each
chunk is folded into a per-pixel state by k_stream_chunk and forgotten.  The result must be, bit for bit, what the
resident path gives without the stale-tail flag (same integers, same single IEEE division: corr.cpp:420-424) and what
the oracle's exact-maths multiTau2 gives; the Filter sums and the normalised g2 / stderr with them.  The flag itself
needs complete rows and must be refused, not ignored."""
import numpy as np
import pytest

from conftest import make_case
from test_gpu_parity import assert_exact, run_oracle

pytestmark = pytest.mark.gpu


def run_resident(pkg, dq, sq, F, off, idx, val, **kw):
    c = pkg.Correlator(dq, sq, F, compat=False, **kw)
    c.push_sparse(idx, val, off)
    sums = c.finish_ingest()
    G = c.multitau(want=True)
    g2, se = c.normalize()
    c.close()
    return sums, G, g2, se


def run_stream(pkg, dq, sq, F, off, idx, val, chunk, push_frames=None, device=False, **kw):
    """push_frames: frames per xpcs_stream_push_sparse call (a multiple of chunk; None = one chunk per call)"""
    c = pkg.Correlator(dq, sq, F, compat=False, **kw)
    c.stream_begin(chunk)
    step = push_frames or chunk
    keep = []
    for a in range(0, F, step):
        b = min(F, a + step)
        o = off[a:b + 1]
        if device:
            import torch
            ti = torch.from_numpy(np.ascontiguousarray(idx[o[0]:o[-1]], np.int32)).cuda()
            tv = torch.from_numpy(np.ascontiguousarray(val[o[0]:o[-1]], np.int16)).cuda()
            to = torch.from_numpy(np.ascontiguousarray(o - o[0], np.int64)).cuda()
            torch.cuda.synchronize()
            c.stream_push_sparse_device(ti.data_ptr(), tv.data_ptr(), to.data_ptr(), int(o[-1] - o[0]), b - a)
            keep.append((ti, tv, to))
        else:
            c.stream_push_sparse(idx, val, o)   # offsets into the whole arrays: the library rebases them
    sums = c.stream_finish()
    info = c.info()
    G = c.multitau(want=True)
    g2, se = c.normalize()
    launches = c.launch_count()
    c.close()
    return sums, G, g2, se, info, launches


def compare(got, ref, tag, stderr_exact=True):
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_exact(got[1][k], ref[1][k], "%s vs %s" % (name, tag))
    for key in ("frame_sum", "pixel_sum", "part_total", "part_partial"):
        assert_exact(got[0][key], ref[0][key], "%s vs %s" % (key, tag))
    assert_exact(got[2], ref[2], "norm-0-g2 vs %s" % tag)
    if stderr_exact:
        assert_exact(got[3], ref[3], "norm-0-stderr vs %s" % tag)


@pytest.mark.parametrize("h,w,F,occ,seed,dpl,chunk", [
    (64, 64, 1000, 0.01, 3, 8, 64),      # levels 6.. arrive one bin per chunk (or rarer)
    (48, 40, 601, 0.02, 2, 8, 128),      # short last chunk
    (96, 96, 4000, 0.002, 6, 8, 256),
    (40, 56, 2049, 0.004, 5, 4, 512),    # one frame beyond four chunks
    (24, 24, 33, 0.3, 7, 8, 64),         # a single, short chunk
    (32, 32, 5000, 0.2, 9, 8, 2048),     # dense rows: every level below 11 through the 4-bin groups
    (16, 16, 9000, 0.05, 10, 4, 4096),
    (16, 24, 20000, 0.01, 12, 8, 8192),
])
def test_stream_equals_resident_and_oracle(pkg, oracle, h, w, F, occ, seed, dpl, chunk):
    dq, sq, off, idx, val = make_case(pkg, h, w, F, occ, seed)
    st = run_stream(pkg, dq, sq, F, off, idx, val, chunk, dpl=dpl)
    res = run_resident(pkg, dq, sq, F, off, idx, val, dpl=dpl)
    compare(st, res, "resident")
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=False)
    compare(st, (rs, rG, rg2, rse), "oracle", stderr_exact=False)
    ok = np.isfinite(rse)
    assert np.allclose(st[3][ok], rse[ok], rtol=1e-5, atol=0)
    assert st[4].events_pushed == idx.size and st[4].raw_frames_seen == F


def test_pushes_of_several_chunks_and_device_pushes(pkg):
    dq, sq, off, idx, val = make_case(pkg, 48, 48, 3000, 0.01, 21)
    ref = run_stream(pkg, dq, sq, 3000, off, idx, val, 128)
    many = run_stream(pkg, dq, sq, 3000, off, idx, val, 128, push_frames=128 * 5)   # 4 pushes of 5 chunks, one of 3.4
    dev = run_stream(pkg, dq, sq, 3000, off, idx, val, 128, device=True)
    whole = run_stream(pkg, dq, sq, 3000, off, idx, val, 128, push_frames=3072)      # everything in one call
    for other, tag in ((many, "multi-chunk pushes"), (dev, "device pushes"), (whole, "one push")):
        compare(other, ref, tag)
    assert ref[5] > 0


def test_chunk_length_does_not_matter(pkg):
    dq, sq, off, idx, val = make_case(pkg, 32, 48, 2500, 0.03, 22)
    ref = run_stream(pkg, dq, sq, 2500, off, idx, val, 64)
    for chunk in (256, 1024, 4096):
        compare(run_stream(pkg, dq, sq, 2500, off, idx, val, chunk), ref, "chunk %d vs 64" % chunk)


def test_event_walk_and_bin_walk_agree(pkg):
    """a level of a row's chunk is either walked bin by bin or event by event (multitau_stream_core.h: mac_groups /
    mac_events); XPCS_ST_EVENTS moves the switch-over: 0 = bins only, 1000 = events wherever a level has 32 bins"""
    import os
    dq, sq, off, idx, val = make_case(pkg, 32, 32, 6000, 0.04, 26)
    ref = run_stream(pkg, dq, sq, 6000, off, idx, val, 1024)
    old = os.environ.get("XPCS_ST_EVENTS")
    try:
        for knob in ("0", "1000", "3"):
            os.environ["XPCS_ST_EVENTS"] = knob
            compare(run_stream(pkg, dq, sq, 6000, off, idx, val, 1024), ref, "XPCS_ST_EVENTS=" + knob)
    finally:
        os.environ.pop("XPCS_ST_EVENTS", None)
        if old is not None:
            os.environ["XPCS_ST_EVENTS"] = old


def test_bright_and_silent_pixels(pkg, oracle):
    """a pixel lit in every frame with large counts, duplicates of a pixel inside a frame, frames without any event,
    pixels that never fire"""
    h, w, F = 16, 16, 1500
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.02, 23)
    frames = np.repeat(np.arange(F), np.diff(off))
    keep = (frames < 300) | (frames >= 420)          # 120 silent frames (almost two chunks of 64)
    idx, val, frames = idx[keep], val[keep].copy(), frames[keep]
    valid = np.flatnonzero((dq.ravel() > 0) & (sq.ravel() > 0))
    hot = valid[len(valid) // 2]
    idx = np.concatenate([idx, np.full(F, hot, np.int32), np.full(F // 2, hot, np.int32)])
    val = np.concatenate([val, np.full(F, 900, np.int16), np.full(F // 2, 7, np.int16)])
    frames = np.concatenate([frames, np.arange(F), np.arange(0, F - 1, 2)[: F // 2]])
    order = np.argsort(frames, kind="stable")
    idx, val, frames = idx[order].astype(np.int32), val[order].astype(np.int16), frames[order]
    off = np.zeros(F + 1, np.int64)
    off[1:] = np.cumsum(np.bincount(frames, minlength=F))
    st = run_stream(pkg, dq, sq, F, off, idx, val, 64)
    res = run_resident(pkg, dq, sq, F, off, idx, val)
    compare(st, res, "resident")
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val, compat=False)
    for k, name in enumerate(("IP", "IF")):
        assert_exact(st[1][k + 1], rG[k + 1], "%s vs oracle" % name)
    # the hot pixel's G2 numerators exceed 2^24: the reference's fp32 running sums round there
    assert np.allclose(st[1][0], rG[0], rtol=1e-5, atol=0)
    cold = np.ones(h * w, bool)
    cold[hot] = False
    assert_exact(st[1][0][:, cold], rG[0][:, cold], "G2 of the other pixels vs oracle")


def test_stream_on_a_shard(pkg):
    """a sharded handle streams the whole detector's chunks and keeps its own pixels"""
    dq, sq, off, idx, val = make_case(pkg, 48, 48, 1200, 0.01, 24)
    whole = run_stream(pkg, dq, sq, 1200, off, idx, val, 128)
    G2 = np.zeros_like(whole[1][0])
    for r in range(3):
        part = run_stream(pkg, dq, sq, 1200, off, idx, val, 128, shard_index=r, shard_count=3)
        G2 += part[1][0]
    assert_exact(G2, whole[1][0], "G2 summed over three shards")


def test_stream_refusals(pkg):
    dq, sq, off, idx, val = make_case(pkg, 16, 16, 300, 0.05, 25)
    c = pkg.Correlator(dq, sq, 300, compat=True)
    with pytest.raises(pkg.XpcsError) as e:
        c.stream_begin(64)
    assert e.value.code == -1 and "STALE_TAIL" in str(e.value)
    c.close()
    c = pkg.Correlator(dq, sq, 300, compat=False)
    for bad in (63, 100, 16384, 32):
        with pytest.raises(pkg.XpcsError):
            c.stream_begin(bad)
    with pytest.raises(pkg.XpcsError):   # no stream open
        c.stream_push_sparse(idx, val, off[:65])
    c.stream_begin(64)
    with pytest.raises(pkg.XpcsError):   # resident pushes are closed while a stream is open
        c.push_sparse(idx, val, off)
    with pytest.raises(pkg.XpcsError):   # nor can the resident finish end it
        c.finish_ingest()
    c.stream_push_sparse(idx, val, off[:65])
    with pytest.raises(pkg.XpcsError) as e:   # 64 of 300 frames
        c.stream_finish()
    assert e.value.code == -3
    c.stream_push_sparse(idx, val, off[64:64 + 41])      # a short chunk ...
    with pytest.raises(pkg.XpcsError):
        c.stream_push_sparse(idx, val, off[104:104 + 65])  # ... has to be the last one
    c.reset()
    c.stream_begin(64)
    with pytest.raises(pkg.XpcsError):   # more frames than the job has
        c.stream_push_sparse(idx, val, np.concatenate([off, off[-1:] + np.arange(1, 30)]))
    c.reset()
    # after a reset the handle streams (and ingests) again
    c.stream_begin(128)
    c.stream_push_sparse(idx, val, off)
    c.stream_finish(want=False)
    G = c.multitau(want=True)
    with pytest.raises(pkg.XpcsError):   # no rows on the device
        c.frames(2)
    c.reset()
    c.push_sparse(idx, val, off)
    c.finish_ingest(want=False)
    G2 = c.multitau(want=True)
    c.close()
    for a, b in zip(G, G2):
        assert_exact(a, b, "stream vs resident on the same handle")
    bright = val.copy()
    bright[5] = 5000
    c = pkg.Correlator(np.ones_like(dq), np.ones_like(sq), 300, compat=False)
    c.stream_begin(64)
    with pytest.raises(pkg.XpcsError) as e:
        c.stream_push_sparse(idx, bright, off[:65])
    assert "packed" in str(e.value)
    c.close()


def test_corr_stream_frames(pkg, tmp_path):
    """`corr --stream_frames K`: every result dataset equals, bit for bit, what `corr --no_compat` writes for the same
    input (the resident path with exact pair sums) -- on a sparse IMM file and on a Rigaku event file."""
    import golden_util as G
    from test_gpu_corr_host import _run_corr
    for name, K in (("sparse_staletail_32x32", 128), ("rigaku_compact_32x40", 64)):
        c = G.Case(name)
        a = tmp_path / ("stream_" + name)
        b = tmp_path / ("resident_" + name)
        a.mkdir()
        b.mkdir()
        res_s, log = _run_corr(pkg, c, a, extra=["--stream_frames", str(K)])
        res_r, _ = _run_corr(pkg, c, b, extra=["--no_compat"])
        assert "exact multi-tau sums" in log
        assert sorted(res_s) == sorted(res_r)
        for k in res_r:
            assert G.n_diff(res_s[k], res_r[k]) == 0, "%s: %s" % (name, k)
