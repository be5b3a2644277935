"""Pipelined sparse ingest (xpcs_push_sparse cut into chunks of frames that are ingested while the next
chunk crosses PCIe, chunk stores concatenated at xpcs_finish_ingest) must give exactly what the one-pass
ingest gives, and what the oracle gives: every sum of the integer path is an exact integer, so the
chunking cannot change a bit.  The chunking is forced on small inputs through the environment knobs the
library reads at push time."""
import os

import numpy as np
import pytest

from conftest import make_case
from test_gpu_parity import assert_exact, run_oracle

pytestmark = pytest.mark.gpu


def run(pkg, dq, sq, F, off, idx, val, pushes, env, **kw):
    old = {k: os.environ.get(k) for k in ("XPCS_PIPELINE_MIN_EVENTS", "XPCS_PIPELINE_CHUNKS", "XPCS_NO_PIPELINE")}
    for k in old:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        c = pkg.Correlator(dq, sq, F, **kw)
        cuts = np.linspace(0, F, pushes + 1).astype(int)
        for a, b in zip(cuts[:-1], cuts[1:]):
            o = off[a:b + 1]
            c.push_sparse(idx[o[0]:o[-1]], val[o[0]:o[-1]], o - o[0])
        sums = c.finish_ingest()
        G = c.multitau(want=True)
        g2, se = c.normalize()
        info = c.info()
        c.close()
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    return sums, G, g2, se, info


@pytest.mark.parametrize("h,w,F,occ,seed,dpl,chunks,pushes", [
    (64, 64, 1000, 0.01, 3, 8, 4, 1),
    (48, 40, 601, 0.02, 2, 8, 7, 1),
    (96, 96, 4000, 0.002, 6, 8, 3, 2),    # two pushes of three chunks each
    (40, 56, 2049, 0.004, 5, 4, 16, 1),   # as many chunks as the table holds
    (24, 24, 33, 0.3, 7, 8, 5, 1),
])
def test_pipelined_ingest_equals_one_pass_and_oracle(pkg, oracle, h, w, F, occ, seed, dpl, chunks, pushes):
    dq, sq, off, idx, val = make_case(pkg, h, w, F, occ, seed)
    piped = run(pkg, dq, sq, F, off, idx, val, pushes,
                {"XPCS_PIPELINE_MIN_EVENTS": "1", "XPCS_PIPELINE_CHUNKS": str(chunks)}, dpl=dpl, compat=True)
    plain = run(pkg, dq, sq, F, off, idx, val, pushes, {"XPCS_NO_PIPELINE": "1"}, dpl=dpl, compat=True)
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=True)
    assert piped[4].value_kind == 0 and plain[4].value_kind == 0
    assert piped[4].events_stored == plain[4].events_stored
    for ref, tag in ((plain, "one-pass"), ((rs, rG, rg2, rse), "oracle")):
        for k, name in enumerate(("G2", "IP", "IF")):
            assert_exact(piped[1][k], ref[1][k], "%s vs %s" % (name, tag))
        for key in ("frame_sum", "pixel_sum", "part_total", "part_partial"):
            assert_exact(piped[0][key], ref[0][key], "%s vs %s" % (key, tag))
        assert_exact(piped[2], ref[2], "norm-0-g2 vs %s" % tag)
    assert_exact(piped[3], plain[3], "stderr vs one-pass")


def test_pipeline_abandoned_on_large_counts(pkg, oracle):
    """A count beyond the packed word (>= 4096) in a later chunk: the chunks already ingested are dropped
    and the one-pass ingest (float values) takes over at finish_ingest."""
    dq, sq, off, idx, val = make_case(pkg, 32, 32, 400, 0.05, 11)
    val = val.copy()
    valid = (dq.ravel()[idx] > 0) & (sq.ravel()[idx] > 0)
    e = int(off[300]) + int(np.argmax(valid[int(off[300]):int(off[301])]))
    assert valid[e]
    val[e] = 5000
    piped = run(pkg, dq, sq, 400, off, idx, val, 1, {"XPCS_PIPELINE_MIN_EVENTS": "1", "XPCS_PIPELINE_CHUNKS": "4"},
                dpl=8, compat=True)
    plain = run(pkg, dq, sq, 400, off, idx, val, 1, {"XPCS_NO_PIPELINE": "1"}, dpl=8, compat=True)
    assert piped[4].value_kind == 1 and plain[4].value_kind == 1
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_exact(piped[1][k], plain[1][k], name)
    for key in ("frame_sum", "pixel_sum", "part_total", "part_partial"):
        assert_exact(piped[0][key], plain[0][key], key)
    assert_exact(piped[2], plain[2], "norm-0-g2")


def test_pipelined_shards_sum_to_single_gpu_partials(pkg):
    """Pixel shards each running the pipelined ingest over the full event stream (foreign pixels
    skipped): the sum of their normalisation partials is the unsharded one-pass buffer bit for bit."""
    import torch
    dq, sq, off, idx, val = make_case(pkg, 64, 64, 700, 0.02, 61, n_dynamic=5, static_per_dynamic=4)
    F = 700

    def partials(k, K, env):
        old = {n: os.environ.get(n) for n in ("XPCS_PIPELINE_MIN_EVENTS", "XPCS_PIPELINE_CHUNKS", "XPCS_NO_PIPELINE")}
        for n in old:
            os.environ.pop(n, None)
        os.environ.update(env)
        try:
            c = pkg.Correlator(dq, sq, F, shard_index=k, shard_count=K)
            c.push_sparse(idx, val, off)
            c.finish_ingest(want=False)
            c.multitau(want=False)
            ptr, n = c.normalize_partials()
            torch.cuda.synchronize()
            host = pkg.torchio.device_view(ptr, n, "float64", "cuda:0").cpu().numpy().copy()
            c.normalize_finish()
            c.close()
        finally:
            for n, v in old.items():
                os.environ.pop(n, None)
                if v is not None:
                    os.environ[n] = v
        return host

    full = partials(0, 1, {"XPCS_NO_PIPELINE": "1"})
    total = sum(partials(k, 3, {"XPCS_PIPELINE_MIN_EVENTS": "1", "XPCS_PIPELINE_CHUNKS": "5"}) for k in range(3))
    assert_exact(total, full, "sum of pipelined shard partials")


def test_pipelined_ingest_with_empty_frames_and_reset(pkg, oracle):
    """Frames without events at chunk boundaries, an empty trailing push, and a second ingest on the same
    handle after xpcs_reset (the chunk stores are reused)."""
    dq, sq, off, idx, val = make_case(pkg, 32, 32, 240, 0.04, 17)
    # empty out frames 58..63 and 119..121 (chunk cuts of a 4-chunk split fall near 60, 120, 180)
    keep = np.ones(idx.size, bool)
    for a, b in ((58, 64), (119, 122)):
        keep[int(off[a]):int(off[b])] = False
    cnt = np.diff(off)
    for a, b in ((58, 64), (119, 122)):
        cnt[a:b] = 0
    idx2, val2 = idx[keep], val[keep]
    off2 = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    old = {n: os.environ.get(n) for n in ("XPCS_PIPELINE_MIN_EVENTS", "XPCS_PIPELINE_CHUNKS", "XPCS_NO_PIPELINE")}
    os.environ.pop("XPCS_NO_PIPELINE", None)
    os.environ.update({"XPCS_PIPELINE_MIN_EVENTS": "1", "XPCS_PIPELINE_CHUNKS": "4"})
    try:
        c = pkg.Correlator(dq, sq, 240, dpl=8, compat=True)
        outs = []
        for it in range(2):
            if it:
                c.reset()
            c.push_sparse(idx2, val2, off2)
            c.push_sparse(idx2[:0], val2[:0], np.zeros(1, np.int64))  # a push of zero frames
            sums = c.finish_ingest()
            G = c.multitau(want=True)
            g2, se = c.normalize()
            outs.append((sums, G, g2))
        c.close()
    finally:
        for n, v in old.items():
            os.environ.pop(n, None)
            if v is not None:
                os.environ[n] = v
    rs, rG, rg2, rse = run_oracle(oracle, dq, sq, 240, off2, idx2, val2, dpl=8, compat=True)
    for sums, G, g2 in outs:
        for k, name in enumerate(("G2", "IP", "IF")):
            assert_exact(G[k], rG[k], name)
        for key in ("frame_sum", "pixel_sum", "part_total", "part_partial"):
            assert_exact(sums[key], rs[key], key)
        assert_exact(g2, rg2, "norm-0-g2")
