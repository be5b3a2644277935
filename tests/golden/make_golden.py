"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/corr_ref, built
by oracle/ref/Makefile from /root/reference) on small seeded inputs.  Each fixture holds the
inputs (so nothing has to be regenerated bit-for-bit at test time) and every result dataset
the reference wrote.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The fixtures pin the CPU oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_golden.py, -m gpu) to the reference's own outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle import refdrv  # noqa: E402

pkg = entry.load_package()
S = pkg.synth


def save(name, inputs, res, info):
    out = {("in_" + k): v for k, v in inputs.items() if v is not None}
    for k, v in res.items():
        out["ref_" + k.replace("/", "__")] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, name + ".npz"))
    print("%-22s %6.0f kB  datasets: %s" % (name, sz / 1e3, " ".join(sorted(res))))


def sparse_case(name, h, w, F_raw, occ, seed, dpl=8, nd=4, spd=3, r_min=2.0, flat=False, stride=1, avg=1,
                norm=False, events=None):
    dq, sq = S.annular_qmaps(h, w, n_dynamic=nd, static_per_dynamic=spd, r_min=r_min)
    if events is None:
        off, idx, val = S.sparse_frames(h * w, F_raw, occ, seed=seed)
    else:
        dq, sq, off, idx, val = events
    ff = S.flatfield(h * w, seed=seed + 100) if flat else None
    block = stride * avg if (stride > 1 and avg > 1) else max(stride, avg)
    F = F_raw // block
    sw = max(1, F // 10)
    res, info = refdrv.run_case(S, dq, sq, F_raw, sparse=(off, idx, val), g2out=True, dpl=dpl, stride=stride, avg=avg,
                                static_window=sw, flatfield=ff, normalize_by_framesum=norm)
    save(name, dict(kind=np.array("sparse"), dq=dq, sq=sq, off=off, idx=idx, val=val, flat=ff,
                    params=np.array([F_raw, dpl, stride, avg, sw, int(norm)], np.int64)), res, info)


def ufxc_case(name, h, w, F, occ, seed, f0=1900):
    """--ufxc (io/ufxc.cpp): the event words go through the reader's frame-counter unwrapping (the counter
    wraps at 2048: f0 = 1900 puts the wrap inside the run), empty frames, a pixel hit twice in a frame, a
    zero count.  The fixture is a "sparse" one (the decoded events in file order) that also carries the raw
    words the reference read."""
    dq, sq = S.annular_qmaps(h, w, n_dynamic=4, static_per_dynamic=3, r_min=2.0)
    off, idx, val = S.sparse_frames(h * w, F, occ, seed=seed)
    val = np.minimum(val, 3).astype(np.int16)
    cnt = np.diff(off)
    keep = np.ones(idx.size, bool)
    for f in (10, 11, 200):                       # empty frames (the first frame must have events)
        keep[int(off[f]):int(off[f + 1])] = False
        cnt[f] = 0
    idx, val = idx[keep], val[keep]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    val[int(off[5])] = 0                           # an event with count 0
    # frame 7: its first pixel once more at the end of the frame (duplicate, not adjacent in the file)
    a, b = int(off[7]), int(off[8])
    idx = np.concatenate([idx[:b], idx[a:a + 1], idx[b:]])
    val = np.concatenate([val[:b], np.array([2], np.int16), val[b:]])
    off = off.copy()
    off[8:] += 1
    words = S.ufxc_words(h, w, off, idx, val, f0=f0)
    sw = max(1, F // 10)
    res, info = refdrv.run_case(S, dq, sq, F, ufxc=words, g2out=True, dpl=8, static_window=sw)
    save(name, dict(kind=np.array("sparse"), fmt=np.array("ufxc"), words=words, dq=dq, sq=sq, off=off,
                    idx=idx.astype(np.int32), val=val, params=np.array([F, 8, 1, 1, sw, 0], np.int64)), res, info)


def rigaku_case(name, h, w, F, occ, seed, stride=1, avg=1, flat=False):
    """--rigaku (io/rigaku.cpp): 64-bit event words; the reader itself plays the Filter stage.  Frames without
    events vanish (the output frames are the non-empty ones, renumbered), events inside a frame come unsorted
    and a pixel twice, the file holds more frames than are asked for, and the static windows follow the
    reader's own rule (XPCS_COMPAT_LATE_WINDOW).  Stored as a "sparse" fixture: the events the reader's
    restatement (oracle.rigaku_frames) delivers, plus the raw words the reference read."""
    from oracle import oracle as O
    dq, sq = S.annular_qmaps(h, w, n_dynamic=4, static_per_dynamic=3, r_min=2.0)
    block = stride * avg if (stride > 1 and avg > 1) else max(stride, avg)
    n_file = F + 40      # F = raw frames to do (the reader folds `block` of them into one output frame)
    off, idx, val = S.sparse_frames(h * w, n_file, occ, seed=seed)
    val = np.minimum(val, 7).astype(np.int16)
    ff = S.flatfield(h * w, seed=seed + 100) if flat else None
    rng = np.random.default_rng(seed)
    cnt = np.diff(off)
    keep = np.ones(idx.size, bool)
    for f in (3, 4, 57, 300):                     # file frames without events
        keep[int(off[f]):int(off[f + 1])] = False
        cnt[f] = 0
    idx, val = idx[keep], val[keep]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    # shuffle the events inside every frame, and hit the first pixel of frame 9 twice
    order = np.concatenate([a + rng.permutation(b - a) for a, b in zip(off[:-1], off[1:])]).astype(np.int64)
    idx, val = idx[order], val[order]
    a, b = int(off[9]), int(off[10])
    idx = np.concatenate([idx[:b], idx[a:a + 1], idx[b:]])
    val = np.concatenate([val[:b], np.array([3], np.int16), val[b:]])
    off = off.copy()
    off[10:] += 1
    words = S.rigaku_words(h, w, np.arange(1, n_file + 1), off, idx, val)
    sw = max(1, (F // block) // 10)
    res, info = refdrv.run_case(S, dq, sq, F, rigaku=words, g2out=True, dpl=8, static_window=sw, stride=stride, avg=avg,
                                flatfield=ff)
    qm = O.QMap(dq, sq)
    eoff, eidx, evalv = O.rigaku_frames(words, h, w, 0, F // block, qm.mask, stride=stride, avg=avg)
    save(name, dict(kind=np.array("sparse"), fmt=np.array("rigaku"), words=words, dq=dq, sq=sq, off=eoff, idx=eidx,
                    val=evalv, flat=ff, params=np.array([F, 8, stride, avg, sw, 0], np.int64)), res, info)


def hdf5_case(name, h, w, F, occ, seed, begin=4, dtype=np.uint16):
    """--hdf5 (io/hdf5.cpp): a dense frame stack /entry/data/data; every non-zero sample is an event, the frames
    before data_begin_todo are skipped.  Stored as a "sparse" fixture (the events of the frames that are read)
    plus the stack itself."""
    dq, sq = S.annular_qmaps(h, w, n_dynamic=4, static_per_dynamic=3, r_min=2.0)
    n_file = F + begin - 1 + 5
    off, idx, val = S.sparse_frames(h * w, n_file, occ, seed=seed)
    stack = np.zeros((n_file, h * w), dtype)
    fr = np.repeat(np.arange(n_file), np.diff(off))
    stack[fr, idx] = val
    stack = stack.reshape(n_file, h, w)
    sw = max(1, F // 10)
    res, info = refdrv.run_case(S, dq, sq, F, hdf5=stack, g2out=True, dpl=8, static_window=sw, begin=begin)
    a, b = int(off[begin - 1]), int(off[begin - 1 + F])
    save(name, dict(kind=np.array("sparse"), fmt=np.array("hdf5"), stack=stack, dq=dq, sq=sq,
                    off=(off[begin - 1: begin + F] - a).astype(np.int64), idx=idx[a:b].astype(np.int32), val=val[a:b],
                    begin=np.array(begin), params=np.array([F, 8, 1, 1, sw, 0], np.int64)), res, info)


def twotime_cases():
    """two-time, symmetric smoothing, with and without the "Average" filter; round 2: StaticMap smoothing (both
    filters) and one run of the reference's other summation order (--frame_threading, corr.cpp:574-779)."""
    h = w = 16
    F = 200
    dq, sq = S.annular_qmaps(h, w, n_dynamic=3, static_per_dynamic=2, r_min=1.0)
    off, idx, val = S.sparse_frames(h * w, F, 0.06, seed=10)
    cases = [("twotime_symmetric_none", "symmetric", "None", ()), ("twotime_symmetric_average", "symmetric", "Average", ()),
             ("twotime_staticmap_none", "StaticMap", "None", ()), ("twotime_staticmap_average", "StaticMap", "Average", ()),
             ("twotime_symmetric_framethreading", "symmetric", "None", ("--frame_threading",))]
    for name, method, filt, extra in cases:
        if not wanted(name):
            continue
        res, info = refdrv.run_case(S, dq, sq, F, sparse=(off, idx, val), g2out=False, dpl=8, static_window=20,
                                    twotime=dict(qbins=[1, 3], wsize=10, method=method, filter=filt), extra_args=extra)
        save(name, dict(kind=np.array("twotime"), dq=dq, sq=sq, off=off, idx=idx, val=val,
                        params=np.array([F, 8, 1, 1, 20, 0], np.int64), qbins=np.array([1, 3], np.int32),
                        wsize=np.array(10), filt=np.array(filt), method=np.array(method.lower())), res, info)


ONLY = set(sys.argv[1:])


def wanted(name):
    return not ONLY or name in ONLY


def main():
    if not refdrv.available():
        raise SystemExit("oracle/_ref/corr_ref missing: run `make -C oracle ref` (needs /root/reference)")
    if ONLY:   # python make_golden.py name [name ...]: only these fixtures (the others stay byte-identical)
        twotime_cases()
        if wanted("rigaku_stride2_32x40"):
            rigaku_case("rigaku_stride2_32x40", 32, 40, 600, 0.02, 12, stride=2)
        if wanted("rigaku_avg3_flat_32x40"):
            rigaku_case("rigaku_avg3_flat_32x40", 32, 40, 600, 0.02, 13, avg=3, flat=True)
        if wanted("rigaku_stride2_avg2_32x40"):
            rigaku_case("rigaku_stride2_avg2_32x40", 32, 40, 800, 0.02, 14, stride=2, avg=2)
        if wanted("hdf5_stack_u32_24x32"):
            hdf5_case("hdf5_stack_u32_24x32", 24, 32, 200, 0.03, 15, begin=2, dtype=np.uint32)
        return
    sparse_case("sparse_int_24x24", 24, 24, 600, 0.03, 1)
    sparse_case("sparse_odd_dpl4", 20, 28, 1001, 0.012, 2, dpl=4)
    sparse_case("sparse_staletail_32x32", 32, 32, 4000, 0.004, 3)   # rows hit the lower_bound quirk
    # the hand example of SURVEY.md A.4: one pixel, unit counts at frames 0..31, 400, 440, F = 512
    F = 512
    frames = list(range(32)) + [400, 440]
    off = np.zeros(F + 1, np.int64)
    for f in frames:
        off[f + 1:] += 1
    ev = (np.ones((1, 2), np.int32), np.ones((1, 2), np.int32), off, np.zeros(len(frames), np.int32),
          np.ones(len(frames), np.int16))
    sparse_case("staletail_hand_example", 1, 2, F, 0, 0, events=ev)
    sparse_case("sparse_flat_stride2_avg2", 24, 24, 1200, 0.03, 5, flat=True, stride=2, avg=2)
    sparse_case("sparse_flat_avg3", 24, 24, 900, 0.03, 6, flat=True, avg=3)
    sparse_case("sparse_framesum_norm", 24, 24, 500, 0.04, 7, norm=True)
    ufxc_case("ufxc_wrap_48x40", 48, 40, 400, 0.02, 9)
    rigaku_case("rigaku_compact_32x40", 32, 40, 400, 0.02, 10)
    hdf5_case("hdf5_stack_24x32", 24, 32, 300, 0.03, 11)

    # dense int16 source with dark frames, flat-field and threshold (DenseFilter + DarkImage)
    h = w = 16
    darks, F = 12, 300
    dq, sq = S.annular_qmaps(h, w, n_dynamic=3, static_per_dynamic=2, r_min=1.0)
    fr = S.dense_frames(h * w, F, darks=darks, mu=0.05, seed=8)
    ff = S.flatfield(h * w, seed=108)
    lld, sigma = 5.0, 3.0
    res, info = refdrv.run_case(S, dq, sq, F, dense=fr, g2out=True, darkout=True, dpl=8, darks=darks, lld=lld,
                                sigma=sigma, flatfield=ff, static_window=30)
    save("dense_dark_flat_16x16", dict(kind=np.array("dense"), dq=dq, sq=sq, frames=fr, flat=ff,
                                       params=np.array([F, 8, 1, 1, 30, 0, darks], np.int64),
                                       thresh=np.array([lld, sigma], np.float32)), res, info)
    # dense without darks (threshold 0)
    fr2 = (S.dense_frames(h * w, 200, darks=0, mu=0.05, offset=0.0, read_noise=0.4, seed=9)).astype(np.int16)
    res, info = refdrv.run_case(S, dq, sq, 200, dense=fr2, g2out=True, dpl=8, static_window=20)
    save("dense_nodark_16x16", dict(kind=np.array("dense"), dq=dq, sq=sq, frames=fr2,
                                    params=np.array([200, 8, 1, 1, 20, 0, 0], np.int64)), res, info)

    twotime_cases()
    rigaku_case("rigaku_stride2_32x40", 32, 40, 600, 0.02, 12, stride=2)
    rigaku_case("rigaku_avg3_flat_32x40", 32, 40, 600, 0.02, 13, avg=3, flat=True)
    rigaku_case("rigaku_stride2_avg2_32x40", 32, 40, 800, 0.02, 14, stride=2, avg=2)
    hdf5_case("hdf5_stack_u32_24x32", 24, 32, 200, 0.03, 15, begin=2, dtype=np.uint32)


if __name__ == "__main__":
    main()
