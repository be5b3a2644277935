"""ctypes/numpy front end of the CPU oracle (oracle/xpcs_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package.
Each wrapper names the reference routine (file:line under /root/reference) the C
function it calls restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libxpcs_oracle.so")
    src = os.path.join(_HERE, "xpcs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libxpcs_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    L.xo_level_max.argtypes = [C.c_int, C.c_int]
    L.xo_level_max.restype = C.c_int
    L.xo_delay_schedule.argtypes = [C.c_int, C.c_int, _i32p, _i32p, C.c_int]
    L.xo_delay_schedule.restype = C.c_int
    L.xo_build_qmap.argtypes = [C.c_int, _i32p, _i32p, _i16p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                _i32p, C.c_int, _i32p, _i32p, _i64p, _i32p]
    L.xo_build_qmap.restype = C.c_int
    L.xo_dark_image.argtypes = [C.c_int, C.c_int, _i16p, _f64p, _f64p, _f64p]
    L.xo_dark_image.restype = None
    L.xo_sparse_filter.argtypes = [C.c_int] * 6 + [_i16p, _i32p, _f64p, _i64p, _i32p, _i16p, C.c_int64,
                                                   _i64p, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p]
    L.xo_sparse_filter.restype = C.c_int64
    L.xo_set_late_window.argtypes = [C.c_int]
    L.xo_set_late_window.restype = None
    L.xo_dense_filter.argtypes = [C.c_int] * 6 + [_i16p, _i32p, _f64p, C.c_void_p, C.c_void_p, C.c_float,
                                                  C.c_float, _i16p, C.c_int64, _i64p, _i32p, _f32p, _f32p,
                                                  _f32p, _f32p, _f32p]
    L.xo_dense_filter.restype = C.c_int64
    L.xo_post_scale.argtypes = [C.c_int] * 5 + [_i32p, _i64p, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p]
    L.xo_post_scale.restype = None
    L.xo_multitau.argtypes = [C.c_int, C.c_int, C.c_int, _i64p, _i32p, _f32p, _f32p, _f32p, _f32p,
                              C.c_int, C.c_int]
    L.xo_multitau.restype = None
    L.xo_normalize.argtypes = [C.c_int] * 4 + [_i32p, _i32p, _i64p, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p]
    L.xo_normalize.restype = None
    L.xo_twotime_smooth.argtypes = [C.c_int] * 4 + [_i32p, _i64p, _i32p, C.c_int, _i32p, _i64p, _i32p,
                                                    _f32p, _f32p]
    L.xo_twotime_smooth.restype = C.c_int
    L.xo_twotime_bin.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i64p, _i32p, _f32p, _f32p, _f32p, _f32p]
    L.xo_twotime_bin.restype = None
    L.xo_max_threads.restype = C.c_int
    _LIB = L
    return L


def level_max(frames, dpl):
    """corr.cpp:1156-1160."""
    return lib().xo_level_max(frames, dpl)


def delay_schedule(frames, dpl):
    """corr.cpp:1133-1154 -> (level[T], tau[T])."""
    cap = 64 * (2 * dpl + 2)
    lv = np.zeros(cap, np.int32)
    tv = np.zeros(cap, np.int32)
    n = lib().xo_delay_schedule(frames, dpl, lv, tv, cap)
    return lv[:n].copy(), tv[:n].copy()


class QMap:
    """configuration.cpp:244-381 (BuildQMap) result."""

    def __init__(self, dq, sq):
        dq = np.ascontiguousarray(dq, np.int32).ravel()
        sq = np.ascontiguousarray(sq, np.int32).ravel()
        P = dq.size
        self.P = P
        self.dq, self.sq = dq, sq
        self.mask = np.zeros(P, np.int16)
        smax = int(max(sq.max(), 1))
        self.pixels_per_sbin = np.zeros(smax, np.int32)
        seg_dq = np.zeros(P + 1, np.int32)
        seg_sq = np.zeros(P + 1, np.int32)
        seg_start = np.zeros(P + 2, np.int64)
        seg_pixels = np.zeros(max(P, 1), np.int32)
        S, Q = C.c_int(0), C.c_int(0)
        n = lib().xo_build_qmap(P, dq, sq, self.mask, C.byref(S), C.byref(Q), self.pixels_per_sbin, smax,
                                seg_dq, seg_sq, seg_start, seg_pixels)
        self.S, self.Q, self.nseg = S.value, Q.value, n
        self.pixels_per_sbin = self.pixels_per_sbin[: self.S].copy()
        self.seg_dq = seg_dq[:n].copy()
        self.seg_sq = seg_sq[:n].copy()
        self.seg_start = seg_start[: n + 1].copy()
        self.seg_pixels = seg_pixels[: int(seg_start[n])].copy()
        self.sbin_of_pixel = np.where(self.mask != 0, sq, 0).astype(np.int32)


class Rows:
    """Pixel-major event rows (data_structure/sparse_data.cpp, row.h) as CSR."""

    def __init__(self, row_ptr, t, v):
        self.row_ptr, self.t, self.v = row_ptr, t, v

    def copy(self):
        return Rows(self.row_ptr.copy(), self.t.copy(), self.v.copy())


class FilterOut:
    pass


def _filter_alloc(P, F, S, swindow, cap):
    o = FilterOut()
    o.row_ptr = np.zeros(P + 1, np.int64)
    o.t = np.zeros(max(cap, 1), np.int32)
    o.v = np.zeros(max(cap, 1), np.float32)
    o.pixel_sum = np.zeros(P, np.float32)
    o.frame_sum = np.zeros(2 * F, np.float32)
    o.part_total = np.zeros(max(S, 1), np.float32)
    o.part_partial = np.zeros(max(int(np.ceil(F / swindow)) * S, 1), np.float32)
    return o


def ufxc_frames(words, h, w, nraw):
    """The frames the UFXC reader (io/ufxc.cpp:59-99, 144-153) delivers for raw frames 0 .. nraw-1: a 32-bit word
    carries an 11-bit frame counter (bits 31..21; the first word is frame 0, a jump of more than 2000 between
    consecutive words moves the counter's base by 2048), the count in bits 16..15 and the column-major pixel in
    bits 14..0; events keep their file order inside a frame, a frame without words is empty.
    -> (frame_off int64[nraw + 1], idx int32, val int16)."""
    words = np.asarray(words, np.uint32).astype(np.int64)
    per = [[] for _ in range(nraw)]
    if words.size:
        first = int(words[0] >> 21)
        prev, base = first, 0
        for k, wd in enumerate(words.tolist()):
            c = wd >> 21
            if k > 0:
                d = c - prev
                if d < -2000:
                    base += 2048
                elif d > 2000:
                    base -= 2048
            prev = c
            ff = c + base - first
            if 0 <= ff < nraw:
                pix = wd & 0x7FFF
                per[ff].append(((pix % h) * w + pix // h, (wd >> 15) & 3))
    off = np.zeros(nraw + 1, np.int64)
    off[1:] = np.cumsum([len(p) for p in per])
    flat = [e for p in per for e in p]
    idx = np.asarray([e[0] for e in flat], np.int32)
    val = np.asarray([e[1] for e in flat], np.int16)
    return off, idx, val


def rigaku_frames(words, h, w, frame_start_todo, frames, mask, stride=1, avg=1):
    """The event stream the Rigaku reader (io/rigaku.cpp:139-267) turns into `frames` output frames.  A 64-bit
    word carries the frame in bits 63..40, the column-major pixel in bits 35..16 and the count in bits 10..0.
    Words of frames <= frame_start_todo are skipped; with stride > 1 words of frames that are not a multiple of
    the stride are skipped too (:161-162); an output frame ends when the frame number of a word differs from the
    previous one (avg == 1) or passes the next boundary frame_start_todo + k * block (avg > 1) (:166-168) -- so
    frames without events vanish, except that a run not starting where the reader expects opens with an empty
    output frame; reading stops once `frames` output frames are complete; events on masked pixels are dropped
    after the frame bookkeeping; what is still open at the end of the file is kept only if it holds more than one
    pixel (:232).  The reader sums the words of an output frame per pixel and divides by avg, which is what the
    Filter stage does with a block of `block` raw frames (sparse_filter.cpp:143-172): the output frames come
    back as blocks of `block` = stride * avg (or max) raw frames whose first raw frame holds the words.
    -> (frame_off int64[frames * block + 1], idx int32, val int16) in file order, as for xpcs_push_sparse."""
    words = np.asarray(words, np.uint64)
    block = stride * avg if (stride > 1 and avg > 1) else max(stride, avg)
    off, idx, val = [0], [], []
    prev = frame_start_todo + 1
    nxt = frame_start_todo + block
    done = 0
    cur_idx, cur_val = [], []

    def flush():
        idx.extend(cur_idx)
        val.extend(cur_val)
        off.extend([len(idx)] * block)   # raw frame 0 of the block carries the words, the others are empty
        del cur_idx[:], cur_val[:]

    for wd in words.tolist():
        frame = (wd >> 40) & 0xFFFFFFFF
        if frame <= frame_start_todo:
            continue
        if done >= frames:
            break
        if stride > 1 and frame != 0 and frame % stride != 0:
            continue
        if (avg > 1 and frame > nxt) or (avg == 1 and frame != prev):
            flush()
            prev = frame
            done += 1
            nxt += block
        pix = (wd >> 16) & 0xFFFFF
        pix = (pix % h) * w + pix // h
        if not mask[pix]:
            continue
        cur_idx.append(pix)
        cur_val.append(wd & 0x7FF)
    if done < frames and len(set(cur_idx)) > 1:
        flush()
    while len(off) < frames * block + 1:
        off.append(len(idx))
    return np.asarray(off, np.int64), np.asarray(idx, np.int32), np.asarray(val, np.int16)


def sparse_filter(qm, F, frame_off, idx, val, flat=None, stride=1, avg=1, swindow=1, late_window=False):
    """filter/sparse_filter.cpp:115-193 over the ingest loop main.cpp:258-268.  late_window: the static-window
    rule of the Rigaku reader (io/rigaku.cpp:190-193), whose per-frame sums are otherwise the same."""
    P = qm.P
    flat = np.ones(P, np.float64) if flat is None else np.ascontiguousarray(flat, np.float64).ravel()
    frame_off = np.ascontiguousarray(frame_off, np.int64)
    idx = np.ascontiguousarray(idx, np.int32)
    val = np.ascontiguousarray(val, np.int16)
    cap = int(idx.size)
    o = _filter_alloc(P, F, qm.S, swindow, cap)
    lib().xo_set_late_window(1 if late_window else 0)
    try:
        n = lib().xo_sparse_filter(P, F, stride, avg, swindow, qm.S, qm.mask, qm.sbin_of_pixel, flat, frame_off,
                                   idx if idx.size else np.zeros(1, np.int32),
                                   val if val.size else np.zeros(1, np.int16), cap, o.row_ptr, o.t, o.v,
                                   o.pixel_sum, o.frame_sum, o.part_total, o.part_partial)
    finally:
        lib().xo_set_late_window(0)
    assert n >= 0
    o.rows = Rows(o.row_ptr, o.t[:n], o.v[:n])
    o.n = int(n)
    return o


def dark_image(frames, flat=None):
    """data_structure/dark_image.cpp:81-106 -> (avg[P], std[P]) float64."""
    frames = np.ascontiguousarray(frames, np.int16)
    darks, P = frames.shape[0], int(np.prod(frames.shape[1:]))
    flat = np.ones(P, np.float64) if flat is None else np.ascontiguousarray(flat, np.float64).ravel()
    avg = np.zeros(P, np.float64)
    std = np.zeros(P, np.float64)
    lib().xo_dark_image(P, darks, frames.reshape(darks, P), flat, avg, std)
    return avg, std


def dense_filter(qm, F, frames, flat=None, dark=None, lld=0.0, sigma=0.0, stride=1, avg=1, swindow=1):
    """filter/dense_filter.cpp:121-210; frames = the data frames only, int16 [n_raw][P]."""
    P = qm.P
    frames = np.ascontiguousarray(frames, np.int16).reshape(-1, P)
    flat = np.ones(P, np.float64) if flat is None else np.ascontiguousarray(flat, np.float64).ravel()
    cap = int(frames.shape[0]) * P
    o = _filter_alloc(P, F, qm.S, swindow, cap)
    da = ds = None
    if dark is not None:
        da = np.ascontiguousarray(dark[0], np.float64)
        ds = np.ascontiguousarray(dark[1], np.float64)
    n = lib().xo_dense_filter(P, F, stride, avg, swindow, qm.S, qm.mask, qm.sbin_of_pixel, flat,
                              da.ctypes.data if da is not None else None,
                              ds.ctypes.data if ds is not None else None, lld, sigma, frames, cap, o.row_ptr,
                              o.t, o.v, o.pixel_sum, o.frame_sum, o.part_total, o.part_partial)
    assert n >= 0
    o.rows = Rows(o.row_ptr, o.t[:n], o.v[:n])
    o.n = int(n)
    return o


def post_scale(qm, F, swindow, fo, normalize_by_framesum=False):
    """main.cpp:313-343, :360-378 (in place on the FilterOut)."""
    t = fo.rows.t if fo.rows.t.size else np.zeros(1, np.int32)
    v = fo.rows.v if fo.rows.v.size else np.zeros(1, np.float32)
    lib().xo_post_scale(qm.P, F, qm.S, swindow, int(bool(normalize_by_framesum)), qm.pixels_per_sbin,
                        fo.rows.row_ptr, t, v, fo.pixel_sum, fo.frame_sum, fo.part_total, fo.part_partial)


def multitau(P, F, dpl, rows, compat=True, nthreads=0):
    """corr.cpp:315-431 -> G2, IP, IF each [T][P]; the rows are consumed (mutated copy)."""
    lv, _ = delay_schedule(F, dpl)
    T = lv.size
    r = rows.copy()
    G2 = np.zeros((T, P), np.float32)
    IP = np.zeros((T, P), np.float32)
    IF = np.zeros((T, P), np.float32)
    t = r.t if r.t.size else np.zeros(1, np.int32)
    v = r.v if r.v.size else np.zeros(1, np.float32)
    lib().xo_multitau(P, F, dpl, r.row_ptr, t, v, G2, IP, IF, int(bool(compat)), nthreads)
    return G2, IP, IF


def normalize(qm, G2, IP, IF):
    """corr.cpp:927-1091 -> norm-0-g2, norm-0-stderr each (T, Q)."""
    T, P = G2.shape
    g2 = np.zeros((T, qm.Q), np.float32)
    se = np.zeros((T, qm.Q), np.float32)
    lib().xo_normalize(P, T, qm.Q, qm.nseg, qm.seg_dq, qm.seg_sq, qm.seg_start,
                       qm.seg_pixels if qm.seg_pixels.size else np.zeros(1, np.int32),
                       np.ascontiguousarray(G2), np.ascontiguousarray(IP), np.ascontiguousarray(IF), g2, se)
    return g2, se


def twotime(qm, F, rows, qproc, wsize, method="symmetric", average=False):
    """corr.cpp:781-924 (+ Smoothing :433-560, ComputeSG* :1166-1305).

    Returns dict(sg, C{q: (F,F)}, g2full (F,B), g2partials (wsize, partials, B), bins).
    """
    r = rows.copy()
    static_map = method.lower() == "staticmap"
    qproc = np.ascontiguousarray(qproc, np.int32)
    sg = np.zeros((qm.nseg + 1) * (1 if average else F), np.float32)
    t = r.t if r.t.size else np.zeros(1, np.int32)
    v = r.v if r.v.size else np.zeros(1, np.float32)
    nrows = lib().xo_twotime_smooth(F, int(static_map), int(average), qm.nseg, qm.seg_dq, qm.seg_start,
                                    qm.seg_pixels, qproc.size, qproc, r.row_ptr, t, v, sg)
    sg = sg[: nrows * (1 if average else F)].reshape(nrows, -1)
    bins = sorted(set(int(q) for q in qm.seg_dq if int(q) in set(int(x) for x in qproc)))
    partials = (F - wsize) // wsize
    Cs = {}
    g2full = np.zeros((F, len(bins)), np.float32)
    g2part = np.zeros((wsize, max(partials, 0), len(bins)), np.float32)
    for b, q in enumerate(bins):
        segs = np.nonzero(qm.seg_dq == q)[0]
        pix = np.ascontiguousarray(qm.seg_pixels[qm.seg_start[segs[0]]: qm.seg_start[segs[-1] + 1]])
        Cm = np.zeros((F, F), np.float32)
        gf = np.zeros(F, np.float32)
        gp = np.zeros(max(wsize * partials, 1), np.float32)
        lib().xo_twotime_bin(F, wsize, pix.size, pix, r.row_ptr, t, v, Cm, gf, gp)
        Cs[q] = Cm
        g2full[:, b] = gf
        if partials > 0:
            g2part[:, :, b] = gp[: wsize * partials].reshape(wsize, partials)
    return dict(sg=sg, C=Cs, g2full=g2full, g2partials=g2part, bins=bins)


def max_threads():
    return lib().xo_max_threads()
