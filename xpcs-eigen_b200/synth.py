"""Deterministic synthetic XPCS inputs: partition maps, sparse and
dense IMM frame streams, and an IMM file writer/reader following the on-disk layout the
reference consumes (io/imm.h:63-144 header, io/imm.cpp:70-118 payload).

Host-side utility only (numpy); nothing here computes correlations.
"""
import struct

import numpy as np

IMM_HEADER_BYTES = 1024
# offsets measured with offsetof() on the reference's Header
_OFF_MODE, _OFF_COMPRESSION, _OFF_ROWS, _OFF_COLS, _OFF_BYTES = 0, 4, 108, 112, 116
_OFF_ELAPSED, _OFF_PRESET, _OFF_DLEN, _OFF_BUFNO, _OFF_IMMVER, _OFF_COREC = 128, 136, 152, 160, 616, 620


def annular_qmaps(h, w, n_dynamic=36, static_per_dynamic=10, r_min=8.0, r_max=None):
    """Concentric annuli: dq in 1..Q over r in [r_min, r_max), each split radially into
    `static_per_dynamic` nested static bins (sq in 1..S); 0 outside."""
    r_max = min(h, w) / 2.0 if r_max is None else r_max
    y, x = np.mgrid[0:h, 0:w]
    r = np.hypot(y - (h - 1) / 2.0, x - (w - 1) / 2.0)
    S = n_dynamic * static_per_dynamic
    s = np.floor((r - r_min) / (r_max - r_min) * S).astype(np.int64)
    valid = (r >= r_min) & (s >= 0) & (s < S)
    sq = np.where(valid, s + 1, 0).astype(np.int32)
    dq = np.where(valid, s // static_per_dynamic + 1, 0).astype(np.int32)
    return dq, sq


def sparse_frames(P, F, occupancy, seed=1234, frames_per_block=None, mean_extra=0.1):
    """Sparse frame stream: every (frame, pixel) cell fires with probability `occupancy`
    (geometric gaps over the flattened cell index), count = 1 + Poisson(mean_extra).
    Returns frame_off int64[F+1], idx int32[E] (ascending within a frame), val int16[E]."""
    if frames_per_block is None:
        frames_per_block = max(1, int(4e6 / max(P * occupancy, 1e-9)))
    frame_off = np.zeros(F + 1, np.int64)
    idx_parts, val_parts = [], []
    for b0 in range(0, F, frames_per_block):
        nb = min(frames_per_block, F - b0)
        rng = np.random.default_rng([seed, b0])
        cells = P * nb
        expect = cells * occupancy
        n = int(expect + 6.0 * np.sqrt(expect + 1.0) + 16)
        pos = np.cumsum(rng.geometric(occupancy, n).astype(np.int64)) - 1
        while pos.size and pos[-1] < cells:  # practically never
            more = np.cumsum(rng.geometric(occupancy, n).astype(np.int64)) + pos[-1]
            pos = np.concatenate([pos, more])
        pos = pos[pos < cells]
        fr = pos // P
        idx_parts.append((pos - fr * P).astype(np.int32))
        val_parts.append((1 + rng.poisson(mean_extra, pos.size)).clip(1, 32767).astype(np.int16))
        cnt = np.bincount(fr, minlength=nb)
        frame_off[b0 + 1: b0 + nb + 1] = cnt
    frame_off = np.cumsum(frame_off)
    idx = np.concatenate(idx_parts) if idx_parts else np.zeros(0, np.int32)
    val = np.concatenate(val_parts) if val_parts else np.zeros(0, np.int16)
    return frame_off, idx, val


def dense_frames(P, F, darks=0, mu=0.3, adu=20.0, offset=100.0, read_noise=2.0, seed=1234):
    """Dense int16 stream [darks+F][P]: dark frames = offset + N(0, read_noise); data frames
    add adu * Poisson(mu) photons."""
    rng = np.random.default_rng(seed)
    out = np.empty((darks + F, P), np.int16)
    for f in range(darks + F):
        fr = offset + read_noise * rng.standard_normal(P)
        if f >= darks:
            fr = fr + adu * rng.poisson(mu, P)
        out[f] = np.clip(np.rint(fr), -32768, 32767).astype(np.int16)
    return out


def flatfield(P, sigma=0.05, seed=7):
    rng = np.random.default_rng(seed)
    return (1.0 + sigma * rng.standard_normal(P)).astype(np.float64)


def _header(compression, rows, cols, dlen, frame_no, elapsed, corecotick):
    h = bytearray(IMM_HEADER_BYTES)
    struct.pack_into("<i", h, _OFF_MODE, 2)
    struct.pack_into("<i", h, _OFF_COMPRESSION, compression)
    struct.pack_into("<i", h, _OFF_ROWS, rows)
    struct.pack_into("<i", h, _OFF_COLS, cols)
    struct.pack_into("<i", h, _OFF_BYTES, 2)
    struct.pack_into("<d", h, _OFF_ELAPSED, elapsed)
    struct.pack_into("<d", h, _OFF_PRESET, 1e-3)
    struct.pack_into("<I", h, _OFF_DLEN, dlen)
    struct.pack_into("<i", h, _OFF_BUFNO, frame_no)
    struct.pack_into("<i", h, _OFF_IMMVER, 12)
    struct.pack_into("<i", h, _OFF_COREC, corecotick)
    h[1012:1024] = b"\xff" * 12
    return bytes(h)


def frame_clock(n_frames, dt=1e-3):
    """(elapsed f64[n], corecotick i32[n]) written into / expected from the headers."""
    k = np.arange(n_frames)
    return (k + 1) * dt, (1000 + 7 * k).astype(np.int32)


def write_imm_sparse(path, h, w, frame_off, idx, val, dt=1e-3):
    elapsed, tick = frame_clock(len(frame_off) - 1, dt)
    with open(path, "wb") as fh:
        for f in range(len(frame_off) - 1):
            a, b = int(frame_off[f]), int(frame_off[f + 1])
            fh.write(_header(6, h, w, b - a, f, float(elapsed[f]), int(tick[f])))
            fh.write(np.ascontiguousarray(idx[a:b], "<i4").tobytes())
            fh.write(np.ascontiguousarray(val[a:b], "<i2").tobytes())


def ufxc_words(h, w, frame_off, idx, val, f0=0):
    """Events -> the 32-bit words of a UFXC file (reference io/ufxc.cpp:59-99, 144-153): bits 31..21 the
    frame counter (11 bits, wraps at 2048; the first word's counter is frame 0), bits 16..15 the count
    (0..3), bits 14..0 the pixel in column-major order (index = (pix % h) * w + pix // h)."""
    assert h * w <= 1 << 15
    idx = np.asarray(idx, np.int64)
    fr = np.repeat(np.arange(len(frame_off) - 1, dtype=np.int64), np.diff(frame_off))
    pix = (idx % w) * h + idx // w
    v = np.asarray(val, np.int64)
    assert v.min(initial=0) >= 0 and v.max(initial=0) <= 3
    return ((((f0 + fr) & 0x7FF) << 21) | (v << 15) | pix).astype("<u4")


def rigaku_words(h, w, frame_ids, frame_off, idx, val):
    """Events -> the 64-bit words of a Rigaku file (reference io/rigaku.cpp:143, 210-217): frame number in
    bits 63..40 (frame_ids[f] for the events of frame f), column-major pixel in bits 35..16, count (0..2047)
    in bits 10..0."""
    idx = np.asarray(idx, np.int64)
    fr = np.repeat(np.asarray(frame_ids, np.int64), np.diff(frame_off))
    pix = (idx % w) * h + idx // w
    v = np.asarray(val, np.int64)
    assert v.min(initial=0) >= 0 and v.max(initial=0) <= 0x7FF and h * w <= 1 << 20
    return ((fr.astype(np.uint64) << np.uint64(40)) | (pix.astype(np.uint64) << np.uint64(16)) | v.astype(np.uint64)).astype("<u8")


def write_ufxc(path, h, w, frame_off, idx, val, f0=0):
    ufxc_words(h, w, frame_off, idx, val, f0).tofile(path)


def write_imm_dense(path, h, w, frames, dt=1e-3):
    frames = np.asarray(frames).reshape(-1, h * w)
    elapsed, tick = frame_clock(frames.shape[0], dt)
    with open(path, "wb") as fh:
        for f in range(frames.shape[0]):
            fh.write(_header(0, h, w, h * w, f, float(elapsed[f]), int(tick[f])))
            fh.write(np.ascontiguousarray(frames[f], "<i2").tobytes())


def read_imm(path):
    """Parse an IMM file the way io/imm.cpp does (compression flag from the first header).
    Returns dict(sparse, frame_off, idx, val | frames, elapsed, corecotick)."""
    buf = np.fromfile(path, np.uint8)
    n = buf.size
    pos, first = 0, True
    sparse = False
    offs, idxs, vals, el, ct = [0], [], [], [], []
    while pos + IMM_HEADER_BYTES <= n:
        hb = buf[pos: pos + IMM_HEADER_BYTES].tobytes()
        if first:
            sparse = struct.unpack_from("<i", hb, _OFF_COMPRESSION)[0] != 0
            first = False
        dlen = struct.unpack_from("<I", hb, _OFF_DLEN)[0]
        el.append(struct.unpack_from("<d", hb, _OFF_ELAPSED)[0])
        ct.append(struct.unpack_from("<i", hb, _OFF_COREC)[0])
        pos += IMM_HEADER_BYTES
        if sparse:
            idxs.append(buf[pos: pos + 4 * dlen].view("<i4"))
            pos += 4 * dlen
        vals.append(buf[pos: pos + 2 * dlen].view("<i2"))
        pos += 2 * dlen
        offs.append(offs[-1] + dlen)
    out = dict(sparse=sparse, elapsed=np.array(el), corecotick=np.array(ct, np.float64))
    if sparse:
        out.update(frame_off=np.array(offs, np.int64),
                   idx=np.concatenate(idxs).astype(np.int32) if idxs else np.zeros(0, np.int32),
                   val=np.concatenate(vals).astype(np.int16) if vals else np.zeros(0, np.int16))
    else:
        out.update(frames=np.stack(vals).astype(np.int16) if vals else np.zeros((0, 0), np.int16))
    return out
