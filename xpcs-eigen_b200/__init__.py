"""xpcs-eigen_b200: B200-native XPCS correlation hot path (IMM ingest -> Filter -> multi-tau
G2/IP/IF -> q-bin normalisation -> two-time) behind the C-ABI of include/xpcs_b200.h.

This Python layer is a thin ctypes mirror used by tests/ and bench.py; the reference-facing
host program is the C++ `corr` under host/.  Importing the package does not need a GPU;
creating a Correlator does, and raises XpcsError when there is none (no CPU fallback).

The directory name carries a hyphen; load it with `__graft_entry__.load_package()` (module
name `xpcs_eigen_b200`).
"""
import ctypes as C

import numpy as np

from . import cabi, h5lite, synth, torchio  # noqa: F401
from .cabi import XPCS_COMPAT_STALE_TAIL, XpcsError, XpcsInfo, XpcsParams, XpcsShardPlan  # noqa: F401


def make_params(dqmap, sqmap, frames, dpl=8, flatfield=None, stride=1, avg=1, static_window=None,
                normalize_by_framesum=False, compat=True, lld=0.0, sigma=0.0, shard_index=0, shard_count=1, late_window=False,
                reserve_events=0, lane_multitau=False, scalar_dense=False):
    """XpcsParams from numpy maps; returns (params, keepalive list)."""
    dq = np.ascontiguousarray(dqmap, np.int32)
    sq = np.ascontiguousarray(sqmap, np.int32)
    assert dq.ndim == 2 and dq.shape == sq.shape
    ff = None if flatfield is None else np.ascontiguousarray(flatfield, np.float64).ravel()
    p = XpcsParams()
    p.struct_size = C.sizeof(XpcsParams)
    p.width, p.height = dq.shape[1], dq.shape[0]
    p.frames = int(frames)
    p.delays_per_level = dpl
    p.stride_frames, p.avg_frames = stride, avg
    p.static_window = int(static_window) if static_window else max(1, int(frames) // 10)
    p.normalize_by_framesum = int(bool(normalize_by_framesum))
    p.compat_flags = (XPCS_COMPAT_STALE_TAIL if compat else 0) | (cabi.XPCS_FLAG_LANE_MULTITAU if lane_multitau else 0) | \
        (cabi.XPCS_COMPAT_LATE_WINDOW if late_window else 0) | \
        (cabi.XPCS_FLAG_SCALAR_DENSE if scalar_dense else 0)
    p.lld, p.sigma = float(lld), float(sigma)
    p.dqmap, p.sqmap = dq.ctypes.data, sq.ctypes.data
    p.flatfield = None if ff is None else ff.ctypes.data
    p.shard_index, p.shard_count = shard_index, shard_count
    p.reserve_events = int(reserve_events)
    return p, [dq, sq, ff]


def plan_shard(dqmap, sqmap, frames, shard_index=0, shard_count=1, dpl=8):
    """Host-only: (XpcsShardPlan, row_pixels) of one pixel shard -- what xpcs_create would build."""
    lib = cabi.load()
    p, keep = make_params(dqmap, sqmap, frames, dpl=dpl, shard_index=shard_index, shard_count=shard_count)
    plan = XpcsShardPlan()
    rc = lib.xpcs_plan_shard(C.byref(p), C.byref(plan), None, 0)
    if rc != 0:
        raise XpcsError(rc, (lib.xpcs_last_error(None) or b"").decode())
    rows = np.zeros(max(plan.n_rows, 1), np.int32)
    rc = lib.xpcs_plan_shard(C.byref(p), C.byref(plan), rows.ctypes.data, rows.size)
    if rc != 0:
        raise XpcsError(rc, (lib.xpcs_last_error(None) or b"").decode())
    return plan, rows[: plan.n_rows]


def level_max(frames, dpl):
    """Corr::calculateLevelMax (reference corr.cpp:1156-1160)."""
    return cabi.load().xpcs_level_max(frames, dpl)


def delay_schedule(frames, dpl):
    """Corr::delaysPerLevel (reference corr.cpp:1133-1154) -> (level[T], tau[T])."""
    lib = cabi.load()
    n = lib.xpcs_delay_schedule(frames, dpl, None, None, 0)
    lv = np.zeros(max(n, 1), np.int32)
    tv = np.zeros(max(n, 1), np.int32)
    lib.xpcs_delay_schedule(frames, dpl, lv.ctypes.data, tv.ctypes.data, n)
    return lv[:n], tv[:n]


def comm_unique_id():
    """ncclGetUniqueId through the library: 128 bytes for Correlator.comm_init on every rank."""
    buf = C.create_string_buffer(128)
    rc = cabi.load().xpcs_comm_unique_id(buf)
    if rc != 0:
        raise XpcsError(rc, (cabi.load().xpcs_last_error(None) or b"").decode())
    return buf.raw


def _ptr(a):
    return None if a is None else a.ctypes.data


class Correlator:
    """One handle = one GPU = one pixel shard.  Method names follow the C-ABI, which in turn
    names the reference routine each call replaces (see include/xpcs_b200.h)."""

    def __init__(self, dqmap, sqmap, frames, dpl=8, flatfield=None, stride=1, avg=1, static_window=None,
                 normalize_by_framesum=False, compat=True, lld=0.0, sigma=0.0, device=0, shard_index=0,
                 shard_count=1, reserve_events=0, lane_multitau=False, scalar_dense=False, late_window=False):
        self._lib = cabi.load()
        p, self._maps = make_params(dqmap, sqmap, frames, dpl=dpl, flatfield=flatfield, stride=stride, avg=avg,
                                    static_window=static_window, normalize_by_framesum=normalize_by_framesum,
                                    compat=compat, lld=lld, sigma=sigma, shard_index=shard_index,
                                    shard_count=shard_count, reserve_events=reserve_events,
                                    lane_multitau=lane_multitau, scalar_dense=scalar_dense, late_window=late_window)
        self.height, self.width = p.height, p.width
        self.P = p.width * p.height
        self.F = int(frames)
        self.params = p
        self.static_window = p.static_window
        h = C.c_void_p()
        rc = self._lib.xpcs_create(C.byref(p), device, C.byref(h))
        if rc != 0:
            raise XpcsError(rc, (self._lib.xpcs_last_error(None) or b"").decode())
        self._h = h
        self._keep = []  # host buffers that must outlive asynchronous copies
        self._pinned = {}  # name -> (pointer, numpy view): persistent page-locked result buffers (pinned_results())
        i = self.info()
        self.T, self.S, self.Q = i.n_delays, i.n_static, i.n_dynamic

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise XpcsError(rc, (self._lib.xpcs_last_error(self._h) or b"").decode())

    def pinned_results(self, on=True):
        """Keep the result arrays of finish_ingest() / normalize() in page-locked buffers owned by this object
        (allocated once, reused by every call: the arrays returned are views, valid until the next call)."""
        self._use_pinned = bool(on)

    def _out(self, name, n, dtype=np.float32):
        if not getattr(self, "_use_pinned", False):
            return np.empty(n, dtype)
        have = self._pinned.get(name)
        nbytes = int(n) * np.dtype(dtype).itemsize
        if have is None or have[2] < nbytes:
            if have is not None:
                self._lib.xpcs_host_free(have[0])
            p = self._lib.xpcs_host_alloc(max(nbytes, 1))
            if not p:
                raise XpcsError(-4, "xpcs_host_alloc(%d) failed" % nbytes)
            self._pinned[name] = have = (p, (C.c_char * max(nbytes, 1)).from_address(p), nbytes)
        return np.frombuffer(have[1], dtype=dtype, count=int(n))

    def close(self):
        for p, _, _ in getattr(self, "_pinned", {}).values():
            self._lib.xpcs_host_free(p)
        self._pinned = {}
        if getattr(self, "_h", None):
            self._lib.xpcs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = XpcsInfo()
        self._check(self._lib.xpcs_get_info(self._h, C.byref(i)))
        return i

    def row_pixels(self):
        out = np.zeros(max(self.info().n_rows, 1), np.int32)
        self._check(self._lib.xpcs_get_row_pixels(self._h, out.ctypes.data))
        return out[: self.info().n_rows]

    def set_stream(self, cuda_stream):
        self._check(self._lib.xpcs_set_stream(self._h, C.c_void_p(cuda_stream)))

    def reset(self):
        self._keep = []
        self._check(self._lib.xpcs_reset(self._h))

    # -- Filter stage --
    def set_dark(self, frames):
        f = np.ascontiguousarray(frames, np.int16).reshape(-1, self.P)
        self._check(self._lib.xpcs_set_dark(self._h, f.ctypes.data, f.shape[0]))

    def get_dark(self):
        avg = np.zeros(self.P, np.float64)
        std = np.zeros(self.P, np.float64)
        self._check(self._lib.xpcs_get_dark(self._h, avg.ctypes.data, std.ctypes.data))
        return avg, std

    def push_sparse(self, idx, val, frame_off, clock=None, ticks=None):
        idx = np.ascontiguousarray(idx, np.int32)
        val = np.ascontiguousarray(val, np.int16)
        off = np.ascontiguousarray(frame_off, np.int64)
        ck = None if clock is None else np.ascontiguousarray(clock, np.float64)
        tk = None if ticks is None else np.ascontiguousarray(ticks, np.float64)
        self._keep += [idx, val]
        self._check(self._lib.xpcs_push_sparse(self._h, idx.ctypes.data, val.ctypes.data, off.ctypes.data,
                                               _ptr(ck), _ptr(tk), off.size - 1))

    def push_sparse_raw(self, idx_ptr, val_ptr, off_ptr, nframes):
        """Host pointers as integers (e.g. pinned torch tensors); caller keeps them alive."""
        self._check(self._lib.xpcs_push_sparse(self._h, idx_ptr, val_ptr, off_ptr, None, None, nframes))

    def push_sparse_device(self, d_idx, d_val, d_off, n_events, nframes):
        """Device pointers (integers); buffers stay valid until finish_ingest returns."""
        self._check(self._lib.xpcs_push_sparse_device(self._h, d_idx, d_val, d_off, n_events, nframes))

    # -- multi-GPU (one Correlator per GPU; comm.cu) --
    def comm_init(self, nranks, rank, unique_id):
        """Collective over all ranks: joins the NCCL communicator made from `unique_id` (128 bytes from
        comm_unique_id() on rank 0)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.xpcs_comm_init(self._h, nranks, rank, buf))

    def comm_transport(self):
        """1 = direct NVLink stores into the owners' buffers, 0 = staged ncclSend/ncclRecv, -1 = no communicator."""
        return int(self._lib.xpcs_comm_transport(self._h))

    def push_sparse_slab(self, first_raw_frame, idx, val, frame_off, clock=None, ticks=None):
        """Raw frames [first_raw_frame, +len(frame_off)-1) of the WHOLE detector; finish_ingest redistributes."""
        idx = np.ascontiguousarray(idx, np.int32)
        val = np.ascontiguousarray(val, np.int16)
        off = np.ascontiguousarray(frame_off, np.int64)
        ck = None if clock is None else np.ascontiguousarray(clock, np.float64)
        tk = None if ticks is None else np.ascontiguousarray(ticks, np.float64)
        self._keep += [idx, val, off]
        self._check(self._lib.xpcs_push_sparse_slab(self._h, first_raw_frame, idx.ctypes.data, val.ctypes.data,
                                                    off.ctypes.data, _ptr(ck), _ptr(tk), off.size - 1))

    def push_sparse_slab_raw(self, first_raw_frame, idx_ptr, val_ptr, off_ptr, nframes):
        """Host pointers as integers (pinned torch tensors); caller keeps them alive until finish_ingest."""
        self._check(self._lib.xpcs_push_sparse_slab(self._h, first_raw_frame, idx_ptr, val_ptr, off_ptr, None, None, nframes))

    def push_sparse_slab_device(self, first_raw_frame, d_idx, d_val, d_off, n_events, nframes):
        self._check(self._lib.xpcs_push_sparse_slab_device(self._h, first_raw_frame, d_idx, d_val, d_off, n_events, nframes))

    def push_dense(self, frames, clock=None, ticks=None):
        f = np.ascontiguousarray(frames, np.int16).reshape(-1, self.P)
        ck = None if clock is None else np.ascontiguousarray(clock, np.float64)
        tk = None if ticks is None else np.ascontiguousarray(ticks, np.float64)
        self._check(self._lib.xpcs_push_dense(self._h, f.ctypes.data, _ptr(ck), _ptr(tk), f.shape[0]))

    def push_dense_raw(self, frames_ptr, nframes):
        """Host pointer as an integer (e.g. a pinned torch tensor); caller keeps it alive."""
        self._check(self._lib.xpcs_push_dense(self._h, frames_ptr, None, None, nframes))

    def push_dense_device(self, d_frames, nframes):
        self._check(self._lib.xpcs_push_dense_device(self._h, d_frames, nframes))

    # -- online multi-tau (multitau_stream.cu): frame streams that do not fit the device --
    def stream_begin(self, chunk_frames):
        """Chunks of chunk_frames = 2^k frames (64..8192) follow through stream_push_sparse; stream_finish() ends the
        stream, after which multitau() / normalize() hand out the results as after a resident ingest."""
        self._check(self._lib.xpcs_stream_begin(self._h, int(chunk_frames)))

    def stream_push_sparse(self, idx, val, frame_off, clock=None, ticks=None):
        idx = np.ascontiguousarray(idx, np.int32)
        val = np.ascontiguousarray(val, np.int16)
        off = np.ascontiguousarray(frame_off, np.int64)
        ck = None if clock is None else np.ascontiguousarray(clock, np.float64)
        tk = None if ticks is None else np.ascontiguousarray(ticks, np.float64)
        self._check(self._lib.xpcs_stream_push_sparse(self._h, idx.ctypes.data, val.ctypes.data, off.ctypes.data,
                                                      _ptr(ck), _ptr(tk), off.size - 1))

    def stream_push_sparse_device(self, d_idx, d_val, d_off, n_events, nframes):
        """Device pointers (integers), one chunk, offsets starting at 0; the buffers are free again on return."""
        self._check(self._lib.xpcs_stream_push_sparse_device(self._h, d_idx, d_val, d_off, n_events, nframes))

    def stream_finish(self, want=True):
        return self.finish_ingest(want, _fn=self._lib.xpcs_stream_finish)

    def finish_ingest(self, want=True, _fn=None):
        """-> dict(pixel_sum (h,w), frame_sum (2,F), part_total (S,), part_partial (F//window, S))."""
        fn = _fn or self._lib.xpcs_finish_ingest
        if not want:
            self._check(fn(self._h, None, None, None, None))
            self._keep = []
            return None
        F, S = self.F, self.S
        ps = self._out("pixel_sum", self.P)   # fully written by the library
        fs = self._out("frame_sum", 2 * F)
        pt = self._out("part_total", max(S, 1))
        pp = self._out("part_partial", max((F // self.static_window) * S, 1))
        pt[:] = 0
        pp[:] = 0
        self._check(fn(self._h, ps.ctypes.data, fs.ctypes.data, pt.ctypes.data, pp.ctypes.data))
        self._keep = []
        return dict(pixel_sum=ps.reshape(self.height, self.width), frame_sum=fs.reshape(2, F),
                    part_total=pt[:S], part_partial=pp[: (F // self.static_window) * S].reshape(-1, S))

    def frames(self, n):
        """First n post-filter frames, (n, P) float32 (the --frameout dump, main.cpp:276-310)."""
        out = np.zeros((n, self.P), np.float32)
        self._check(self._lib.xpcs_get_frames(self._h, n, out.ctypes.data))
        return out

    def timestamps(self):
        n = self.info().raw_frames_seen
        ck = np.zeros(2 * n, np.float64)
        tk = np.zeros(2 * n, np.float64)
        self._check(self._lib.xpcs_get_timestamps(self._h, ck.ctypes.data, tk.ctypes.data))
        return ck.reshape(2, n), tk.reshape(2, n)

    # -- Correlation --
    def multitau(self, want=True):
        """Corr::multiTau2 -> (G2, IP, IF) each (T, P), or None when want is False."""
        if not want:
            self._check(self._lib.xpcs_multitau(self._h, None, None, None))
            return None
        out = [np.zeros((self.T, self.P), np.float32) for _ in range(3)]
        self._check(self._lib.xpcs_multitau(self._h, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data))
        return tuple(out)

    def correlators(self, pixels):
        """G2, IP, IF of the listed detector pixels, each (T, n) -- the columns of multitau()'s arrays."""
        pix = np.ascontiguousarray(pixels, np.int32)
        out = [np.zeros((self.T, pix.size), np.float32) for _ in range(3)]
        self._check(self._lib.xpcs_get_correlators(self._h, pix.ctypes.data, pix.size, out[0].ctypes.data,
                                                   out[1].ctypes.data, out[2].ctypes.data))
        return tuple(out)

    def normalize(self):
        """Corr::normalizeG2s -> (g2, stderr) each (T, Q)."""
        g2 = self._out("g2", self.T * max(self.Q, 1)).reshape(self.T, max(self.Q, 1))
        se = self._out("se", self.T * max(self.Q, 1)).reshape(self.T, max(self.Q, 1))
        g2[:] = 0
        se[:] = 0
        self._check(self._lib.xpcs_normalize(self._h, g2.ctypes.data, se.ctypes.data))
        return g2[:, : self.Q], se[:, : self.Q]

    def normalize_partials(self):
        """-> (device pointer, count of float64) for the cross-shard SUM all-reduce."""
        p = C.c_void_p()
        n = C.c_int64()
        self._check(self._lib.xpcs_normalize_partials(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def normalize_finish(self):
        g2 = np.zeros((self.T, max(self.Q, 1)), np.float32)
        se = np.zeros((self.T, max(self.Q, 1)), np.float32)
        self._check(self._lib.xpcs_normalize_finish(self._h, g2.ctypes.data, se.ctypes.data))
        return g2[:, : self.Q], se[:, : self.Q]

    def twotime(self, qbin, wsize, method="symmetric", average=False, want_c=True):
        """Corr::twotime for one dynamic bin.  method: "none" | "symmetric" | "staticmap".
        sg comes back as (rows, F or 1): one row for symmetric, one per static partition for staticmap."""
        F = self.F
        partials = max((F - wsize) // wsize, 0)
        Cm = np.zeros((F, F), np.float32) if want_c else None
        gf = np.zeros(F, np.float32)
        gp = np.zeros(max(wsize * partials, 1), np.float32)
        m = {"none": 0, "symmetric": 1, "staticmap": 2}[method.lower()]
        cols = 1 if average else F
        sg = np.zeros((max(self.S, 1) if m == 2 else 1) * cols, np.float32)
        rows = C.c_int(0)
        self._check(self._lib.xpcs_twotime_sg(self._h, qbin, wsize, m, int(bool(average)), _ptr(Cm),
                                              gf.ctypes.data, gp.ctypes.data, sg.ctypes.data, C.byref(rows)))
        sg = sg[: rows.value * cols].reshape(rows.value, cols)
        return dict(C=Cm, g2full=gf, g2partials=gp[: wsize * partials].reshape(wsize, partials),
                    sg=sg if m == 2 else sg.ravel())

    # -- measurement --
    def kernel_timing(self, on=True):
        self._check(self._lib.xpcs_kernel_timing(self._h, int(on)))

    def multitau_fallback_slices(self):
        """slices the warp-per-row multi-tau kernel left to the lane-per-row one (-1: it did not run)."""
        return int(self._lib.xpcs_multitau_fallback_slices(self._h))

    def launch_count(self):
        return int(self._lib.xpcs_launch_count(self._h))

    def kernel_report(self, reset=False):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        ln = (C.c_int64 * cap)()
        n = self._lib.xpcs_kernel_report(self._h, names, ms, ln, cap)
        out = {names[i].decode(): (ms[i], ln[i]) for i in range(min(n, cap))}
        if reset:
            self._check(self._lib.xpcs_kernel_report_reset(self._h))
        return out
