"""Driver for oracle/_ref/corr_ref -- the UNMODIFIED reference `corr` compiled from
/root/reference by oracle/ref/Makefile against the directory-backed HDF5 shim
(oracle/ref/h5dir.cpp).

TEST INFRASTRUCTURE ONLY (see oracle/xpcs_oracle.c header for who may import oracle/).
It writes a configuration container with the keys of Configuration::init
(configuration.cpp:80-242; on-disk types as SURVEY.md B.1: int32 for getInteger keys, float32
for getFloat keys, int64 for getLong keys, fixed strings), an IMM file
(xpcs-eigen_b200/synth.py follows io/imm.h:63-144), runs the binary the way a user runs
`corr config.hdf5 --imm data.imm [--g2out] [--darkout]`, and reads the result datasets back.
"""
import os
import re
import shutil
import struct
import subprocess
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "_ref", "corr_ref")

_TCODE = {np.dtype("<i4"): 1, np.dtype("<i8"): 2, np.dtype("<f4"): 3, np.dtype("<f8"): 4, np.dtype("<u4"): 5,
          np.dtype("<u2"): 6, np.dtype("<i2"): 7, np.dtype("<u8"): 8}
_DTYPE = {v: k for k, v in _TCODE.items()}


def available():
    return os.path.exists(BIN) and os.access(BIN, os.X_OK)


# ---- the container layout of oracle/ref/h5dir.cpp ----
def put(root, path, value):
    """Write one dataset.  str -> fixed-length string; numpy array -> its dtype, rank <= 4."""
    full = os.path.join(root, path.strip("/"))
    os.makedirs(os.path.dirname(full), exist_ok=True)
    if isinstance(value, str):
        raw = value.encode()
        tcode, esize, dims = 9, len(raw), [1]
    else:
        a = np.ascontiguousarray(value)
        tcode, esize, dims = _TCODE[a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else np.dtype(a.dtype.str.replace("=", "<").replace("|", "<"))], a.dtype.itemsize, list(a.shape)
        raw = a.tobytes()
    d4 = dims + [0] * (4 - len(dims))
    hdr = b"H5DIRDS1" + struct.pack("<IIII4QQ", tcode, esize, len(dims), 0, *d4, len(raw))
    assert len(hdr) == 64
    with open(full, "wb") as fh:
        fh.write(hdr)
        fh.write(raw)


def get(root, path):
    full = os.path.join(root, path.strip("/"))
    with open(full, "rb") as fh:
        hdr = fh.read(64)
        assert hdr[:8] == b"H5DIRDS1", full
        tcode, esize, rank, _ = struct.unpack_from("<IIII", hdr, 8)
        dims = struct.unpack_from("<4Q", hdr, 24)[:rank]
        raw = fh.read()
    if tcode == 9:
        return raw.decode()
    return np.frombuffer(raw, _DTYPE[tcode]).reshape(dims).copy()


def listing(root, group):
    g = os.path.join(root, group.strip("/"))
    out = {}
    for dirpath, _, files in os.walk(g):
        for f in files:
            rel = os.path.relpath(os.path.join(dirpath, f), g)
            out[rel] = get(root, os.path.join(group, rel))
    return out


def config_items(dq, sq, frames, imm_path, dpl=8, begin=1, darks=0, lld=0.0, sigma=0.0, stride=1, avg=1,
                 static_window=None, flatfield=None, normalize_by_framesum=False, twotime=None, entry="/xpcs",
                 output="/exchange"):
    """[(dataset path, value)] of one configuration: every key Configuration::init reads
    (configuration.cpp:80-242) with the on-disk types of SURVEY.md B.1.
    frames = number of RAW data frames to do (data_begin_todo = begin, 1-based, after the darks).
    twotime = dict(qbins=[...], wsize=int, method="symmetric", filter="None") or None."""
    dq = np.ascontiguousarray(dq, "<i4")
    sq = np.ascontiguousarray(sq, "<i4")
    h, w = dq.shape
    e = entry
    i32 = lambda v: np.array([[v]], "<i4")  # noqa: E731
    f32 = lambda v: np.array([[v]], "<f4")  # noqa: E731
    i64 = lambda v: np.array([[v]], "<i8")  # noqa: E731
    first = begin + darks
    block = stride * avg if (stride > 1 and avg > 1) else max(stride, avg)
    F = frames // block
    it = [
        (e + "/compression", "ENABLED"),
        (e + "/output_data", output),
        ("/measurement/instrument/detector/x_dimension", i32(w)),
        ("/measurement/instrument/detector/y_dimension", i32(h)),
        (e + "/dqmap", dq),
        (e + "/sqmap", sq),
        (e + "/data_begin", i32(first)),
        (e + "/data_end", i32(first + frames - 1)),
        (e + "/data_begin_todo", i32(first)),
        (e + "/data_end_todo", i32(first + frames - 1)),
        (e + "/delays_per_level", i32(dpl)),
        (e + "/dark_begin_todo", i32(1 if darks else 0)),
        (e + "/dark_end_todo", i32(darks if darks else 0)),
        (e + "/lld", f32(lld)),
        (e + "/sigma", f32(sigma)),
        (e + "/stride_frames", i64(stride)),
        (e + "/avg_frames", i64(avg)),
        (e + "/normalize_by_framesum", i32(1 if normalize_by_framesum else 0)),
    ]
    for k, v in (("x_pixel_size", 7.5e-5), ("y_pixel_size", 7.5e-5), ("adu_per_photon", 1.0), ("exposure_time", 1e-3),
                 ("efficiency", 1.0), ("distance", 4.0)):
        it.append(("/measurement/instrument/detector/" + k, f32(v)))
    it.append(("/measurement/instrument/source_begin/beam_intensity_transmitted", f32(1e10)))
    it.append(("/measurement/sample/thickness", f32(1.0)))
    it.append((e + "/static_mean_window_size", i32(static_window or max(1, F // 10))))
    if flatfield is not None:
        it.append((e + "/flatfield_enabled", "ENABLED"))
        it.append(("/measurement/instrument/detector/flatfield", np.ascontiguousarray(flatfield, "<f8").reshape(h, w)))
    else:
        it.append((e + "/flatfield_enabled", "DISABLED"))
    if twotime:
        it.append((e + "/analysis_type", "Twotime"))
        it.append((e + "/smoothing_method", twotime.get("method", "symmetric")))
        it.append((e + "/smoothing_filter", twotime.get("filter", "None")))
        it.append((e + "/qphi_bin_to_process", np.asarray(twotime["qbins"], "<i8").reshape(-1, 1)))
        it.append((e + "/twotime2onetime_window_size", i32(twotime["wsize"])))
    else:
        it.append((e + "/analysis_type", "Multitau"))
        it.append((e + "/twotime2onetime_window_size", i32(1)))
    it.append((e + "/input_file_local", imm_path))
    return it, F


def write_config(root, dq, sq, frames, imm_path, **kw):
    """The configuration as a directory-backed container (oracle/ref/h5dir.cpp) for corr_ref."""
    if os.path.exists(root):
        shutil.rmtree(root)
    os.makedirs(root)
    items, F = config_items(dq, sq, frames, imm_path, **kw)
    for path, value in items:
        put(root, path, value)
    return F


_SCOPE = re.compile(r"\[info\] (.+?) took (\d+(?:\.\d+)?)\s*(ms|s|m)\b")


def run(root, imm_path=None, g2out=False, darkout=False, threads=None, extra=(), cwd=None):
    """Run corr_ref; returns dict(seconds, scopes{name: seconds}, stdout)."""
    if not available():
        raise RuntimeError("oracle/_ref/corr_ref not built (make -C oracle ref needs /root/reference)")
    cmd = [BIN, root]
    if imm_path:
        cmd.append("--imm=" + imm_path)
    if g2out:
        cmd.append("--g2out")
    if darkout:
        cmd.append("--darkout")
    cmd += list(extra)
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd or os.path.dirname(root), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("corr_ref failed (%d):\n%s" % (p.returncode, p.stdout[-2000:]))
    scopes = {}
    for name, val, unit in _SCOPE.findall(p.stdout):
        scopes[name.strip()] = float(val) * {"ms": 1e-3, "s": 1.0, "m": 60.0}[unit]
    return dict(seconds=dt, scopes=scopes, stdout=p.stdout)


def scratch_dir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix="xpcs_ref_", dir=base)


def run_case(synth, dq, sq, frames, sparse=None, dense=None, g2out=True, darkout=False, threads=None, keep=False,
             ufxc=None, rigaku=None, hdf5=None, extra_args=(), **cfg):
    """Write IMM + config, run the reference, return (results dict of the output group, run info).
    sparse = (frame_off, idx, val); dense = int16 frames [darks + frames][P]."""
    d = scratch_dir()
    try:
        imm = os.path.join(d, "data.imm")
        h, w = np.asarray(dq).shape
        extra = ()
        if ufxc is not None:  # raw 32-bit words of a UFXC file, read through --ufxc
            np.asarray(ufxc, "<u4").tofile(imm)
            extra = ("--ufxc",)
        elif hdf5 is not None:  # frame stack [frames][h][w] (uint16 / uint32) at /entry/data/data, read through --hdf5
            put(imm, "/entry/data/data", np.ascontiguousarray(hdf5))
            extra = ("--hdf5",)
        elif rigaku is not None:  # raw 64-bit words of a Rigaku file, read through --rigaku
            np.asarray(rigaku, "<u8").tofile(imm)
            extra = ("--rigaku",)
        elif sparse is not None:
            synth.write_imm_sparse(imm, h, w, *sparse)
        else:
            synth.write_imm_dense(imm, h, w, dense)
        root = os.path.join(d, "case.h5dir")
        write_config(root, dq, sq, frames, imm, **cfg)
        info = run(root, imm, g2out=g2out, darkout=darkout, threads=threads, extra=tuple(extra) + tuple(extra_args), cwd=d)
        res = listing(root, cfg.get("output", "/exchange"))
        return res, info
    finally:
        if not keep:
            shutil.rmtree(d, ignore_errors=True)


class SparseJob:
    """bench.py helper: the files of one multi-tau job are written once; run() executes the
    reference binary on them and returns its own stage scopes (seconds).  Sparse by default;
    dense=(int16 frames [darks + frames][P]) with darks / lld / sigma / flatfield for the dense path."""

    def __init__(self, dq, sq, frames, off=None, idx=None, val=None, dpl=8, swindow=None, dense=None, **cfg):
        from __graft_entry__ import load_package
        synth = load_package().synth
        self.dir = scratch_dir()
        self.imm = os.path.join(self.dir, "data.imm")
        h, w = np.asarray(dq).shape
        if dense is not None:
            synth.write_imm_dense(self.imm, h, w, dense)
        else:
            synth.write_imm_sparse(self.imm, h, w, off, idx, val)
        self.root = os.path.join(self.dir, "case.h5dir")
        write_config(self.root, dq, sq, frames, self.imm, dpl=dpl, static_window=swindow, **cfg)
        import atexit
        atexit.register(self.close)

    def run(self, threads=None):
        info = run(self.root, self.imm, g2out=False, threads=threads, cwd=self.dir)
        s = info["scopes"]
        return {"total_s": s.get("Total"), "load_s": s.get("Loading data"),
                "multitau_s": s.get("Computing G2 MultiTau"), "normalize_s": s.get("Normalizing Data"),
                "wall_s": info["seconds"]}

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)
