#!/usr/bin/env python
"""bench.py -- multi-tau G2 throughput of the B200-native XPCS hot path (BASELINE.json metric).

One "step" = one whole correlation job over one batch of synthetic input:
    frame-major sparse IMM events -> Filter/ingest (pixel-major store) -> multi-tau G2/IP/IF
    -> q-bin normalisation (norm-0-g2, norm-0-stderr)
Workload (default `c3`): BASELINE.json configs[2], the configuration north_star's target is
quoted on -- sparse 1 Mpixel detector (1024x1024), 100 000 frames, 0.1 % occupancy, dpl 8,
36 dynamic / 360 static annular q-bins.  With N GPUs it is the SAME job ("strong"): ONE 1-Mpixel
detector, every GPU holds 1/N of the frames (its share of the file) and owns 1/N of the pixel rows
(cut at static-bin boundaries); the library moves the events to their owners over NVLink and
all-reduces the sums and normalisation partials (NCCL inside the C-ABI, csrc/comm.cu).  `value` =
F / t.  The data is the same for every N (8 slabs seeded by slab number) and the run asserts that
g2 / stderr equal a one-GPU run bit for bit, plus a sampled-row check against the oracle
("parity").  `--scaling weak` keeps round 1's N-modules construction.
`--workload c2` runs BASELINE.json configs[1] (dense 1024x1024 int16, dark/flat/threshold).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c4|c5]

Contract keys: metric/value/unit/n_gpus/steps/warmup/ms_per_step/higher_is_better/scaling/
vs_baseline/dtype/data/config + clocks, e2e, gpu_launches, roofline, cpu_baseline.
The oracle (oracle/) is executed here only for the cpu_baseline leg and for --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "multi-tau G2 frames/sec"
UNIT = "frames/s"

WORKLOADS = {
    # name: (module h, module w, frames, occupancy, description)
    "c3": dict(h=1024, w=1024, F=100000, occ=0.001, kind="sparse",
               name="sparse IMM 1 Mpixel (1024x1024), 100k frames, 0.1% occupancy, dpl 8, 36 dynamic/360 static q-bins (BASELINE configs[2])"),
    "c1": dict(h=512, w=512, F=10000, occ=0.01, kind="sparse",
               name="sparse IMM 512x512, 10k frames, 1% occupancy, dpl 8, 36 dynamic q-bins (BASELINE configs[0])"),
    "c4": dict(h=256, w=256, F=10000, occ=0.02, kind="twotime",
               name="two-time correlation, 64k-pixel q ROI (256x256), 10k frames, 2% occupancy, symmetric smoothing (BASELINE configs[3])"),
    "c5": dict(h=2048, w=2048, F=1000000, occ=0.0001, kind="sparse",
               name="high-rate sparse sweep 2048x2048, 1M frames, 0.01% occupancy (low end of BASELINE configs[4]; --occupancy overrides)"),
    "c2": dict(h=1024, w=1024, F=20000, occ=None, kind="dense",
               name="non-sparse IMM 1024x1024 int16, 20k frames, dark/flat correction + threshold (BASELINE configs[1])"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index):
    """One process per GPU: run on the CPUs NVML reports as local to the GPU, so that the pinned staging
    buffers (first touch) and the driver threads sit on the GPU's own NUMA node -- with 8 ranks reading
    630 MB each per step from host memory, cross-socket traffic is the first thing to go.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# synthetic input
# ----------------------------------------------------------------------------------------
def module_maps(pkg, wl, n_modules):
    dq, sq = pkg.synth.annular_qmaps(wl["h"], wl["w"], n_dynamic=36, static_per_dynamic=10, r_min=8.0)
    if n_modules > 1:
        dq = np.tile(dq, (n_modules, 1))
        sq = np.tile(sq, (n_modules, 1))
    return np.ascontiguousarray(dq), np.ascontiguousarray(sq)


def gen_sparse_device(torch, pixels_dev, n_pix, F, occ, seed, device):
    """Every (frame, pixel) cell of this rank's pixel universe fires with probability occ
    (geometric gaps), count = 1 + Poisson(0.1).  pixels_dev = sorted int32 pixel ids (or None
    for 0..n_pix-1).  Returns device tensors idx int32[E], val int16[E], off int64[F+1]."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cells = int(n_pix) * int(F)
    expect = cells * occ
    n = int(expect + 6.0 * np.sqrt(expect + 1.0) + 1024)
    gaps = torch.empty(n, dtype=torch.float64, device=device).geometric_(occ, generator=g)
    pos = torch.cumsum(gaps.to(torch.int64), 0) - 1
    del gaps
    pos = pos[pos < cells]
    fr = torch.div(pos, n_pix, rounding_mode="floor")
    j = (pos - fr * n_pix)
    off = torch.searchsorted(pos, torch.arange(F + 1, device=device, dtype=torch.int64) * n_pix)
    del pos, fr
    idx = j.to(torch.int32) if pixels_dev is None else pixels_dev[j]
    del j
    lam = torch.full((idx.numel(),), 0.1, dtype=torch.float32, device=device)
    val = (1 + torch.poisson(lam, generator=g)).clamp_(1, 4000).to(torch.int16)
    return idx.contiguous(), val.contiguous(), off.contiguous()


def gen_sparse_host(pkg, P, F, occ, seed):
    return pkg.synth.sparse_frames(P, F, occ, seed=seed)


# ----------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------
def cpu_sample_job(pkg, O, wl, F_s, seed=1234):
    """One bounded CPU sample of the same workload: same detector, occupancy, q-maps and dpl,
    F_s frames.  Uses oracle/_ref (the compiled reference) when present, else the C port."""
    h, w, occ = wl["h"], wl["w"], wl["occ"]
    P = h * w
    dq, sq = module_maps(pkg, wl, 1)
    off, idx, val = gen_sparse_host(pkg, P, F_s, occ, seed)
    ref = None
    try:
        from oracle import refdrv
        if refdrv.available():
            ref = refdrv
    except Exception:
        ref = None
    # all the host cores this process may use; not OMP_NUM_THREADS, which torchrun pins to 1
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    if ref is not None:
        job = ref.SparseJob(dq, sq, F_s, off, idx, val, dpl=8, swindow=max(1, F_s // 10))

        def run():
            st = job.run(threads=threads)
            return float(st["total_s"]), st  # the reference's own "Total" scope (main.cpp:108)
        kind = "reference"
    else:
        qm = O.QMap(dq, sq)

        def run():
            t0 = time.perf_counter()
            fo = O.sparse_filter(qm, F_s, off, idx, val, swindow=max(1, F_s // 10))
            t1 = time.perf_counter()
            G2, IP, IF = O.multitau(P, F_s, 8, fo.rows, compat=True, nthreads=threads)
            t2 = time.perf_counter()
            O.normalize(qm, G2, IP, IF)
            t3 = time.perf_counter()
            return t3 - t0, {"filter_s": t1 - t0, "multitau_s": t2 - t1, "normalize_s": t3 - t2}
        kind = "port"
    return run, kind, threads, int(idx.size)


def full_config(wl_key):
    """The one full-size run of the reference binary per round (profiles/ref_full_config.py), if recorded."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ref_full_config.json"))).get(wl_key)
    except Exception:
        return None


def run_reference_arm(args, wl, wl_key):
    """--impl reference: the reference's CPU implementation of the path on this box's host
    cores (oracle/_ref when it was compiled, else the oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = entry.load_package()
    O = entry.load_oracle()
    if wl["kind"] != "sparse":
        wl = WORKLOADS["c3"]
        wl_key = "c3"
    # c1 (BASELINE configs[0], the reference's own CPU-runnable case) runs at its full size: a same-config pair;
    # the larger workloads run a bounded sample (5 000 frames of the same detector, occupancy and q-maps)
    F_s = args.cpu_frames or (wl["F"] if wl_key == "c1" else min(5000, wl["F"]))
    run, kind, threads, E = cpu_sample_job(pkg, O, wl, F_s)
    for _ in range(args.warmup):
        run()
    ts, stages = [], None
    for _ in range(args.steps):
        t, stages = run()
        ts.append(t)
    t = float(np.mean(ts))
    value = F_s / t
    sample = "%d of %d frames of the %s detector at %.3g occupancy (%d events), all stages; frames/s of the sample" % (
        F_s, wl["F"], "%dx%d" % (wl["h"], wl["w"]), wl["occ"], E)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "workload_key": wl_key, "sample": sample, "same_config": F_s == wl["F"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "same_config": F_s == wl["F"], "stages_s": stages, "full_config": full_config(wl_key)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
KERNEL_GROUP = {  # kernel name -> stage of SURVEY.md 8(d)
    "k_demux_count": "K0", "k_demux_scan": "K0", "k_demux_scatter": "K0", "k_merge_offsets": "K0", "k_concat": "K1",
    "k_chunk_rows": "K1",
    "k_block_frames": "K1", "k_hist": "K1", "k_slice_len": "K1", "k_slice_scan": "K1", "k_scatter": "K1",
    "k_finalize": "K1", "k_frame_scale": "K1", "k_hist_dense": "K1", "k_scatter_dense": "K1",
    "k_finalize_warp": "K1", "k_scatter_rec": "K1", "k_scatter_rec_dense": "K1", "k_place": "K1", "k_dense_bounds": "K2",
    "k_dense_filter": "K2", "k_dark": "K3", "k_multitau": "K4", "k_multitau_warp": "K4", "k_multitau_slice": "K4", "k_multitau_slicef": "K4", "k_multitau_warpf": "K4",
    "k_unpermute": "K4",
    "k_segment_reduce": "K6", "k_normalize_finish": "K6",
}


def algorithmic_bytes(name, E, T, R, Q, P, F_dense=0):
    """Algorithmic bytes per launch of each kernel (DESIGN.md 'Kernels' table; SURVEY.md 8d:
    one event = 6 B, one correlator value = 4 B)."""
    tbl = {
        "k_demux_count": 4 * E,                # multi-GPU slab partition: the pixel indices of the slab
        "k_demux_scatter": 12 * E,             # read the slab, write the per-owner streams
        "k_hist": 6 * E,                       # one read of the frame-major events
        "k_scatter": 12 * E,                   # read frame-major, write pixel-major
        "k_scatter_rec": 12 * E,               # read frame-major, append to the slice streams (8 B records move)
        "k_place": 12 * E,                     # read the slice streams, write the rows (same 6 B/event accounting)
        "k_finalize": 12 * E,                  # read + write the pixel-major rows once
        "k_multitau": 6 * E + 12 * T * R,      # read each event once, write G2/IP/IF once
        "k_multitau_warp": 6 * E + 12 * T * R,
        "k_multitau_slice": 6 * E + 12 * T * R,
        "k_multitau_warpf": 6 * E + 12 * T * R,  # same algorithmic bytes; the float store holds 8 B/event
        "k_multitau_slicef": 6 * E + 12 * T * R,
        "k_finalize_warp": 12 * E,
        "k_segment_reduce": 12 * T * R,        # read G2/IP/IF once
        "k_normalize_finish": 8 * T * Q,
        "k_dense_filter": 2 * P * F_dense + 16 * P,
        "k_unpermute": 8 * T * R,
    }
    return tbl.get(name, 0)


def gen_dense_device(torch, P, F, darks, mu, dev, seed=99, chunk=200):
    """int16 [darks + F][P] on the device: offset 100 ADU + N(0, 2) read noise; data frames add
    20 ADU per photon, photons ~ Poisson(mu)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((darks + F, P), dtype=torch.int16, device=dev)
    for f0 in range(0, darks + F, chunk):
        n = min(chunk, darks + F - f0)
        fr = 100.0 + 2.0 * torch.randn((n, P), device=dev, generator=g)
        lam = torch.full((n, P), mu, device=dev)
        ph = torch.poisson(lam, generator=g)
        if f0 < darks:
            ph[: max(0, min(n, darks - f0))] = 0.0
        out[f0: f0 + n] = torch.round(fr + 20.0 * ph).clamp_(-32768, 32767).to(torch.int16)
    return out


def bench_dense(args, wl):
    """BASELINE configs[1]: non-sparse IMM 1024x1024 int16, 20k frames, 100 dark frames, flat-field,
    threshold lld + sigma*dark_std, then multi-tau + normalisation -- one B200."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    pkg = entry.load_package()
    h, w, F = wl["h"], wl["w"], wl["F"]
    P = h * w
    darks, mu, lld, sigma = 100, 0.02, 5.0, 3.0
    dev = torch.device("cuda", 0)
    dq, sq = module_maps(pkg, wl, 1)
    # flat-field spread 1 %: the reference subtracts dark_avg = mean(raw*flat) from the un-flat-fielded raw
    # value, so a wide flat-field turns the 100 ADU offset into spurious survivors
    flat = pkg.synth.flatfield(P, sigma=0.01)
    frames = gen_dense_device(torch, P, F, darks, mu, dev)
    E_est = int(P * F * mu * 1.3) + (1 << 20)
    c = pkg.Correlator(dq, sq, F, dpl=8, flatfield=flat, lld=lld, sigma=sigma, device=0, reserve_events=E_est,
                       compat=not args.no_compat)
    c.set_dark(frames[:darks].cpu().numpy())
    data = frames[darks:]
    T, Q, S = c.T, c.Q, c.S
    stream = torch.cuda.Stream(device=dev)
    c.set_stream(stream.cuda_stream)

    def step_device():
        c.reset()
        c.push_dense_device(data.data_ptr(), F)
        c.finish_ingest(want=False)
        c.multitau(want=False)
        return c.normalize()

    for _ in range(max(args.warmup, 1)):
        g2, se = step_device()
    info = c.info()
    E, R = int(info.events_stored), int(info.n_rows)
    c.kernel_report(reset=True)
    c.kernel_timing(True)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        g2, se = step_device()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1) / args.steps
    launches = int(c.launch_count())
    report = c.kernel_report(reset=True)
    c.kernel_timing(False)
    # end to end: 42 GB of pinned host frames through the public API
    e2e = None
    if not args.no_e2e:
        host = torch.empty((F, P), dtype=torch.int16, pin_memory=True)
        host.copy_(data)
        torch.cuda.synchronize()
        nst = max(1, min(args.steps, 3))

        def step_e2e():
            c.reset()
            c.push_dense_raw(host.data_ptr(), F)
            sums = c.finish_ingest(want=True)
            c.multitau(want=False)
            return sums, c.normalize()

        step_e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(nst):
            step_e2e()
        torch.cuda.synchronize()
        ms_e2e = 1e3 * (time.perf_counter() - t0) / nst
        e2e = {"value": F / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * P * F,
               "d2h_bytes_per_step": 4 * (P + 2 * F + S + (F // c.static_window) * S) + 8 * T * Q, "ms_per_step": ms_e2e,
               "steps": nst, "api": "Correlator.push_dense/finish_ingest/multitau/normalize (C-ABI xpcs_*)"}
    clocks = sampler.stop()
    peak, peak_src = peaks()
    kern = {}
    for name, (ms, n) in report.items():
        if n <= 0:
            continue
        b = algorithmic_bytes(name, E, T, R, Q, P, F_dense=F)
        if name == "k_dense_filter":
            b = 2 * P * F + 16 * P          # whole job; launched in frame batches
            per = ms / args.steps
        else:
            per = ms / n
        kern[name] = {"ms_per_launch": ms / n, "launches_per_step": n / args.steps, "ms_per_step": ms / args.steps,
                      "algo_bytes": b, "gbs": (b / (per * 1e-3) / 1e9) if b else None, "stage": KERNEL_GROUP.get(name, "?")}
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"])
    k = kern[dom]
    line = {
        "metric": METRIC, "value": F / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (dark/flat/threshold in f64 where the reference does)", "data": "synthetic",
        "config": {"workload": wl["name"], "workload_key": "c2", "detector_pixels": P, "frames": F, "dark_frames": darks,
                   "photons_per_pixel_frame": mu, "lld": lld, "sigma": sigma, "events_after_threshold": E, "delays": T,
                   "l2": "inputs (%.1f GB/step) exceed the 126 MB L2" % (2.0 * P * F / 1e9)},
        "pixel_frames_per_s": float(P) * F / (ms_dev * 1e-3), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": (k["gbs"] / peak) if k["gbs"] else None, "traffic": None, "peak_source": peak_src,
                     "algo_bytes_per_step": k["algo_bytes"], "kernel_share_of_step": k["ms_per_step"] / ms_dev},
        "kernels": kern, "results_finite": bool(np.isfinite(g2).all()),
    }
    tr = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the ncu --set full capture
    if os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            line["roofline"]["traffic"] = tj.get("c2", {}).get(dom)
            line["roofline"]["traffic_source"] = tj.get("_source_c2", "profiles/traffic.json")
        except Exception:
            pass
    if not args.no_parity:
        # parity of the timed configuration (untimed; the oracle is the checker): the full time series of sampled
        # pixels through the oracle's dark image, dense filter and multiTau2 against the device's G2 / IP / IF
        # (1e-5 relative, north_star; the survivors of the filter must be the same samples)
        try:
            from oracle import oracle as O
            line["parity"] = parity_block_dense(torch, O, c, dq, sq, flat, frames, darks, F, lld, sigma,
                                                n_rows=min(args.parity_rows, 512), compat=not args.no_compat)
        except Exception as ex:
            line["parity"] = {"ok": False, "error": repr(ex)}
        if not line["parity"].get("ok"):
            print(json.dumps(line))
            raise SystemExit("bench.py: parity block failed: %r" % (line["parity"],))
    if not args.no_cpu:
        try:
            from oracle import refdrv
            if refdrv.available():
                F_s, d_s = args.cpu_frames or 300, 20
                fr = gen_dense_device(torch, P, F_s, d_s, mu, dev, seed=7).cpu().numpy()
                job = refdrv.SparseJob(dq, sq, F_s, dense=fr, dpl=8, swindow=max(1, F_s // 10), darks=d_s, lld=lld,
                                       sigma=sigma, flatfield=flat)
                ncores = len(os.sched_getaffinity(0))
                st = job.run(threads=ncores)
                line["cpu_baseline"] = {"value": F_s / st["total_s"], "unit": UNIT, "cores": ncores,
                                        "kind": "reference", "seconds": st["total_s"], "stages_s": st,
                                        "sample": "%d of %d frames (+%d darks), same detector, flat-field and threshold" % (F_s, F, d_s)}
        except Exception as ex:  # the bench line must still be printed
            line["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(line))
    c.close()
    return 0


def bench_twotime(args, wl):
    """BASELINE configs[3]: two-time correlation of one 64k-pixel dynamic partition over 10k
    frames -- the tensor-core stage.  One step = sg + operand build + C = triu(X X^T) scaled +
    diagonal statistics, device-resident (events already ingested); C stays on the device."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    pkg = entry.load_package()
    h, w, F, occ = wl["h"], wl["w"], wl["F"], wl["occ"]
    P = h * w
    dq = np.ones((h, w), np.int32)
    sq = (1 + (np.arange(P) // 4096)).astype(np.int32).reshape(h, w)   # 16 static bins inside the ROI
    dev = torch.device("cuda", 0)
    d_idx, d_val, d_off = gen_sparse_device(torch, None, P, F, occ, 4321, dev)
    E = int(d_idx.numel())
    c = pkg.Correlator(dq, sq, F, dpl=8, device=0, reserve_events=E)
    c.push_sparse_device(d_idx.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), E, F)
    c.finish_ingest(want=False)
    wsize = 100
    for _ in range(max(args.warmup, 1)):
        c.twotime(1, wsize, want_c=False)
    c.kernel_report(reset=True)
    c.kernel_timing(True)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = c.twotime(1, wsize, want_c=False)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    rep = c.kernel_report(reset=True)
    gemm_ms = rep["k_twotime_gemm"][0] / max(rep["k_twotime_gemm"][1], 1)
    flops = float(P) * F * (F + 1)            # upper triangle incl. diagonal, 2 flop per MAC
    tm, tn = (F + 127) // 128, (F + 255) // 256     # 128 x 256 tiles that touch the upper triangle (csrc/twotime.cu)
    n_tiles = sum(1 for mt in range(tm) for nt in range(tn) if nt * 256 + 255 >= mt * 128)
    flops_issued = 2.0 * n_tiles * 128 * 256 * (-(-P // 64) * 64)
    peak_tf, peak_sus = 1402.4, None
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_tf = float(pk["bf16_tflops"])
        peak_sus = float(pk.get("bf16_tflops_sustained", 0)) or None
    except Exception:
        pass
    line = {
        "metric": "two-time correlation frames/sec", "value": F / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16 operands (exact counts) / fp32 TMEM accumulate", "data": "synthetic",
        "config": {"workload": wl["name"], "workload_key": "c4", "roi_pixels": P, "frames": F, "events": E,
                   "l2": "operand matrix %.0f MB + C %.0f MB exceed L2" % (2.0 * P * F / 1e6, 4.0 * F * F / 1e6)},
        "clocks": clocks, "gpu_launches": int(c.launch_count()),
        "roofline": {"kernel": "k_twotime_gemm", "bound": "tensor", "achieved": flops / (gemm_ms * 1e-3) / 1e12,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": flops / (gemm_ms * 1e-3) / 1e12 / peak_tf,
                     "traffic": None, "flops_algorithmic": flops, "flops_issued": flops_issued,
                     "achieved_issued": flops_issued / (gemm_ms * 1e-3) / 1e12, "ms_per_launch": gemm_ms,
                     "peak_sustained": peak_sus, "frac_of_sustained": (flops / (gemm_ms * 1e-3) / 1e12 / peak_sus) if peak_sus else None,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; fp16 runs at the same rate)"},
        "kernels": {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1]} for k, v in rep.items() if v[1]},
        "g2full_head": [float(x) for x in r["g2full"][:4]],
    }
    print(json.dumps(line))
    c.close()
    return 0


# ----------------------------------------------------------------------------------------
# parity block of the sparse bench legs (the oracle as the checker, outside every timed region)
# ----------------------------------------------------------------------------------------
def rows_of_pixels(O, pix_sorted, ev_pix, ev_frame, ev_val):
    """CSR rows (oracle.Rows) of the listed pixels from their events: row i = pixel pix_sorted[i], frames ascending."""
    row = np.searchsorted(pix_sorted, ev_pix)
    order = np.lexsort((ev_frame, row))
    row, t, v = row[order], ev_frame[order].astype(np.int32), ev_val[order].astype(np.float32)
    row_ptr = np.zeros(pix_sorted.size + 1, np.int64)
    np.add.at(row_ptr, row + 1, 1)
    return O.Rows(np.cumsum(row_ptr), t, v)


def parity_block(torch, pkg, O, c, dq, sq, F, ev_source, own_pixels, g2_dev, n_rows=2000, compat=True, seed=5):
    """Bit-for-bit check of the timed configuration itself (not of a small stand-in):
    (1) G2 / IP / IF of ~n_rows sampled pixel rows against oracle.multitau (corr.cpp:315-431) on those rows' events;
    (2) norm-0-g2 of a few dynamic bins recomputed by oracle.normalize (corr.cpp:927-1091) from the device's
        G2 / IP / IF of the bins' pixels.
    ev_source() yields (idx, val, off, first_frame) device tensors that together cover ALL frames of the job (whole
    detector), one slab of frames at a time (a C5 point does not fit a GPU twice)."""
    P = dq.size
    dev = torch.device("cuda", torch.cuda.current_device())
    rng = np.random.default_rng(seed)
    valid = np.flatnonzero((dq.ravel() > 0) & (sq.ravel() > 0))
    own = np.asarray(own_pixels, np.int64)
    samp = np.sort(rng.choice(own, size=min(n_rows, own.size), replace=False)).astype(np.int32)

    def events_of(pixels):
        lut = torch.zeros(P, dtype=torch.bool, device=dev)
        lut[torch.from_numpy(np.asarray(pixels, np.int64)).to(dev)] = True
        ep, ef, ev = [], [], []
        step = 1 << 28   # bounded temporaries
        for d_idx, d_val, d_off, f0 in ev_source():
            for a in range(0, int(d_idx.numel()), step):
                sl = slice(a, min(a + step, int(d_idx.numel())))
                e = torch.nonzero(lut[d_idx[sl].long()]).squeeze(1) + a
                ep.append(d_idx[e].cpu().numpy())
                ef.append((torch.searchsorted(d_off, e, right=True) - 1 + f0).cpu().numpy())
                ev.append(d_val[e].cpu().numpy())
            del d_idx, d_val, d_off
        return np.concatenate(ep), np.concatenate(ef), np.concatenate(ev)

    T = c.T
    out = {"rows": int(samp.size), "row_entries": int(3 * T * samp.size)}
    ep, ef, ev = events_of(samp)
    rows = rows_of_pixels(O, samp, ep, ef, ev)
    rG2, rIP, rIF = O.multitau(samp.size, F, 8, rows, compat=compat)
    G2, IP, IF = c.correlators(samp)
    mism = int((G2 != rG2).sum() + (IP != rIP).sum() + (IF != rIF).sum())
    out["mismatches"] = mism
    out["row_events"] = int(ep.size)
    out["nonzero_g2_entries"] = int((rG2 != 0).sum())

    # (2) dynamic bins wholly owned by this rank: the first two and the last such
    dqf, sqf = dq.ravel(), sq.ravel()
    ownset = np.zeros(P, bool)
    ownset[own] = True
    bins = [q for q in range(1, int(dqf.max()) + 1) if ownset[valid[dqf[valid] == q]].all() and (dqf[valid] == q).any()]
    pick = sorted(set(bins[:2] + bins[-1:]))
    nm = 0
    for q in pick:
        pix = np.flatnonzero((dqf == q) & (sqf > 0)).astype(np.int32)   # ascending: the order inside a static bin
        G2q, IPq, IFq = c.correlators(pix)
        sq_local = np.unique(sqf[pix], return_inverse=True)[1].astype(np.int32) + 1
        qm = O.QMap(np.ones((1, pix.size), np.int32), sq_local.reshape(1, -1))
        rg2, _ = O.normalize(qm, G2q, IPq, IFq)
        nm += int(G.n_diff_arrays(rg2[:, 0], g2_dev[:, q - 1]))
    out["norm_bins"] = pick
    out["norm_mismatches"] = nm
    out["ok"] = mism == 0 and nm == 0
    return out


def parity_block_dense(torch, O, c, dq, sq, flat, frames, darks, F, lld, sigma, n_rows=512, compat=True, seed=5, rtol=1e-5):
    """Dense leg: ~n_rows sampled pixels as a small detector of their own -- oracle.dark_image
    (dark_image.cpp:81-106), oracle.dense_filter (dense_filter.cpp:121-210) and oracle.multitau (corr.cpp:315-431)
    on their full time series -- against the device's G2 / IP / IF columns of those pixels.  Floating point:
    within rtol; the pattern of exact zeros must be identical."""
    rng = np.random.default_rng(seed)
    dqf, sqf = dq.ravel(), sq.ravel()
    valid = np.flatnonzero((dqf > 0) & (sqf > 0))
    pix = np.sort(rng.choice(valid, size=min(n_rows, valid.size), replace=False)).astype(np.int32)
    sub = frames[:, torch.from_numpy(pix.astype(np.int64)).to(frames.device)].cpu().numpy()   # (darks + F, n)
    qm = O.QMap(np.ones((1, pix.size), np.int32), np.ones((1, pix.size), np.int32))   # (the bins do not enter G2 / IP / IF)
    fl = np.ascontiguousarray(np.asarray(flat, np.float64).ravel()[pix])
    dark = O.dark_image(np.ascontiguousarray(sub[:darks]), fl)
    fo = O.dense_filter(qm, F, np.ascontiguousarray(sub[darks:]), flat=fl, dark=dark, lld=lld, sigma=sigma,
                        swindow=max(1, F // 10))
    rG = O.multitau(qm.P, F, 8, fo.rows, compat=compat)
    dG = c.correlators(pix)
    worst, bad, zeros = 0.0, 0, 0
    for a, b in zip(dG, rG):
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        err = np.abs(a64 - b64)
        bad += int((~((np.isnan(a64) & np.isnan(b64)) | (err <= rtol * np.abs(b64)))).sum())
        zeros += int(((a == 0.0) != (b == 0.0)).sum())
        nz = b64 != 0
        if nz.any():
            worst = max(worst, float(np.nanmax(err[nz] / np.abs(b64[nz]))))
    return {"rows": int(pix.size), "row_entries": int(3 * dG[0].size), "row_events": int(fo.n), "rtol": rtol,
            "mismatches": bad, "zero_pattern_mismatches": zeros, "worst_rel_err": worst,
            "nonzero_g2_entries": int((rG[0] != 0).sum()), "ok": bad == 0 and zeros == 0}


class G:  # tiny helper namespace (NaN-aware inequality count)
    @staticmethod
    def n_diff_arrays(a, b):
        a = np.asarray(a).ravel()
        b = np.asarray(b).ravel()
        return int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b)))))


N_SLABS = 64  # the synthetic job is generated as 64 slabs of frames seeded by slab number: the same data for every GPU
              # count, and generator temporaries that stay bounded at 4e10 events


def bench_sparse(args, wl):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    weak = args.scaling == "weak" and world > 1
    if world > 1:
        bind_to_gpu_numa_node(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the one JSON line: NCCL's logs go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    pkg = entry.load_package()
    F, occ = wl["F"], wl["occ"]
    dq, sq = module_maps(pkg, wl, world if weak else 1)
    P = dq.size
    E_est = int(wl["h"] * wl["w"] * F * occ * 1.02 / (1 if weak else world)) + (1 << 16)
    c = pkg.Correlator(dq, sq, F, dpl=8, compat=not args.no_compat, device=local, shard_index=rank,
                       shard_count=world, reserve_events=E_est)
    if world > 1:  # the library's own communicator (NCCL bound at run time: the copy torch loaded)
        uid = [pkg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(world, rank, uid[0])
    c.pinned_results(True)   # result arrays in page-locked buffers of the library (xpcs_host_alloc), reused every step
    info = c.info()
    T, Q, R, S = c.T, c.Q, info.n_rows, c.S
    stream = torch.cuda.Stream(device=dev)
    c.set_stream(stream.cuda_stream)

    # ---- synthetic input, resident in HBM before any timed region ----
    n_slabs = N_SLABS if N_SLABS % world == 0 else world
    slab_frames = [F * (k + 1) // n_slabs - F * k // n_slabs for k in range(n_slabs)]
    slab_first = [F * k // n_slabs for k in range(n_slabs)]

    def gen_slabs(ks):
        parts, base = [], 0
        offs = [torch.zeros(1, dtype=torch.int64, device=dev)]
        for k in ks:
            i, v, o = gen_sparse_device(torch, None, wl["h"] * wl["w"], slab_frames[k], occ, 1234 + k, dev)
            parts.append((i, v))
            offs.append(o[1:] + base)
            base += int(i.numel())
        return (torch.cat([p[0] for p in parts]).contiguous(), torch.cat([p[1] for p in parts]).contiguous(),
                torch.cat(offs).contiguous())

    if weak:
        # N modules of the detector sharing the q-bins, every rank fed its own (demultiplexed) pixels: per-GPU work fixed
        own = c.row_pixels()
        masked = np.nonzero((dq.ravel() < 1) | (sq.ravel() < 1))[0].astype(np.int32)[rank::world]
        uni = np.sort(np.concatenate([own, masked])).astype(np.int32)
        d_idx, d_val, d_off = gen_sparse_device(torch, torch.from_numpy(uni).to(dev), int(uni.size), F, occ, 1234 + rank, dev)
        first, nfr = 0, F
    else:
        my = list(range(rank * n_slabs // world, (rank + 1) * n_slabs // world))
        d_idx, d_val, d_off = gen_slabs(my)
        first, nfr = slab_first[my[0]], sum(slab_frames[k] for k in my)
    E = int(d_idx.numel())
    torch.cuda.empty_cache()   # the generator's temporaries go back to the driver: the library allocates with cudaMalloc
    if not args.no_e2e:
        h_idx = torch.empty(E, dtype=torch.int32, pin_memory=True).copy_(d_idx)
        h_val = torch.empty(E, dtype=torch.int16, pin_memory=True).copy_(d_val)
        h_off = torch.empty(nfr + 1, dtype=torch.int64, pin_memory=True).copy_(d_off)
    torch.cuda.synchronize()

    def step_device():
        c.reset()
        if weak:
            c.push_sparse_device(d_idx.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), E, F)
        else:
            c.push_sparse_slab_device(first, d_idx.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), E, nfr)
        c.finish_ingest(want=False)
        c.multitau(want=False)
        return c.normalize()

    def step_e2e():
        c.reset()
        if weak:
            c.push_sparse_raw(h_idx.data_ptr(), h_val.data_ptr(), h_off.data_ptr(), F)
        else:
            c.push_sparse_slab_raw(first, h_idx.data_ptr(), h_val.data_ptr(), h_off.data_ptr(), nfr)
        sums = c.finish_ingest(want=True)
        c.multitau(want=False)
        g2, se = c.normalize()
        return sums, g2, se

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_over_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    # ---- device-resident timing (value) ----
    for _ in range(args.warmup):
        g2, se = step_device()
    c.kernel_report(reset=True)
    c.kernel_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        g2, se = step_device()
    e1.record(stream)
    barrier()
    wall_dev = 1e3 * (time.perf_counter() - t0) / args.steps
    ms_dev = reduce_over_ranks(e0.elapsed_time(e1) / args.steps, dist.ReduceOp.MAX if world > 1 else None)
    launches = int(reduce_over_ranks(c.launch_count(), dist.ReduceOp.SUM if world > 1 else None))
    report = c.kernel_report(reset=True)
    c.kernel_timing(False)
    finite_g2 = bool(np.isfinite(g2).all())
    g2_timed = np.array(g2, copy=True)
    Es = int(c.info().events_stored)

    # ---- end-to-end timing through the public API with host buffers ----
    ms_e2e, e2e_identical = None, None
    if not args.no_e2e:
        for _ in range(min(args.warmup, 3)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            sums, g2e, see = step_e2e()
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) / args.steps
        ms_e2e = reduce_over_ranks(max(e0.elapsed_time(e1) / args.steps, 1e3 * wall), dist.ReduceOp.MAX if world > 1 else None)
        e2e_identical = bool(np.array_equal(g2e, g2_timed, equal_nan=True))
    clocks = sampler.stop() if rank == 0 else None
    h2d = 4 * E + 2 * E + 8 * (nfr + 1)
    d2h = 4 * (P + 2 * F + S + (F // c.static_window) * S) + 2 * 4 * T * Q
    E_total = reduce_over_ranks(E, dist.ReduceOp.SUM if world > 1 else None)
    R_total = reduce_over_ranks(R, dist.ReduceOp.SUM if world > 1 else None)
    h2d_total = reduce_over_ranks(h2d, dist.ReduceOp.SUM if world > 1 else None)

    jobs = world if weak else 1
    value = jobs * F / (ms_dev * 1e-3)
    e2e_value = jobs * F / (ms_e2e * 1e-3) if ms_e2e else None

    # ---- roofline of the dominant kernel (rank 0's launches; library (NCCL) kernels listed but not ranked) ----
    peak, peak_src = peaks()
    kern = {}
    for name, (ms, n) in report.items():
        if n <= 0:
            continue
        per = ms / n
        Ek = Es if name in ("k_finalize", "k_finalize_warp", "k_multitau", "k_multitau_warp", "k_multitau_slice", "k_multitau_slicef", "k_multitau_warpf") else \
            (E if name.startswith("k_demux") else Es)
        b = algorithmic_bytes(name, Ek, T, R, Q, P)
        kern[name] = {"ms_per_launch": per, "launches_per_step": n / args.steps, "ms_per_step": ms / args.steps,
                      "algo_bytes": b, "gbs": (b / (per * 1e-3) / 1e9) if per > 0 and b else None,
                      "stage": KERNEL_GROUP.get(name, "NCCL" if name.startswith("nccl_") else "?")}
    own_k = [k for k in kern if not k.startswith("nccl_")]
    dom = max(own_k, key=lambda k: kern[k]["ms_per_step"]) if own_k else None
    roof = None
    if dom:
        k = kern[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s",
                "frac": (k["gbs"] / peak) if k["gbs"] else None, "traffic": None, "peak_source": peak_src,
                "algo_bytes_per_launch": k["algo_bytes"], "ms_per_launch": k["ms_per_launch"],
                "kernel_share_of_step": k["ms_per_step"] / ms_dev}
        tr = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the ncu --set full capture
        if os.path.exists(tr):
            try:
                tj = json.load(open(tr))
                roof["traffic"] = tj.get(args.workload, {}).get(dom)
                roof["traffic_source"] = tj.get("_source", "profiles/traffic.json (ncu --set full capture of this command)")
            except Exception:
                pass
    pipeline_bytes = 18 * Es + 12 * T * R
    kernels_ms = sum(k["ms_per_step"] for k in kern.values())

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak" if weak else "strong",
        "vs_baseline": None, "dtype": "u32 numerators / f32 quotients", "data": "synthetic",
        "config": {"workload": wl["name"], "workload_key": args.workload, "detector_pixels": int(P),
                   "rows_this_gpu": int(R), "rows_total": int(R_total), "frames": F, "events_this_gpu_slab": E,
                   "events_total": int(E_total), "delays": T, "q_bins": Q, "static_bins": S,
                   "parallelism": ("%d modules, one per GPU (weak)" % world) if weak else
                   ("ONE detector: frame slabs in (1/%d of the frames per GPU), pixel rows out (static-bin aligned), "
                    "events redistributed over NVLink by the library (ncclSend/ncclRecv), sums and normalisation "
                    "partials all-reduced (NCCL inside the C-ABI)" % world) if world > 1 else "single GPU",
                   "value_definition": "frames of the whole detector per second = F/t, all stages (exchange, ingest, multi-tau, normalise)"
                   if not weak else "1-Mpixel-module frames/s = n_gpus*F/t",
                   "l2": "inputs (%.0f MB/step per GPU) exceed the 126 MB L2; no flush needed" % (6 * E / 1e6)
                   if 6 * E > 126e6 else "inputs %.0f MB/step per GPU; results %.0f MB/step written in between evict them" % (6 * E / 1e6, 12.0 * T * R / 1e6),
                   "event_transport": {1: "direct NVLink stores from the partition kernel into the owners' lists (CUDA IPC mappings)",
                                       0: "staged: per-owner streams + grouped ncclSend/ncclRecv", -1: None}[c.comm_transport()],
                   "compat_stale_tail": not args.no_compat},
        "pixel_frames_per_s": float(R_total) * F / (ms_dev * 1e-3),
        "clocks": clocks,
        "e2e": None if ms_e2e is None else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total), "h2d_bytes_per_step_per_gpu": h2d,
                "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e, "g2_identical_to_device_resident_run": e2e_identical,
                "api": "Correlator.push_sparse_slab/finish_ingest/multitau/normalize (C-ABI xpcs_*), pinned host buffers"},
        "gpu_launches": launches,
        "roofline": roof,
        "pipeline": {"algo_bytes_per_step": pipeline_bytes, "formula": "18*E + 12*T*P_valid (SURVEY 8d PRIMARY), this GPU's share",
                     "kernel_ms_per_step": kernels_ms, "gbs_over_step": pipeline_bytes / (ms_dev * 1e-3) / 1e9,
                     "frac_of_peak_over_step": pipeline_bytes / (ms_dev * 1e-3) / 1e9 / peak,
                     "host_wall_ms_per_step": wall_dev},
        "kernels": kern,
        "results_finite": finite_g2,
    }

    # ---- parity of the timed configuration (rank 0; untimed; the oracle is the checker) ----
    ok = True
    if not args.no_parity and not weak:
        c.kernel_timing(False)
        g2p, sep = step_device()          # leaves G2 / IP / IF of this run on the device
        par = {}
        if rank == 0:
            O = entry.load_oracle()

            def ev_source():
                if world == 1:
                    yield d_idx, d_val, d_off, 0
                else:
                    for k in range(n_slabs):
                        torch.cuda.empty_cache()
                        yield gen_slabs([k]) + (slab_first[k],)

            par = parity_block(torch, pkg, O, c, dq, sq, F, ev_source, c.row_pixels(), g2p, n_rows=args.parity_rows,
                               compat=not args.no_compat)
            if world > 1 and 6.0 * E_total > 40e9:
                par["single_gpu_g2_identical"] = None   # the whole event list does not fit one GPU next to this rank's share
            elif world > 1:
                torch.cuda.empty_cache()
                all_ev = gen_slabs(range(n_slabs))
                # the same job on ONE GPU (this one), same data: g2 / stderr must be bit-identical
                c1 = pkg.Correlator(dq, sq, F, dpl=8, compat=not args.no_compat, device=local, reserve_events=int(all_ev[0].numel()))
                c1.push_sparse_device(all_ev[0].data_ptr(), all_ev[1].data_ptr(), all_ev[2].data_ptr(), int(all_ev[0].numel()), F)
                c1.finish_ingest(want=False)
                c1.multitau(want=False)
                g2_1, se_1 = c1.normalize()
                c1.close()
                par["single_gpu_g2_identical"] = bool(np.array_equal(g2_1, g2p, equal_nan=True))
                par["single_gpu_stderr_identical"] = bool(np.array_equal(se_1, sep, equal_nan=True))
                par["ok"] = par["ok"] and par["single_gpu_g2_identical"] and par["single_gpu_stderr_identical"]
            ok = bool(par["ok"])
        line["parity"] = par
        barrier()

    # ---- CPU baseline (rank 0, N=1 only): the reference binary on a bounded sample of the same workload ----
    if rank == 0 and world == 1 and not args.no_cpu:
        O = entry.load_oracle()
        F_s = args.cpu_frames or min(5000, F)
        run, kind, threads, E_s = cpu_sample_job(pkg, O, wl, F_s)
        t, stages = run()
        line["cpu_baseline"] = {
            "value": F_s / t, "unit": UNIT, "cores": threads, "kind": kind, "same_config": F_s == F,
            "sample": "%d of %d frames, same detector/occupancy/q-maps (%d events), all stages, one pass" % (F_s, F, E_s),
            "seconds": t, "stages_s": stages}
        fc = os.path.join(ROOT, "profiles", "ref_full_config.json")  # one full-F run of the reference per round
        if os.path.exists(fc):
            try:
                line["cpu_baseline"]["full_config"] = json.load(open(fc)).get(args.workload)
            except Exception:
                pass
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    c.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = ONE detector sharded over the GPUs (default); weak = N modules, one per GPU")
    ap.add_argument("--frames", type=int, default=0, help="override the frame count (debug)")
    ap.add_argument("--hw", type=int, nargs=2, default=None, help="override the detector height and width (debug: one GPU's share of a sharded job)")
    ap.add_argument("--occupancy", type=float, default=0.0, help="override the occupancy of a sparse workload (c5 sweep)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (c2: 42 GB pinned)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (sampled rows against the oracle)")
    ap.add_argument("--parity-rows", type=int, default=2000)
    ap.add_argument("--no-compat", action="store_true", help="exact sums instead of the reference's stale-tail behaviour")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.frames:
        wl["F"] = args.frames
    if args.hw:
        wl["h"], wl["w"] = args.hw
        wl["name"] += " [detector %dx%d]" % tuple(args.hw)
    if args.occupancy and wl["kind"] == "sparse":
        wl["occ"] = args.occupancy
        wl["name"] += " [occupancy %g]" % args.occupancy
    if args.impl == "reference":
        return run_reference_arm(args, wl, args.workload)
    if wl["kind"] == "dense":
        return bench_dense(args, wl)
    if wl["kind"] == "twotime":
        return bench_twotime(args, wl)
    return bench_sparse(args, wl)


if __name__ == "__main__":
    sys.exit(main())
