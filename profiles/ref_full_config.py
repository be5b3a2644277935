"""One full-size run of the UNMODIFIED reference binary (oracle/_ref/corr_ref) per configuration on this box's host
cores: the same-config CPU figure quoted in bench.py's cpu_baseline.full_config (the per-step reference arm of the
bench runs a 5 000-frame sample to stay within minutes).  usage: python profiles/ref_full_config.py c1 c3 > out.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402
from oracle import refdrv  # noqa: E402

pkg = entry.load_package()
out = {}
for key in sys.argv[1:] or ["c1"]:
    wl = bench.WORKLOADS[key]
    h, w, F, occ = wl["h"], wl["w"], wl["F"], wl["occ"]
    dq, sq = bench.module_maps(pkg, wl, 1)
    t0 = time.time()
    off, idx, val = pkg.synth.sparse_frames(h * w, F, occ, seed=1234)
    job = refdrv.SparseJob(dq, sq, F, off, idx, val, dpl=8, swindow=max(1, F // 10))
    threads = len(os.sched_getaffinity(0))
    st = job.run(threads=threads)
    out[key] = {"workload": wl["name"], "frames": F, "events": int(idx.size), "cores": threads, "kind": "reference",
                "frames_per_s": F / st["total_s"], "stages_s": st, "same_config": True,
                "prepare_s": time.time() - t0 - st["wall_s"]}
    job.close()
    print(key, out[key], file=sys.stderr)
print(json.dumps(out, indent=1))
