"""Where the warps of k_multitau_slice spend their cycles.  Build the library with
`make -C xpcs-eigen_b200 EXTRA_NVFLAGS=-DXPCS_SL_TRACE` (every 1031st CTA then prints one line per trace point:
`SLT <slice> <warp> <point> <task> <cycles since the CTA started>`), run
`python bench.py --no-cpu --no-e2e --no-parity --steps 1 --warmup 1 | grep ^SLT > trace.txt`, then
`python profiles/trace_slice.py trace.txt`.  Points: 1 staged, 2 merge-level histogram done, 3 live counts done,
4 limits done, 5 tasks start, 6 task taken (task id), 7 no task left, 8 last warp done, 9 G2 written."""
import collections
import statistics as st
import sys

names = {1: "staged", 2: "mlhist", 3: "base", 4: "limits", 5: "go", 6: "take", 7: "idle", 8: "final", 9: "end"}
by = collections.defaultdict(list)
for line in open(sys.argv[1]):
    r = line.split()
    by[int(r[1])].append((int(r[2]), names[int(r[3])], int(r[4]), int(r[5])))
dur, phase = collections.defaultdict(list), collections.defaultdict(list)
for s, evs in sorted(by.items()):
    evs = evs[len(evs) // 2:]  # the timed launch (the first half is the warm-up launch)
    for w in sorted({e[0] for e in evs}):
        seq = [(what, i, c) for (ww, what, i, c) in evs if ww == w]
        for k in range(len(seq) - 1):
            if seq[k][0] == "take":
                dur[seq[k][1]].append(seq[k + 1][2] - seq[k][2])
        d = {what: c for what, i, c in seq if what != "take"}
        for a, b in (("staged", "mlhist"), ("mlhist", "base"), ("base", "limits"), ("limits", "go"), ("go", "idle"),
                     ("idle", "final"), ("final", "end")):
            if a in d and b in d:
                phase[a + " -> " + b].append(d[b] - d[a])
        if "staged" in d:
            phase["start -> staged"].append(d["staged"])
        if "end" in d:
            phase["whole CTA"].append(d["end"])
print("CTAs traced: %d" % len(by))
for k in phase:
    print("%-18s mean %7d  max %7d cycles" % (k, st.mean(phase[k]), max(phase[k])))
for k in sorted(dur):
    print("task %2d  n %3d  mean %7d  max %7d cycles" % (k, len(dur[k]), st.mean(dur[k]), max(dur[k])))
