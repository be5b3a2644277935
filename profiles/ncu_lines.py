"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
instructions executed and stall samples.  usage: python profiles/ncu_lines.py file.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = None; cur_line = None; cur_src = ""
inst = collections.Counter(); samp = collections.Counter(); src = {}
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur_line = (fname, int(r[0])); src[cur_line] = r[1]
    if r[2] == "" and r[0] != "": continue  # pure source row (no SASS)
    try:
        inst[cur_line] += int(r[ii] or 0); samp[cur_line] += int(r[si] or 0)
    except ValueError: pass
tot = sum(inst.values()); ts = sum(samp.values())
print("total warp-instructions %d, samples %d" % (tot, ts))
for k, v in inst.most_common(top):
    print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100.0*v/tot, 100.0*samp[k]/max(ts,1), k[0], k[1], src[k].strip()[:110]))
