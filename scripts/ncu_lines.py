"""Aggregates an ncu --page source --print-source cuda,sass CSV per (kernel, file, line):
instructions executed, stall samples, shared wavefronts.  Usage: ncu_lines.py rep.ncu-rep [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
by = 1 if (len(sys.argv) > 3 and sys.argv[3] == "samples") else 0     # sort key: instructions (default) or stall samples
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, ""])
tot = collections.defaultdict(lambda: [0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1].split("(")[0].split("::")[-1].split("<")[0]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] == "": continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def g(name):
        try: return int(r[hdr[name]])
        except Exception: return 0
    a = agg[(cur_fn, cur_file, line)]
    a[0] += g("Instructions Executed"); a[1] += g("# Samples"); a[2] += g("L1 Wavefronts Shared"); a[3] += g("stall_long_sb"); a[4] = r[1][:90]
    tot[cur_fn][0] += g("Instructions Executed"); tot[cur_fn][1] += g("# Samples")
for fn in tot:
    print("==== %s: %d warp-instr, %d samples" % (fn, tot[fn][0], tot[fn][1]))
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    items.sort(key=lambda kv: -kv[1][by])
    for (f, fl, ln), v in items[:top]:
        print("%5.1f%% inst %5.1f%% smp  smemwf %9d  %s:%d  %s" % (100.0 * v[0] / tot[fn][0], 100.0 * v[1] / max(tot[fn][1], 1), v[2], fl, ln, v[4]))
