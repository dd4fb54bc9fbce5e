"""Stall-reason mix per source region (function) of one kernel from an ncu report.  Usage: ncu_stalls_by_region.py rep kernel-regex"""
import collections, csv, os, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mgnet_b200", "csrc")
def fn_ranges(path):
    names = {}; cur = "?"
    for i, l in enumerate(open(path), 1):
        m = re.search(r"(?:__device__|__global__)[^;(]*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m: cur = m.group(1)
        names[i] = cur
    return names
maps = {f: fn_ranges(os.path.join(base, f)) for f in os.listdir(base) if f.endswith(".cuh")}
cur_file = cur_fn = None; hdr = None
agg = collections.defaultdict(collections.Counter); tot = collections.Counter()
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or cur_fn is None or not re.search(kern, cur_fn): continue
    try: line = int(r[0])
    except ValueError: continue
    region = maps.get(cur_file, {}).get(line, cur_file)
    if region == "__launch_bounds__": region = "kernel body"
    for h, v in zip(hdr, r):
        if h.startswith("stall_") and "(Not Issued)" not in h:
            try: x = int(v)
            except ValueError: continue
            agg[region][h[6:]] += x; tot[h[6:]] += x
total = sum(tot.values())
print("kernel %s: %d samples; overall: %s" % (kern, total, " ".join("%s %.1f%%" % (k, 100.0 * v / total) for k, v in tot.most_common(9))))
for region, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values())):
    n = sum(c.values())
    if n < 0.01 * total: continue
    print("%5.1f%%  %-18s %s" % (100.0 * n / total, region, " ".join("%s %.0f%%" % (k, 100.0 * v / n) for k, v in c.most_common(6))))
