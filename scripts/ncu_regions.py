"""Per-region instruction shares from an ncu report (regions = line ranges per file). Usage: ncu_regions.py rep kernel"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = None; hdr = None
per = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1].split("(")[0].split("::")[-1].split("<")[0]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] == "" or cur_fn != kern: continue
    try: line = int(r[0])
    except ValueError: continue
    def g(name):
        try: return int(r[hdr[name]])
        except Exception: return 0
    per[(cur_file, line)][0] += g("Instructions Executed"); per[(cur_file, line)][1] += g("# Samples"); per[(cur_file, line)][2] += g("L1 Wavefronts Shared")
import re
def fn_ranges(path):
    # crude: map each line to the enclosing __device__/__global__ function name
    names = {}; cur = "?"
    for i, l in enumerate(open(path), 1):
        m = re.search(r"(?:__device__|__global__)[^;(]*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m: cur = m.group(1)
        names[i] = cur
    return names
import os
base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mgnet_b200", "csrc")
maps = {f: fn_ranges(os.path.join(base, f)) for f in os.listdir(base) if f.endswith(".cuh")}
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
for (f, ln), v in per.items():
    key = (f, maps.get(f, {}).get(ln, "?"))
    for k in range(3): agg[key][k] += v[k]; tot[k] += v[k]
print("kernel %s: %d warp-instr %d samples %d smem wavefronts" % (kern, tot[0], tot[1], tot[2]))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%5.1f%% inst %5.1f%% smp %5.1f%% smemwf  %s:%s" % (100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1), 100.0 * v[2] / max(tot[2], 1), key[0], key[1]))
