"""Static SASS instruction mix per kernel of a built library.  Usage: sass_mix.py lib.so [name-substring ...]"""
import collections, re, subprocess, sys
lib, pats = sys.argv[1], sys.argv[2:]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; mix = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur: mix[cur][m.group(1).split(".")[0]] += 1
for fn, c in mix.items():
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    if pats and not any(p in dem for p in pats): continue
    tot = sum(c.values())
    print("%s: %d instr | %s" % (dem, tot, " ".join("%s=%d" % kv for kv in c.most_common(14))))
