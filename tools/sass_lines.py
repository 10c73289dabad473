#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (nvdisasm -g line info of the shipped cubin).  For a kernel whose
hot loop body executes straight through, static counts of the loop's lines approximate the per-iteration instruction cost.
Usage: python tools/sass_lines.py KERNEL_MANGLED_PREFIX CUBIN_STEM [top_n]"""
import collections, os, re, subprocess, sys
kpref, stem = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
OUTER = bool(os.environ.get("SASS_OUTER"))
RANGE = os.environ.get("SASS_RANGE")  # "file:lo-hi": also print the total inside that line range
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "incubator_pegasus_b200", "libpegasus_b200.so")
os.system(f"rm -rf /tmp/xelf && mkdir -p /tmp/xelf && cd /tmp/xelf && cuobjdump -xelf all {so} >/dev/null 2>&1")
cub = [f for f in os.listdir("/tmp/xelf") if f.startswith(stem) and f.endswith(".cubin")][0]
sass = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join("/tmp/xelf", cub)]).decode().split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(kpref) and l.rstrip().endswith(":")][0]
end = len(sass)
for i in range(start + 1, len(sass)):
    if sass[i].startswith("//--------------------- .text."): end = i; break
agg, ops, cur, n = collections.Counter(), collections.defaultdict(collections.Counter), None, 0
for l in sass[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        if OUTER:  # attribute to the outermost call site (the line of the kernel's own body that the code was inlined into)
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            if inl: cur = (os.path.basename(inl[-1][0]), int(inl[-1][1]))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m: agg[cur] += 1; ops[cur][m.group(1)] += 1; n += 1
files = {}
def srcline(c):
    if not c: return ""
    f, l = c
    if f not in files:
        for d in ("csrc", "host"):
            p = os.path.join(ROOT, "incubator_pegasus_b200", d, f)
            if os.path.exists(p): files[f] = open(p).read().split("\n"); break
        else: files[f] = []
    return files[f][l - 1].strip()[:95] if 0 < l <= len(files[f]) else ""
print(f"{n} SASS instructions")
for c, k in agg.most_common(topn):
    top = ",".join(f"{o}{v}" for o, v in ops[c].most_common(4))
    print(f"{k:5d}  {c[0] if c else '?':22s}:{c[1] if c else 0:4d}  [{top:40s}] {srcline(c)}")

if RANGE:
    f, r = RANGE.split(":"); lo, hi = map(int, r.split("-"))
    tot = sum(k for c, k in agg.items() if c and c[0] == f and lo <= c[1] <= hi)
    print(f"total inside {RANGE}: {tot} instructions")
