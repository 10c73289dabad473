#!/usr/bin/env python
"""Per-source-line executed warp instructions / stall samples of one kernel from an ncu capture (--import-source not needed):
SASS instructions of the capture's source page are matched, in order, with `nvdisasm -g` line info of the shipped cubin.
Usage: python tools/ncu_lines.py REP.ncu-rep KERNEL_MANGLED_PREFIX CUBIN_STEM [top_n]      (e.g. _ZN3pgs6k_emit compact)"""
import collections, csv, os, re, subprocess, sys
rep, kpref, stem = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("NCU_LINES_SO") or os.path.join(ROOT, "incubator_pegasus_b200", "libpegasus_b200.so")
os.system(f"rm -rf /tmp/xelf && mkdir -p /tmp/xelf && cd /tmp/xelf && cuobjdump -xelf all {so} >/dev/null 2>&1")
cub = [f for f in os.listdir("/tmp/xelf") if f.startswith(stem) and f.endswith(".cubin")][0]
DEPTH = int(os.environ.get("NCU_LINES_OUTER", "0"))  # >0: attribute to the DEPTH-th frame from the outside of the inline chain
sass = subprocess.check_output(["nvdisasm", "-gi" if DEPTH else "-g", "-c", os.path.join("/tmp/xelf", cub)]).decode().split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(kpref) and l.rstrip().endswith(":")][0]
end = len(sass)
for i in range(start + 1, len(sass)):
    if sass[i].startswith("//--------------------- .text."): end = i; break
insts, cur, chain, in_chain = [], None, [], False
for l in sass[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        if not in_chain: chain = []
        in_chain = True
        chain.append(loc)  # innermost first; the following lines walk outwards
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+\S", l):
        if in_chain:
            cur = chain[0]
            if DEPTH and len(chain) >= DEPTH: cur = chain[-DEPTH]
            in_chain = False
        insts.append((cur, l.split("*/", 1)[1].strip()[:60]))
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ix, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
print(f"{len(insts)} SASS instructions in the cubin, {len(data)} in the capture")
agg, samp, per_inst = collections.Counter(), collections.Counter(), []
for (c, txt), d in zip(insts, data):
    agg[c] += int(d[ix]); samp[c] += int(d[isamp]); per_inst.append((int(d[ix]), int(d[isamp]), c, txt))
tot, ts = sum(agg.values()), sum(samp.values())
files = {}
def srcline(c):
    if not c: return ""
    f, l = c
    if f not in files:
        for d in ("csrc", "host"):
            p = os.path.join(ROOT, "incubator_pegasus_b200", d, f)
            if os.path.exists(p): files[f] = open(p).read().split("\n"); break
        else: files[f] = []
    return files[f][l - 1].strip()[:110] if 0 < l <= len(files[f]) else ""
print(f"total warp instructions {tot}, stall samples {ts}")
for c, n in agg.most_common(topn):
    print(f"{100*n/tot:5.1f}% inst {100*samp[c]/max(1,ts):5.1f}% samp  {c[0] if c else '?':22s}:{c[1] if c else 0:4d}  {srcline(c)}")
if os.environ.get("NCU_LINES_STATIC"):
    st = collections.Counter()
    for n, s, c, txt in per_inst: st[c] += 1
    print("\nstatic SASS count / executions per instruction, top lines:")
    for c, n in agg.most_common(topn):
        print(f"  {c[0] if c else '?':22s}:{c[1] if c else 0:4d}  static {st[c]:4d}  exec/inst {n/max(1,st[c])/1e6:8.2f} M   {srcline(c)[:80]}")
