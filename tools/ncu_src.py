#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report: joins ncu's per-SASS-instruction metrics (--page source --csv) with
nvdisasm's line table of the same kernel in libsmj.so (built with -lineinfo), by instruction order.

    python tools/ncu_src.py gpurun_out/prof_x.ncu-rep <kernel substring> [top_n] [launch index]

The library must be the build the report was captured from."""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel block only
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "ins": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r and r[0].startswith("0x"):
        cur["ins"].append(r)
blk = blocks[0]
h = blk["hdr"]
ii, si = h.index("Instructions Executed"), h.index("# Samples")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pim-sort-merge-join_b200", "libsmj.so")], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    secs = re.split(r"\n//-+ \.text\.", dis)
    for s in secs[1:]:
        head = s.split("\n", 1)[0]
        if kname in head:
            cand, line = [], None
            for ln in s.split("\n"):
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                if m:
                    line = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    cand.append((line, m.group(2).strip()))
            if len(cand) == len(blk["ins"]):
                lines = cand
                break
            if lines is None or abs(len(cand) - len(blk["ins"])) < abs(len(lines) - len(blk["ins"])):
                lines = cand
    if lines and len(lines) == len(blk["ins"]):
        break
if not lines:
    sys.exit(f"kernel {kname} not found in libsmj.so")
if len(lines) != len(blk["ins"]):
    print(f"warning: {len(lines)} SASS instructions in libsmj.so vs {len(blk['ins'])} in the report (different build?)")
agg = {}
tot_i = tot_s = 0
for (line, _), r in zip(lines, blk["ins"]):
    i, s = int(r[ii] or 0), int(r[si] or 0)
    a = agg.setdefault(line, [0, 0])
    a[0] += i
    a[1] += s
    tot_i += i
    tot_s += s
print(f"== {blk['name'][:100]}\n   {tot_i} warp-instructions, {tot_s} stall samples")
src_cache = {}
best = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
for (line, (i, s)) in sorted(best, key=lambda kv: (kv[0] or ("", 0))):
    text = ""
    if line:
        f = os.path.join(ROOT, "pim-sort-merge-join_b200", "csrc", line[0])
        if f not in src_cache and os.path.exists(f):
            src_cache[f] = open(f).read().split("\n")
        if f in src_cache and line[1] - 1 < len(src_cache[f]):
            text = src_cache[f][line[1] - 1].strip()
    print(f"{(line[0] + ':' + str(line[1])) if line else '?':22s} inst {100.0 * i / max(tot_i, 1):5.1f}%  samp {100.0 * s / max(tot_s, 1):5.1f}%  {text[:100]}")
