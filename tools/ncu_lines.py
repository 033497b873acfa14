"""Attribute ncu per-SASS-instruction samples to CUDA source lines.
    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top] [column]
`column` ranks by one stall column of the source page (e.g. stall_long_sb); the
pseudo-column `busy` = all samples minus stall_barrier (where the warps that are NOT
parked at a barrier spend their time, i.e. the critical path of a tick).
Joins `ncu --page source --csv` (SASS rows, in address order) with the line table
`nvdisasm -g` prints for the same function (needs -lineinfo at compile time).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
col = sys.argv[5] if len(sys.argv) > 5 else None

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []          # (file, line) per instruction, in order
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True,
                         text=True).stdout
    infn, cur = False, ("?", 0)
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = kern in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if lines:
        break

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci, ce, cs = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
body = [r for r in rows[2:] if len(r) > ce]
if len(body) != len(lines):
    print(f"warning: {len(body)} SASS rows vs {len(lines)} disassembled instructions", file=sys.stderr)
samp, inst = collections.Counter(), collections.Counter()
cb = hdr.index("stall_barrier")
cc = hdr.index(col) if col and col != "busy" else None
for r, loc in zip(body, lines):
    if col == "busy":
        samp[loc] += float(r[ci] or 0) - float(r[cb] or 0)
    elif cc is not None:
        samp[loc] += float(r[cc] or 0)
    else:
        samp[loc] += float(r[ci] or 0)
    inst[loc] += float(r[ce] or 0)
ts, ti = sum(samp.values()) or 1, sum(inst.values()) or 1
src_cache = {}
def text(loc):
    f, n = loc
    if f not in src_cache:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bayesair_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    s = src_cache[f]
    return s[n - 1].strip()[:110] if 0 < n <= len(s) else ""
print(f"total stall samples {ts:.0f}, warp instructions executed {ti:.0f}")
print(f"{'samples%':>8} {'inst%':>6}  location")
for loc, v in samp.most_common(top):
    print(f"{100 * v / ts:8.1f} {100 * inst[loc] / ti:6.1f}  {loc[0]}:{loc[1]}  {text(loc)}")
