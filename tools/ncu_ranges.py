"""Stall-reason breakdown per source-line range of ba_sim.cuh / ba_kernels.cu.
    python tools/ncu_ranges.py <report.ncu-rep> <lib.so> <kernel-substring> file:lo-hi[=label] ...
"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1:4]
ranges = []
for a in sys.argv[4:]:
    spec, _, label = a.partition("=")
    f, _, r = spec.partition(":")
    lo, _, hi = r.partition("-")
    ranges.append((f, int(lo), int(hi), label or spec))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, ("?", 0)
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m: infn = kern in m.group(1); continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
    if lines: break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
idx = {c: hdr.index(c) for c in cols}
ci, ce = hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > ce]
agg = collections.defaultdict(lambda: collections.Counter())
for r, (f, n) in zip(body, lines):
    lab = "other"
    for (rf, lo, hi, label) in ranges:
        if f == rf and lo <= n <= hi: lab = label; break
    agg[lab]["inst"] += float(r[ce] or 0); agg[lab]["samples"] += float(r[ci] or 0)
    for c in cols: agg[lab][c] += float(r[idx[c]] or 0)
tot_s = sum(a["samples"] for a in agg.values()); tot_i = sum(a["inst"] for a in agg.values())
for lab, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(((c, a[c]) for c in cols), key=lambda kv: -kv[1])[:5]
    print(f"{lab:28s} samples {100*a['samples']/tot_s:5.1f}%  inst {100*a['inst']/tot_i:5.1f}%  " +
          "  ".join(f"{c[6:]}={100*v/max(a['samples'],1):.0f}%" for c, v in top))
