#!/bin/bash
# Rebuild profiles/ from the artefacts a GPU run left in gpurun_out/ (see DESIGN.md section 7):
#   gpurun_out/launches_r1.csv      ncu --metrics gpu__time_duration.sum of the bench command
#   gpurun_out/prof_r1_forward.ncu-rep   ncu --set full of tools/profile_forward.py 2960 0.012 2 3
set -e
cd "$(dirname "$0")/.."
LIB=bayesair_b200/_lib/libbayesair_b200.so
K=ba_forward_kernelILb0ELi128ELb0ELb0
REP=gpurun_out/prof_r1_forward.ncu-rep
python tools/launch_list.py gpurun_out/launches_r1.csv > profiles/r1_launch_list.csv
{
  python tools/ncu_summary.py $REP 2>/dev/null | grep -v "^source columns" | grep -v "^warning" | sed '/^total stall samples/,$d'
  echo; echo "== stall reasons per code region (tools/ncu_ranges.py) =="
  python tools/ncu_ranges.py $REP $LIB $K $(python - <<'PY'
import re
src = open("bayesair_b200/csrc/ba_sim.cuh").read().splitlines()
marks = [("ba_prologue_flight", "prologue"), ("ba_q_push", "ring_push_refill"), ("ba_take_aircraft", "tracked_airports"),
         ("ba_phase_a", "phase_A_service"), ("ba_calendar_scan", "calendar"), ("ba_copy8", "phase_B_staging"),
         ("ba_due_order", "phase_C_pull"), ("ba_airport_busy", "vote"), ("ba_epilogue_flight", "epilogue_1"),
         ("ba_epilogue_airport", "epilogue_2")]
pos = []
for name, label in marks:
    for i, l in enumerate(src):
        if re.match(r"(BA_DEV|template).*", l) and (name + "(") in (l + (src[i + 1] if i + 1 < len(src) else "")):
            pos.append((i + 1, label)); break
pos.sort()
out = []
for (lo, label), nxt in zip(pos, pos[1:] + [(len(src) + 1, "")]):
    out.append(f"ba_sim.cuh:{lo}-{nxt[0] - 1}={label}")
out.append("ba_kernels.cu:1-1000=kernel_body_and_barriers")
print(" ".join(out))
PY
)
  echo; echo "== hottest source lines, all stall samples (tools/ncu_lines.py) =="
  python tools/ncu_lines.py $REP $LIB $K 20 2>&1
  echo; echo "== hottest source lines outside barrier waits (critical path of a tick) =="
  python tools/ncu_lines.py $REP $LIB $K 25 busy 2>&1
} > profiles/r1_forward_kernel_ncu_summary.txt
python - <<'PY'
import subprocess, csv, io, json
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_r1_forward.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
def get(k):
    i = h.index(k)
    return float(v[i].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u[i], 1)
r, w = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
d = {"workload": "cfg3", "source": "profiles/r1_forward_kernel_ncu_summary.txt",
     "capture": "ncu --set full, tools/profile_forward.py 2960 0.012 2 3 (5920 particle-days)",
     "dram_bytes_read": r, "dram_bytes_write": w, "particle_days": 5920,
     "dram_bytes_per_particle_day": (r + w) / 5920}
json.dump(d, open("profiles/forward_kernel_summary.json", "w"), indent=1)
print(d)
PY
