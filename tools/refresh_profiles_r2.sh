#!/bin/bash
# Rebuild the round-2 files of profiles/ from what tools/make_profiles_r2.sh left in gpurun_out/:
#   launches_r2.csv              ncu --metrics gpu__time_duration.sum of `bench.py --steps 2 --warmup 3`
#   prof_r2_forward2.ncu-rep     ncu --set full of tools/profile_forward.py 4736 0.012 2 3
set -e
cd "$(dirname "$0")/.."
LIB=bayesair_b200/_lib/libbayesair_b200.so
K=ba_forward2_kernelILi1EE
REP=${1:-gpurun_out/prof_r2_forward2.ncu-rep}
TAG=${2:-r2}
python tools/launch_list.py gpurun_out/launches_r2.csv > profiles/${TAG}_launch_list.csv
{
  python tools/ncu_summary.py $REP 2>/dev/null | grep -v "^source columns" | grep -v "^warning" | sed '/^total stall samples/,$d'
  echo; echo "== stall reasons per code region (tools/ncu_ranges.py) =="
  python tools/ncu_ranges.py $REP $LIB $K $(python - <<'PY'
import re
src = open("bayesair_b200/csrc/ba_scan2.cuh").read().splitlines()
marks = [("ba2_flight", "table_loads"), ("ba2_init_airport", "airport_setup"), ("ba2_prologue_load", "prologue_flight"),
         ("ba2_bucket_scan", "sort_bucket_scan"), ("ba2_scatter_flight", "sort_scatter"), ("ba2_split_cursor", "sort_split"),
         ("ba2_state_init", "scan_state_init"), ("ba2_insert", "scan_insert"), ("ba2_replay", "scan_replay"),
         ("ba2_airport_tick", "scan_airport_tick"), ("ba2_value_grad_flight", "value_gradients"),
         ("ba2_epilogue_load", "epilogue_flight"),
         ("ba2_epilogue_airport", "epilogue_sums")]
pos = []
for name, label in marks:
    for i, l in enumerate(src):
        if re.match(r"(BA_DEV|template|void|float|static|struct Ba2FlightIn|struct Ba2EpiIn).*", l) and (name + "(") in l:
            pos.append((i + 1, label)); break
pos.sort()
out = []
for (lo, label), nxt in zip(pos, pos[1:] + [(len(src) + 1, "")]):
    out.append(f"ba_scan2.cuh:{lo}-{nxt[0] - 1}={label}")
out.append("ba_kernels.cu:1-2000=kernel_body")
out.append("ba_sim.cuh:1-3000=shared_helpers(philox,log-probs)")
print(" ".join(out))
PY
)
  echo; echo "== hottest source lines, all stall samples (tools/ncu_lines.py) =="
  python tools/ncu_lines.py $REP $LIB $K 30 2>&1
} > profiles/${TAG}_forward2_kernel_ncu_summary.txt
python - "$REP" "$TAG" <<'PY'
import subprocess, csv, io, json, sys
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
def get(k):
    i = h.index(k)
    return float(v[i].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u[i], 1)
r, w = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
pd = int(get("launch__grid_size")) * 2
d = {"workload": "cfg3", "source": f"profiles/{tag}_forward2_kernel_ncu_summary.txt",
     "capture": f"ncu --set full, tools/profile_forward.py 4736 0.012 2 3 ({pd} particle-days)",
     "dram_bytes_read": r, "dram_bytes_write": w, "particle_days": pd,
     "dram_bytes_per_particle_day": (r + w) / pd}
json.dump(d, open("profiles/forward_kernel_summary.json", "w"), indent=1)
print(d)
PY
