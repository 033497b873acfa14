"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log into profiles/ form.
    python tools/launch_list.py gpurun_out/launches_r1.csv > profiles/r1_launch_list.csv
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
cmd = sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
print(f"# command: {cmd}  (under ncu --metrics gpu__time_duration.sum --clock-control none -c 600; "
      "per-launch times are serialised and cold-cache: the kernels' SHARES are what compares)")
print("id,kernel,block,grid,gpu__time_duration.sum [ns]")
tot = collections.Counter()
for r in rows:
    name = r[4].split("(")[0][-60:]
    ns = int(float(r[14].replace(",", "")))
    print(f"{r[0]},{name},{r[7].replace(',', ' ')},{r[8].replace(',', ' ')},{ns}")
    tot[name] += ns
print("# share of GPU time by kernel")
s = sum(tot.values())
for k, v in tot.most_common():
    print(f"# {100 * v / s:6.2f} %  {v / 1e6:10.3f} ms  {k}")
