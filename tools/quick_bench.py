"""Forward + backward timing of one bench workload (no extras):
    python tools/quick_bench.py <workload> <particles> [service_scale] [iters]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_workload  # noqa: E402
from bayesair_b200.plan import SimulationPlan  # noqa: E402
from bayesair_b200.synth import synth_latents  # noqa: E402

name = sys.argv[1]
N, Fs, P, svc, specs = build_workload(name)
P = int(sys.argv[2]) if len(sys.argv) > 2 else P
svc = float(sys.argv[3]) if len(sys.argv) > 3 else svc
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
plan = SimulationPlan(specs, specs[0].codes, delta_t=0.2, device=dev)
lat = synth_latents(N, P, seed=1, service_scale=svc)
tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
la, lr = plan.pack_latents(tl["mean_turnaround"], tl["mean_service"], tl["travel"])
sc = torch.tensor([0.1, 0.1, 0.1], device=dev)
nz = plan.generate_noise(P, 3)
ones = torch.ones((P, len(Fs)), device=dev)
best = 1e9
for it in range(iters + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = plan.forward(la, lr, sc, nz, want_grad=True, want_ticks=True)
    plan.backward(ones, o["tape"])
    torch.cuda.synchronize()
    if it:
        best = min(best, time.perf_counter() - t0)
bad = int((o["status"] != 0).sum())
print(f"{name} N={N} P={P} svc={svc} block={plan.block_threads} smem={plan.smem_bytes} ctas={plan.resident_ctas} "
      f"step={1e3 * best:.2f} ms pd/s={P * len(Fs) / best:.0f} ticks~{float(o['ticks'].float().mean()):.0f} bad={bad}")
