"""Small fixed invocation of the hot path for ncu / timing:
    python tools/profile_forward.py [P] [service_scale] [days] [iters]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_workload  # noqa: E402
from bayesair_b200.plan import SimulationPlan  # noqa: E402
from bayesair_b200.synth import synth_latents  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
svc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
days = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
exact = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
N, Fs, _, _, specs = build_workload("cfg3")
specs = specs[:days]
dev = torch.device("cuda:0")
plan = SimulationPlan(specs, specs[0].codes, delta_t=0.2, device=dev)
lat = synth_latents(N, P, seed=1, service_scale=svc)
tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
la, lr = plan.pack_latents(tl["mean_turnaround"], tl["mean_service"], tl["travel"])
sc = torch.tensor([0.1, 0.1, 0.1], device=dev)
nz = plan.generate_noise(P, 3)
ones = torch.ones((P, days), device=dev)
for it in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = plan.forward(la, lr, sc, nz, want_grad=(os.environ.get("BA_NOGRAD") is None), no_overflow_rerun=True, want_ticks=True,
                     force_exact_lists=exact)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if o["tape"] is not None:
        plan.backward(ones, o["tape"])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
ov = int((o["status"] & 2).ne(0).sum())
print(f"block={plan.block_threads} exact={exact} P={P} days={days} svc={svc} smem={plan.smem_bytes} ctas={plan.resident_ctas} "
      f"fwd={1e3 * (t1 - t0):.2f} ms bwd={1e3 * (t2 - t1):.2f} ms overflow={ov}/{P * days} "
      f"ticks~{float(o['ticks'].float().mean()):.0f} pd/s(fwd)={P * days / (t1 - t0):.0f}")
