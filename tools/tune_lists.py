"""GPU tuning helper: shared-memory list capacities vs overflow rate vs throughput.
    python tools/tune_lists.py [P] [days]
"""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_workload  # noqa: E402
from bayesair_b200.plan import SimulationPlan  # noqa: E402
from bayesair_b200.synth import synth_latents  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
days = int(sys.argv[2]) if len(sys.argv) > 2 else 3
svcs = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.02]
N, Fs, _, svc, specs = build_workload("cfg3")
specs = specs[:days]
dev = torch.device("cuda:0")
grid = [(16, 8, 0, 0), (16, 4, 0, 0), (32, 4, 0, 0), (8, 8, 0, 0)]
for spread, svc in itertools.product((0.3,), svcs):
    lat = synth_latents(N, P, seed=1, service_scale=svc, spread=spread)
    tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
    for qdiv, qmin, bfrac, bmin in grid:
        os.environ.update(BA_Q_DIV=str(qdiv), BA_Q_MIN=str(qmin), BA_BAG_FRAC=str(bfrac),
                          BA_BAG_MIN=str(bmin))
        plan = SimulationPlan(specs, specs[0].codes, delta_t=0.2, device=dev)
        la, lr = plan.pack_latents(tl["mean_turnaround"], tl["mean_service"], tl["travel"])
        sc = torch.tensor([0.1, 0.1, 0.1], device=dev)
        nz = plan.generate_noise(P, 3)
        for rerun in (False, True):
            for _ in range(2):
                o = plan.forward(la, lr, sc, nz, want_grad=True, no_overflow_rerun=not rerun)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                o = plan.forward(la, lr, sc, nz, want_grad=True, no_overflow_rerun=not rerun)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            ov = int((o["status"] & 2).ne(0).sum())
            print(f"svc={svc} spread={spread} qdiv={qdiv} qmin={qmin} bag={bfrac}/{bmin} smem={plan.smem_bytes} "
                  f"ctas={plan.resident_ctas} rerun={rerun} overflow={ov}/{P * days} ms={ms:.2f} "
                  f"pd/s={P * days / ms * 1e3:.0f}", flush=True)
        plan.close()
