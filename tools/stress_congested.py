"""One-off GPU check: the bench network (N=100, 2 days) at three congestion levels against the
C oracle -- log-prob, d/d mean_service, tick counts; the heaviest level runs past the 384-tick
calendar and takes the exact-list re-run."""
import sys, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from bench import build_workload, std_noise
from bayesair_b200.plan import SimulationPlan
from bayesair_b200.synth import synth_latents
from oracle import ba_oracle
dev = torch.device("cuda:0")
N, Fs, _, svc, specs = build_workload("cfg3")
specs = specs[:2]
for scale in (0.012, 0.05, 0.1):
    P = 48
    plan = SimulationPlan(specs, specs[0].codes, delta_t=0.2, device=dev)
    lat = synth_latents(N, P, seed=77, service_scale=scale)
    tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
    la, lr = plan.pack_latents(tl["mean_turnaround"], tl["mean_service"], tl["travel"])
    sc = torch.tensor([0.1, 0.1, 0.1], device=dev)
    nz = std_noise(P, 2, max(s.n_flights for s in specs), 9)
    out = plan.forward(la, lr, sc, torch.tensor(nz, device=dev), want_grad=True, want_ticks=True)
    torch.cuda.synchronize()
    ref = ba_oracle.run_batch(specs, lat, (0.1, 0.1, 0.1), nz, delta_t=0.2, want_grad=True, n_threads=16)
    lp = out["logp"].double().cpu().numpy()
    e = np.abs(lp - ref["logp"]).max() / np.abs(ref["logp"]).max()
    ga, gr, gs = plan.backward(torch.ones_like(out["logp"]), out["tape"])
    eg = np.abs(ga[:, 1].double().cpu().numpy() - ref["grads"]["mean_service"]).max() / np.abs(ref["grads"]["mean_service"]).max()
    print(f"service {scale}: logp rel {e:.1e} grad rel {eg:.1e} ticks equal {np.array_equal(out['ticks'].cpu().numpy(), ref['n_ticks'])} "
          f"max ticks {int(out['ticks'].max())} status {int(out['status'].abs().sum())}", flush=True)
    plan.close()
