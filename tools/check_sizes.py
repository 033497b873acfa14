"""GPU sanity run over the BASELINE configs that are not the bench workload:
cfg1 (top-10), cfg2 (one day, 1024 particles) and a reduced cfg5 (N=500, F=20000):
parity of a few particle-days against the C oracle, then throughput."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_workload, std_noise  # noqa: E402
from bayesair_b200.plan import SimulationPlan  # noqa: E402
from bayesair_b200.synth import synth_latents  # noqa: E402
from oracle import ba_oracle  # noqa: E402

dev = torch.device("cuda:0")
for name, P, days in (("cfg1", 1024, 1), ("cfg2", 1024, 1), ("cfg5s", 256, 2)):
    N, Fs, _, svc, specs = build_workload(name)
    specs = specs[:days]
    t0 = time.time()
    plan = SimulationPlan(specs, specs[0].codes, delta_t=0.2, device=dev)
    t_plan = time.time() - t0
    lat = synth_latents(N, P, seed=4, service_scale=svc)
    tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
    la, lr = plan.pack_latents(tl["mean_turnaround"], tl["mean_service"], tl["travel"])
    sc = torch.tensor([0.1, 0.1, 0.1], device=dev)
    Pc = 4
    nz = std_noise(Pc, days, max(s.n_flights for s in specs), 5)
    o = plan.forward(la[:Pc].contiguous(), lr[:Pc].contiguous(), sc, torch.tensor(nz, device=dev),
                     want_grad=True, want_ticks=True)
    ga, gr, gs = plan.backward(torch.ones_like(o["logp"]), o["tape"])
    torch.cuda.synchronize()
    ref = ba_oracle.run_batch(specs, {k: v[:Pc] for k, v in lat.items()}, (0.1, 0.1, 0.1), nz,
                              delta_t=0.2, n_threads=8)
    lp = o["logp"].cpu().numpy().astype(np.float64)
    err = np.abs(lp - ref["logp"]).max() / np.abs(ref["logp"]).max()
    gerr = np.abs(ga[:, 1].cpu().numpy() - ref["grads"]["mean_service"]).max() / \
        np.abs(ref["grads"]["mean_service"]).max()
    ticks_ok = np.array_equal(o["ticks"].cpu().numpy(), ref["n_ticks"])
    noise = plan.generate_noise(P, 9)
    ones = torch.ones((P, days), device=dev)
    for _ in range(2):
        oo = plan.forward(la, lr, sc, noise, want_grad=True)
        plan.backward(ones, oo["tape"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        oo = plan.forward(la, lr, sc, noise, want_grad=True)
        plan.backward(ones, oo["tape"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    bad = int(oo["status"].ne(0).sum())
    print(f"{name}: N={N} F={Fs[0]} days={days} P={P} block={plan.block_threads} smem={plan.smem_bytes} "
          f"ctas={plan.resident_ctas} plan={t_plan:.2f}s | logp rel err {err:.1e} grad {gerr:.1e} "
          f"ticks_equal={ticks_ok} | {ms:.2f} ms/step {P * days / ms * 1e3:.0f} pd/s bad_status={bad}",
          flush=True)
    plan.close()

# BASELINE config 4: posterior-predictive forward simulation (all observations stripped,
# cancellations on) of the 10-day network
N, Fs, _, svc, specs = build_workload("cfg3")
import copy
pspecs = []
for sp in specs[:4]:
    q = copy.deepcopy(sp)
    q.act_dep[:] = np.nan; q.act_arr[:] = np.nan; q.cancel_obs[:] = -1
    pspecs.append(q)
P = 2048
plan = SimulationPlan(pspecs, pspecs[0].codes, delta_t=0.2, include_cancellations=True, device=dev)
lat = synth_latents(N, P, seed=4, service_scale=svc, include_cancellations=True, n_aircraft_mean=8.0)
tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
la, lr = plan.pack_latents(tl["mean_turnaround"], tl["mean_service"], tl["travel"],
                           tl["log_initial_aircraft"], tl["base_cancel_logprob"])
sc = torch.tensor([0.1, 0.1, 0.1], device=dev)
for _ in range(2):
    o = plan.predict(la, lr, sc, None, None, seed=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    o = plan.predict(la, lr, sc, None, None, seed=3)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
sim = o["sim"]
canc = float((sim[..., 0] == 1).float().mean())
delay = float(torch.nanmean(sim[..., 1] - torch.tensor(np.stack([np.pad(s.sched_dep, (0, plan.max_flights - s.n_flights)) for s in pspecs]), device=dev)[None]))
print(f"cfg4 (predictive, cancellations on): P={P} days={len(pspecs)} {ms:.1f} ms/step "
      f"{P * len(pspecs) / ms * 1e3:.0f} pd/s bad_status={int(o['status'].ne(0).sum())} "
      f"cancelled fraction {canc:.3f} mean departure delay {delay:.2f} h", flush=True)
