"""One-off GPU stress: random networks of awkward sizes against the C oracle
(logp, gradients, tick counts), both list variants, with and without cancellations."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesair_b200.compiler import day_spec_from_frame
from bayesair_b200.plan import SimulationPlan
from bayesair_b200.synth import synth_day_frame, synth_latents
from oracle import ba_oracle
from tests.util import std_noise

dev = torch.device("cuda:0")
bad = 0
for N, F, canc, svc in ((3, 40, False, 0.05), (31, 600, False, 0.03), (33, 700, True, 0.03),
                        (64, 1500, False, 0.02), (65, 1500, False, 0.05), (129, 2500, False, 0.02),
                        (200, 4000, True, 0.02), (257, 5000, False, 0.02)):
    specs = [day_spec_from_frame(synth_day_frame(N, F, d, seed=100 + N + d)) for d in range(2)]
    codes = specs[0].codes
    for s in specs[1:]:
        codes = codes + [c for c in s.codes if c not in codes]
    specs = [day_spec_from_frame(synth_day_frame(N, F, d, seed=100 + N + d), codes) for d in range(2)]
    P = 6
    lat = synth_latents(len(codes), P, seed=N, service_scale=svc, include_cancellations=canc,
                        n_aircraft_mean=6.0)
    Fm = max(s.n_flights for s in specs)
    noise = std_noise(P, 2, Fm, seed=N)
    sc = (0.1, 0.08, 0.12)
    ref = ba_oracle.run_batch(specs, lat, sc, noise, delta_t=0.2, include_cancellations=canc,
                              want_grad=True, n_threads=8)
    plan = SimulationPlan(specs, codes, delta_t=0.2, include_cancellations=canc, device=dev)
    tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
    args = [tl["mean_turnaround"], tl["mean_service"], tl["travel"]]
    if canc:
        args += [tl["log_initial_aircraft"], tl["base_cancel_logprob"]]
    la, lr = plan.pack_latents(*args)
    sct = torch.tensor(np.asarray(sc, np.float32), device=dev)
    for exact in (False, True):
        out = plan.forward(la, lr, sct, torch.tensor(noise, device=dev), want_grad=True,
                           want_ticks=True, force_exact_lists=exact)
        torch.cuda.synchronize()
        lp = out["logp"].double().cpu().numpy()
        e_lp = np.abs(lp - ref["logp"]).max() / np.abs(ref["logp"]).max()
        ga, gr, gs = plan.backward(torch.ones_like(out["logp"]), out["tape"])
        g_serv = ga[:, 1].double().cpu().numpy()
        e_g = np.abs(g_serv - ref["grads"]["mean_service"]).max() / (np.abs(ref["grads"]["mean_service"]).max() + 1e-30)
        ticks_ok = np.array_equal(out["ticks"].cpu().numpy(), ref["n_ticks"])
        st = int(out["status"].abs().sum())
        ok = e_lp < 5e-6 and e_g < 1e-4 and ticks_ok and st == 0
        bad += not ok
        print(f"N={N} F={F} canc={canc} exact={exact}: logp {e_lp:.1e} grad {e_g:.1e} ticks {ticks_ok} status {st} {'ok' if ok else 'FAIL'}", flush=True)
    plan.close()
print("failures:", bad)
