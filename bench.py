#!/usr/bin/env python
"""Benchmark of the hot path: WN-network particle-days/s (log-joint + gradient).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--workload cfg3|cfg2|cfg1|cfg5s]

One "step" = one pass of the hot path over one batch: forward simulation +
fused log-joint (`ba_forward`) and the VJP contraction (`ba_backward`) for
P particles x D days of a synthetic WN-like network (the WN data is absent from
the reference mount).  Prints ONE JSON line (contract in the task
statement): `value` = whole-job particle-days/s with inputs resident in HBM,
`e2e` = the same through the host-buffer C-ABI call `ba_logjoint_host` (H2D of
latents, both kernels, D2H of log-probs + gradients inside the timed region),
`roofline` for the forward kernel, `cpu_baseline` = the C oracle on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_days_per_s"
UNIT = "particle-days/s"

# name -> (N airports, flights per day, particles per GPU, service scale)
WORKLOADS = {
    # BASELINE.json configs[2] (the north star's stated target size): full WN network,
    # Dec 21-30 2022 meltdown flight counts, 4096 particles x 10 days
    "cfg3": (100, [3876, 3876, 3876, 3200, 3552, 3876, 3876, 3876, 3876, 3876], 4096, 0.012),
    # configs[1]: full WN network, one day, 1024 particles
    "cfg2": (100, [3876], 1024, 0.012),
    # configs[0]: top-10 airports, one day (the reference's own CPU-runnable case)
    "cfg1": (10, [480], 1024, 0.03),
    # configs[4] scaled to fit a quick run: 500 airports, 20k flights/day, 4 days; 1332
    # particles = four full waves of the 1332 particle-days a B200 holds at this size (the full
    # config is 184 waves: a batch of one and a half waves would measure the idle tail)
    "cfg5s": (500, [20000] * 4, 1332, 0.01),
}


def build_workload(name, seed=21):
    from bayesair_b200.compiler import day_spec_from_frame
    from bayesair_b200.synth import synth_day_frame
    N, Fs, P, svc = WORKLOADS[name]
    frames = [synth_day_frame(N, F, d, seed=seed) for d, F in enumerate(Fs)]
    specs = [day_spec_from_frame(frames[0])]
    specs += [day_spec_from_frame(f, specs[0].codes) for f in frames[1:]]
    return N, Fs, P, svc, specs


def workload_config(name, particles=0):
    """The `config` both arms print (identical dicts: the driver compares them)."""
    N, Fs, P, svc = WORKLOADS[name]
    P = particles or P
    return {"workload": f"{name}: WN-like synthetic network N={N} airports, D={len(Fs)} days, "
                        f"F/day={Fs}, P={P} particles per GPU, delta_t=0.2, no cancellations, "
                        f"mean service {svc} h x LogNormal(0, 0.3), network seed 21",
            "particles_per_gpu": P, "days": len(Fs), "delta_t": 0.2,
            "l2_policy": "inputs larger than L2 (the per-flight noise tensor alone is "
                         f"{P * sum(Fs) * 16 / 1e9:.2f} GB per GPU)"}


def algorithmic_bytes_per_particle_day(N, Fs, R, with_noise=True):
    """16 F (noise in) + 4 (2N+R) (latents in) + 4 (2N+R) (grads out)
    + 4 (logp) + 12 (scale grads), static flight tables amortised over particles."""
    F = float(np.mean(Fs))
    return (16.0 * F if with_noise else 0.0) + 8.0 * (2 * N + R) + 16.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:                                   # set up before the timed region starts
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self._nvml = (pynvml, h, reasons, mx)
        except Exception:
            self._nvml = None

    def _run(self):
        # NVML in-process (a sample every 20 ms); nvidia-smi (a few per second) if it is absent
        if self._nvml is not None:
            pynvml, h, get_reasons, mx = self._nvml
            bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40),
                    ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
            try:
                while True:
                    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    r = int(get_reasons(h))
                    self.rows.append([str(sm), str(mx), "0"] +
                                     ["Active" if r & b else "Not Active" for _, b in bits])
                    if self._stop.wait(0.02):
                        break
                return
            except Exception:
                pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def std_noise(P, D, F, seed):
    rng = np.random.default_rng([seed, 99])
    z = np.empty((P, D, F, 4), np.float32)
    z[..., 0] = rng.standard_exponential((P, D, F), dtype=np.float32)
    z[..., 1] = rng.standard_normal((P, D, F), dtype=np.float32)
    z[..., 2] = rng.standard_exponential((P, D, F), dtype=np.float32)
    z[..., 3] = rng.standard_normal((P, D, F), dtype=np.float32)
    return z


def cpu_baseline(specs, N, svc, n_particles, threads, seed=3):
    """The C oracle (a port of the reference algorithm) on the host cores, on a
    bounded sample of the same workload."""
    from bayesair_b200.synth import synth_latents
    from oracle import ba_oracle
    lat = synth_latents(N, n_particles, seed=seed, service_scale=svc)
    Fmax = max(s.n_flights for s in specs)
    noise = std_noise(n_particles, len(specs), Fmax, seed)
    t0 = time.perf_counter()
    out = ba_oracle.run_batch(specs, lat, (0.1, 0.1, 0.1), noise, delta_t=0.2, want_grad=True,
                              n_threads=threads)
    dt = time.perf_counter() - t0
    assert (out["status"] == 0).all()
    return n_particles * len(specs) / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The
    reference is pure Python + Pyro and cannot travel to the GPU box (Pyro is not
    installable, /root/reference is not mounted there), so this arm times the C
    port of its algorithm (oracle/ba_oracle.c, pinned to the reference by golden
    vectors) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    N, Fs, P, svc, specs = build_workload(name)
    cores = os.cpu_count() or 1
    # bounded sample per step: about 2-4 s of CPU work
    probe_rate, _ = cpu_baseline(specs, N, svc, cores, cores)
    n_particles = max(1, int(probe_rate * 3.0 / len(specs)))
    for _ in range(args.warmup):
        cpu_baseline(specs, N, svc, max(1, n_particles // 4), cores)
    t_tot = 0.0
    for k in range(args.steps):
        _, dt = cpu_baseline(specs, N, svc, n_particles, cores, seed=10 + k)
        t_tot += dt
    value = args.steps * n_particles * len(specs) / t_tot
    sample = f"{n_particles} particles x {len(specs)} days per step ({name} network), {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args.particles),
        "reference_kind": f"C port of the reference algorithm (oracle/ba_oracle.c), {cores} host threads",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C port of the reference algorithm (oracle/ba_oracle.c); the reference's own "
                "Pyro path measured 13.7 s per particle-day (0.073 particle-days/s, 1 core) "
                "at this network size in the build container (tests/golden/wn100.npz ref_times)",
    }
    print(json.dumps(line))


def run_native(args):
    import torch
    import torch.distributed as dist

    from bayesair_b200 import _capi
    from bayesair_b200.plan import SimulationPlan
    from bayesair_b200.synth import synth_latents

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; bayesair_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    N, Fs, P, svc, specs = build_workload(name)
    if args.particles:
        P = args.particles
    D = len(Fs)
    plan = SimulationPlan(specs, specs[0].codes, delta_t=0.2, device=dev)
    R, C = plan.n_routes, plan.n_airport_latents
    lat = synth_latents(N, P, seed=1000 + rank, service_scale=svc)
    tl = {k: torch.tensor(v, device=dev) for k, v in lat.items()}
    la, lr = plan.pack_latents(tl["mean_turnaround"], tl["mean_service"], tl["travel"])
    la, lr = la.contiguous(), lr.contiguous()
    scales = torch.tensor([0.1, 0.1, 0.1], device=dev)
    noise = plan.generate_noise(P, seed=77 + rank)        # resident in HBM, like the latents
    ones = torch.ones((P, D), device=dev)
    # guide-parameter gradient buffer of a mean-field guide over 2N + N^2 log-latents
    # (scripts/wn/train.py:228-232): the only thing that crosses GPUs per step
    n_latent = 2 * N + N * N
    guide_grad = torch.zeros(2 * n_latent + 4, device=dev)

    def step():
        out = plan.forward(la, lr, scales, noise, want_grad=True)
        ga, gr, gs = plan.backward(ones, out["tape"])
        if world > 1:
            guide_grad[:N * 2] = ga.sum(0).reshape(-1)[:N * 2]
            dist.all_reduce(guide_grad)
        return out, ga, gr, gs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out, ga, gr, gs = step()
    barrier()
    SimulationPlan.check_status(out["status"])

    launches0 = _capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _capi.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * P * D / (ms_per_step * 1e-3)

    # ---- strong scaling: the SAME total batch (P particles x D days) split over the ranks --
    strong = {"value": value, "particles_total": P, "ms_per_step": ms_per_step,
              "note": "fixed total work split over n_gpus ranks (contiguous particle shards)"}
    if world > 1:
        from bayesair_b200.parallel import shard_particles
        s0, cnt = shard_particles(P, rank, world)
        la_s, lr_s = la[s0:s0 + cnt].contiguous(), lr[s0:s0 + cnt].contiguous()
        nz_s, ones_s = noise[s0:s0 + cnt], ones[:cnt]

        def strong_step():
            o = plan.forward(la_s, lr_s, scales, nz_s, want_grad=True)
            g3 = plan.backward(ones_s, o["tape"])
            guide_grad[:N * 2] = g3[0].sum(0).reshape(-1)[:N * 2]
            dist.all_reduce(guide_grad)

        for _ in range(3):
            strong_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            strong_step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong["ms_per_step"] = float(t.item())
        strong["value"] = P * D / (strong["ms_per_step"] * 1e-3)
        del la_s, lr_s

    # ---- forward kernel alone (roofline): one launch per call, no re-run kernels --
    fwd_ms = []
    for _ in range(max(3, min(args.steps, 10))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.forward(la, lr, scales, noise, want_grad=True, no_overflow_rerun=True)
        e1.record()
        torch.cuda.synchronize()
        fwd_ms.append(e0.elapsed_time(e1))
    fwd = float(np.mean(fwd_ms))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    B = algorithmic_bytes_per_particle_day(N, Fs, R)
    achieved = B * P * D / (fwd * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "forward_kernel_summary.json")))
        if prof.get("workload") == name:
            # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of
            # this kernel on this network, per particle-day, scaled to this launch's P x D
            traffic = prof["dram_bytes_per_particle_day"] * P * D
    except Exception:
        pass

    # ---- same workload, prior-like (congested) service times: every hub queue saturates,
    #      the shared-memory rings overflow and the exact-list variant does the work ----
    lat_c = synth_latents(N, P, seed=2000 + rank, service_scale=0.05)
    tc = {k: torch.tensor(v, device=dev) for k, v in lat_c.items()}
    la_c, lr_c = plan.pack_latents(tc["mean_turnaround"], tc["mean_service"], tc["travel"])
    la_c, lr_c = la_c.contiguous(), lr_c.contiguous()
    for _ in range(2):
        o = plan.forward(la_c, lr_c, scales, noise, want_grad=True)
        plan.backward(ones, o["tape"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_c = max(2, args.steps // 3)
    for _ in range(n_c):
        o = plan.forward(la_c, lr_c, scales, noise, want_grad=True)
        plan.backward(ones, o["tape"])
    e1.record()
    torch.cuda.synchronize()
    congested_value = P * D * n_c / (e0.elapsed_time(e1) * 1e-3)
    del la_c, lr_c, tc

    # ---- Philox variant (no noise tensor read): production configuration --------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(max(2, args.steps // 2)):
        o = plan.forward(la, lr, scales, None, seed=5, want_grad=True)
        plan.backward(ones, o["tape"])
    e1.record()
    torch.cuda.synchronize()
    philox_value = P * D * max(2, args.steps // 2) / (e0.elapsed_time(e1) * 1e-3)

    # ---- "SVI step ms" (BASELINE metric, cfg2 shape): mean-field guide over the 2N+N^2
    #      log-latents, 1024 particles, one day: guide sample -> model log-joint (kernel) ->
    #      backward -> gradient all-reduce -> Adam step (scripts/wn/train.py:301-326 batched)
    svi = None
    if not args.no_svi:
        from bayesair_b200.objective import MeanFieldGuide, n_latent, objective_fn
        from bayesair_b200.parallel import allreduce_gradients
        from bayesair_b200.svi import FusedMeanFieldSVI
        plan1 = SimulationPlan(specs[:1], specs[0].codes, delta_t=0.2, device=dev)
        init = torch.zeros(n_latent(N), device=dev)
        init[:N] = float(np.log(0.5)); init[N:2 * N] = float(np.log(svc))
        init[2 * N:] = torch.log(torch.tensor(np.maximum(lat["travel"][0], 0.5), device=dev)).reshape(-1)

        def time_steps(fn, n=20, warm=6):
            for _ in range(warm):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # fused: ba_mf_sample -> ba_forward -> ba_mf_backward -> all-reduce(flat) -> fused Adam,
        # replayed as one CUDA graph; the status flag is read once after the timed steps
        fused = FusedMeanFieldSVI(plan1, scales, 1024, init_loc=init, init_scale=0.05, lr=1e-3, seed=11)
        fused_ms = time_steps(fused.step)
        fused.check()
        graph = fused._graph is not None
        # strong: the same 1024 particles split over the ranks
        strong_ms = fused_ms
        if world > 1:
            fs = FusedMeanFieldSVI(plan1, scales, max(1, 1024 // world), init_loc=init,
                                   init_scale=0.05, lr=1e-3, seed=11)
            strong_ms = time_steps(fs.step)
            fs.check()
            del fs
        # framework-autograd variant of the same step (round-1 path, for comparison)
        guide = MeanFieldGuide(n_latent(N), init_loc=init, init_scale=0.05).to(dev)
        opt = torch.optim.Adam(guide.parameters(), lr=1e-3)

        def svi_step():
            opt.zero_grad(set_to_none=True)
            loss = objective_fn(guide, plan1, scales, 1024)
            loss.backward()
            allreduce_gradients(guide.parameters(), average=True)
            opt.step()
            return loss

        auto_ms = time_steps(svi_step, n=10, warm=3)
        svi = {"value": fused_ms, "strong_value": strong_ms, "autograd_value": auto_ms,
               "cuda_graph": graph,
               "config": "cfg2: N=100, one day F=3876, mean-field guide over 10200 log-latents, Adam. "
                         "value: 1024 particles per GPU, fused library step (sample, simulate, guide "
                         "backward into the flat all-reduce buffer, fused Adam) as one CUDA graph, "
                         "status flag checked after the timed steps; strong_value: 1024 particles "
                         "in total split over the GPUs; autograd_value: the same step through "
                         "torch autograd with a host status check per step (round-1 path)"}
        del fused
        plan1.close()

    # ---- CalNF step (BASELINE configs[2]): conditional spline-flow guide over the 10 200
    #      log-latents, K = 2 bootstrap permutations of the 10 failure days, 2 nominal days,
    #      P particles per ELBO: label substep + 2 permutation ELBOs + nominal ELBO +
    #      calibration ELBO (scripts/training.py:177-296), ONE all-reduce of the flat
    #      flow-gradient buffer, clipping, fused Adam ---------------------------------------
    calnf = None
    if not args.no_calnf and name == "cfg3":
        try:
            from bayesair_b200.calnf import CalNFTrainer
            from bayesair_b200.compiler import day_spec_from_frame
            from bayesair_b200.flows import ConditionalSplineFlow
            from bayesair_b200.objective import n_latent as n_lat
            from bayesair_b200.synth import synth_day_frame
            torch.manual_seed(1234)                      # identical flow weights on every rank
            tf32_prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = True  # conditioner GEMMs (K = 64) on tensor cores
            nom = [day_spec_from_frame(synth_day_frame(N, 3876, 20 + d, seed=21), specs[0].codes) for d in range(2)]
            nplan = SimulationPlan(nom, specs[0].codes, delta_t=0.2, device=dev)
            init = torch.zeros(n_lat(N))
            init[:N] = float(np.log(0.5)); init[N:2 * N] = float(np.log(svc))
            init[2 * N:] = torch.log(torch.tensor(np.maximum(lat["travel"][0], 0.5))).reshape(-1)
            flow = ConditionalSplineFlow(n_lat(N), 2, hidden_features=(64, 64), transforms=4, bins=8,
                                         init_loc=init, init_scale=0.05).to(dev)
            Pc = args.calnf_particles
            tr = CalNFTrainer(flow, plan, nplan, scales, n_particles=Pc, n_permutations=2, lr=1e-4,
                              calibration_lr=1e-3, seed=5)
            for _ in range(2):
                tr.step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n_c = 3
            for _ in range(n_c):
                tr.step()
            e1.record()
            barrier()
            tr.check()
            t = torch.tensor([e0.elapsed_time(e1) / n_c], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sim_pd = Pc * (4 * D + len(nom))             # label + 2 permutations + calibration, nominal
            calnf = {"value": float(t.item()), "flow_parameters": int(tr.fp.flat.numel()),
                     "allreduce_bytes_per_step": int(tr.fp.nbytes) if world > 1 else 0,
                     "simulated_particle_days_per_step_per_gpu": int(sim_pd),
                     "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
                     "config": f"cfg3 network, coupling spline flow (4 layers, 8 bins, hidden 64x64) over "
                               f"{n_lat(N)} log-latents, K=2 permutations, {Pc} particles per ELBO per GPU, "
                               "5 ELBO evaluations per step (label substep, 2 permutations, nominal, "
                               "calibration); conditioner GEMMs cuBLAS TF32, splines ba_rqs_* kernels"}
            del tr, flow
            nplan.close()
            torch.cuda.empty_cache()
            torch.backends.cuda.matmul.allow_tf32 = tf32_prev
        except Exception as exc:      # an optional section must not take the headline line down
            calnf = {"value": None, "error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()

    # ---- posterior-predictive forward simulation (BASELINE configs[3]): every observation
    #      stripped, cancellations on (aircraft bookkeeping at every airport), 10 days ------
    predictive = None
    if not args.no_predictive and name == "cfg3":
        try:
            import copy
            pspecs = []
            for sp in specs:
                q = copy.deepcopy(sp)
                q.act_dep[:] = np.nan; q.act_arr[:] = np.nan; q.cancel_obs[:] = -1
                pspecs.append(q)
            Pp = args.predictive_samples
            pplan = SimulationPlan(pspecs, pspecs[0].codes, delta_t=0.2, include_cancellations=True, device=dev)
            plat = synth_latents(N, Pp, seed=4 + rank, service_scale=svc, include_cancellations=True,
                                 n_aircraft_mean=8.0)
            pt = {k: torch.tensor(v, device=dev) for k, v in plat.items()}
            pla, plr = pplan.pack_latents(pt["mean_turnaround"], pt["mean_service"], pt["travel"],
                                          pt["log_initial_aircraft"], pt["base_cancel_logprob"])
            for _ in range(2):
                po = pplan.predict(pla, plr, scales, None, None, seed=3)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                po = pplan.predict(pla, plr, scales, None, None, seed=3)
            e1.record()
            barrier()
            SimulationPlan.check_status(po["status"])
            pms = e0.elapsed_time(e1) / 2
            if world > 1:
                t = torch.tensor([pms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                pms = float(t.item())
            predictive = {"value": world * Pp * D / (pms * 1e-3), "unit": UNIT, "ms_per_step": pms,
                          "cancelled_fraction": float((po["sim"][..., 0] == 1).float().mean()),
                          "config": f"cfg4 at a reduced sample count: the cfg3 network, all 10 days, every "
                                    f"observation stripped, cancellations on, {Pp} samples per GPU (BASELINE "
                                    "configs[3] asks for 65 536: the same kernel, 8x the work), in-kernel Philox"}
            del pla, plr, po
            pplan.close()
        except Exception as exc:      # an optional section must not take the headline line down
            predictive = {"value": None, "error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()

    # ---- the drop-in model() call itself (cfg2 shape): 2N + N(N-1) pyro.sample statements
    #      under trace(condition(...)), scripts/wn/train.py:306-314 -------------------------
    model_call_ms = None
    if not args.no_svi:
        try:
            import bayesair_b200 as ba
            from bayesair_b200.ppl import poutine, pyro
            from bayesair_b200.synth import synth_day_frame as _frame
            fl, aps = ba.parse_schedule(_frame(N, Fs[0], 0, seed=21))
            states = [ba.NetworkState({a.code: a for a in aps}, fl)]
            codes = list(states[0].airports.keys())
            Pm = 1024
            zt = torch.tensor(lat["mean_turnaround"][:Pm], device=dev)
            zs = torch.tensor(lat["mean_service"][:Pm], device=dev)
            ztr = torch.tensor(lat["travel"][:Pm], device=dev)
            data = {}
            for i, c in enumerate(codes):
                data[f"{c}_mean_turnaround_time"] = zt[:, i]
                data[f"{c}_mean_service_time"] = zs[:, i]
                for j, c2 in enumerate(codes):
                    if i != j:
                        data[f"travel_time_{c}_{c2}"] = ztr[:, i, j]

            def call():
                tr_ = poutine.trace(poutine.condition(ba.air_traffic_network_model, data=data)).get_trace(
                    states, 0.2, device=dev)
                return tr_.log_prob_sum()
            pyro.clear_param_store()
            call()                                           # compiles and caches the plan
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            model_call_ms = 1e3 * (time.perf_counter() - t0) / 3
            from bayesair_b200.plan import clear_plan_cache
            clear_plan_cache()
        except Exception as exc:      # optional section
            model_call_ms = None
            print(f"[bench] model-call timing skipped: {type(exc).__name__}: {exc}", file=sys.stderr)

    # ---- end to end through the host-buffer C ABI --------------------------------
    h_la = torch.empty(la.shape, dtype=torch.float32, pin_memory=True).copy_(la.cpu())
    h_lr = torch.empty(lr.shape, dtype=torch.float32, pin_memory=True).copy_(lr.cpu())
    h_sc = np.array([0.1, 0.1, 0.1], np.float32)
    outbuf = {
        "logp": torch.empty((P, D), dtype=torch.float32, pin_memory=True).numpy(),
        "status": torch.empty((P, D), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
        "g_airport": torch.empty((P, C, N), dtype=torch.float32, pin_memory=True).numpy(),
        "g_route": torch.empty((P, R), dtype=torch.float32, pin_memory=True).numpy(),
        "g_scales": torch.empty((P, 3), dtype=torch.float32, pin_memory=True).numpy(),
    }
    for _ in range(2):
        plan.logjoint_host(h_la.numpy(), h_lr.numpy(), h_sc, None, seed=5, out=outbuf)
    barrier()
    n_e2e = max(2, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        plan.logjoint_host(h_la.numpy(), h_lr.numpy(), h_sc, None, seed=5, out=outbuf)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * P * D * n_e2e / e2e_s
    h2d = int(h_la.numel() * 4 + h_lr.numel() * 4 + 12)
    d2h = int(sum(v.nbytes for v in outbuf.values()))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(name, args.particles),
            "launch": {
                "routes": R, "kernel": "ba_forward2_kernel (one warp per particle-day up to 128 "
                                       "airports, NW warps beyond)",
                "noise": "explicit [P,D,F,4] fp32 tensor resident in HBM "
                         f"({noise.numel() * 4 / 1e9:.2f} GB per GPU)",
                "parallelism": f"particles sharded over {world} GPU(s); one all-reduce of the "
                               f"{guide_grad.numel()}-float mean-field guide-gradient buffer per step"
                               if world > 1 else "single GPU",
                "block_threads": plan.block_threads, "smem_bytes": plan.smem_bytes,
                "resident_ctas": plan.resident_ctas,
            },
            "reference_kind": "the --impl reference arm and cpu_baseline are a C port of the "
                              "reference algorithm (oracle/ba_oracle.c) on all host threads; the "
                              "Pyro reference itself cannot run on the GPU box",
            "strong": strong,
            "clocks": clocks.summary(),
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "note": "ba_logjoint_host: pinned host latents in, Philox noise in-kernel, "
                            "log-probs + gradients out, synchronised"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "ba_forward2_kernel", "kernel_ms": fwd,
                         "algorithmic_bytes_per_particle_day": B,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"
                         if peaks else "fallback 6650 GB/s (of fallback)",
                         "note": "serial tick loop per particle-day: memory-latency bound by "
                                 "construction (SURVEY.md §8d; every pipe below 30 %, DRAM at 23 % "
                                 "of peak, 1.5 MB of DRAM traffic per particle-day against 74 KB "
                                 "algorithmic); see profiles/ and DESIGN.md §7"},
            "svi_step_ms": svi,
            "calnf_step_ms": calnf,
            "predictive": predictive,
            "model_call_ms": {"value": model_call_ms,
                              "config": "drop-in air_traffic_network_model under trace(condition(...)) + "
                                        "log_prob_sum, cfg2 shape (N=100, 1024 particles): 10 100 "
                                        "pyro.sample statements on the host around one kernel call"},
            "philox_value": philox_value,
            "congested_value": congested_value,
            "regimes": {"value": f"posterior-like latents: mean service {svc} h x LogNormal(0, 0.3); "
                                 "no particle-day is re-run",
                        "congested_value": "prior-like latents: mean service 0.05 h, every hub "
                                           "queue saturates (~230 ticks per day instead of ~137, "
                                           "draw histories replayed); same kernel, no re-runs"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            probe, _ = cpu_baseline(specs, N, svc, cores, cores)
            n_cpu = max(cores, int(probe * 12.0 / D))          # about 12 s of CPU work
            rate, dt = cpu_baseline(specs, N, svc, n_cpu, cores)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n_cpu} particles x {D} days of the same network in {dt:.1f} s "
                          f"(oracle/ba_oracle.c, {cores} threads)",
                "reference_python_recorded": "13.7 s per particle-day = 0.073 particle-days/s on 1 "
                                             "core for the unmodified Pyro reference at N=100/F=3876 "
                                             "(build container, tests/golden/wn100.npz ref_times)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-svi", action="store_true")
    ap.add_argument("--no-calnf", action="store_true")
    ap.add_argument("--no-predictive", action="store_true")
    ap.add_argument("--predictive-samples", type=int, default=8192, help="predictive samples per GPU")
    ap.add_argument("--calnf-particles", type=int, default=1024, help="particles per ELBO per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
