"""Value-mode gradient fixtures: the UNMODIFIED reference with its four latent per-flight
sites conditioned on given values (autograd leaves), as a guide over those sites would do
(``scripts/wn/wn_network.py:285-302``).  Written next to the fixtures of make_golden.py as
``<name>_valuegrad.npz``:

    values        [D, Fmax, 4]  the site values (the draws of the matching golden)
    total         the reference's per-flight log-probability at those values
    grad_values   [D, Fmax, 4]  its gradient w.r.t. the values
    grad_<latent>, grad_scales  its gradient w.r.t. the global latents / the three scales

    python tests/golden/make_value_golden.py [names]
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bayesair_b200.synth import synth_day_frame, synth_latents  # noqa: E402
from oracle.reference_runner import run_reference  # noqa: E402
from tests.golden.make_golden import CASES  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ("tiny4", "twoday5", "congested6", "saturated5", "obscancel4", "top10")


def make(name):
    N, Fs, svc, canc, nac, cfrac, scales, dt, seed = CASES[name]
    assert not canc, "value gradients are defined for plans without aircraft bookkeeping"
    frames = [synth_day_frame(N, F, d, seed=seed, cancel_frac=cfrac) for d, F in enumerate(Fs)]
    lat = {k: v[0] for k, v in synth_latents(N, 1, seed=seed, service_scale=svc).items()}
    gold = np.load(os.path.join(HERE, name + ".npz"))
    values = gold["ref_values"]
    ref = run_reference(frames, lat, None, delta_t=dt, scales=scales, values=values)
    assert abs(ref["per_flight"] - float(gold["ref_per_flight"])) <= 2e-5 * abs(ref["per_flight"]), \
        (ref["per_flight"], float(gold["ref_per_flight"]))
    out = {"values": values, "total": ref["per_flight"], "grad_values": ref["grad_values"],
           "grad_scales": ref["grad_scales_per_flight"]}
    for k, v in ref["grad_per_flight"].items():
        out["grad_" + k] = v
    path = os.path.join(HERE, name + "_valuegrad.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: total={ref['per_flight']:.4f} |grad_values|max={np.abs(ref['grad_values']).max():.3f} "
          f"-> {os.path.getsize(path)} B")


if __name__ == "__main__":
    for n in sys.argv[1:] or NAMES:
        make(n)
