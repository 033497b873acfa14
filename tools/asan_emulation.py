import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import tests.hostemu as he
he._OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', '_build', 'libba_hostemu_asan.so')
he.build = lambda: he._OUT
from tests.util import Golden, golden_names, std_noise
from tests.test_host_side import _burst_day
for name in [n for n in golden_names() if n != 'wn100']:
    g = Golden(name)
    plan = he.EmuPlan(g.specs, len(g.global_codes), g.delta_t, g.max_t, g.include_cancellations)
    lat = {k: v[None] for k, v in g.latents.items()}
    for exact in (False, True):
        o = plan.forward(lat, g.scales, g.noise[None], exact=exact, shuffle_seed=5, want_pop=True, replay_waits=exact)
        assert (o['status'] == 0).all(), (name, o['status'])
    o = plan.forward(lat, g.scales, None, seed=7)
    if plan.scan2_ok:
        # the ring-less scan: shuffled lanes, with / without diagnostics, Philox, value mode with
        # value gradients
        for pop in (True, False):
            o = plan.forward(lat, g.scales, g.noise[None], scan2=True, shuffle_seed=5, want_pop=pop)
            assert (o['status'] == 0).all(), (name, 'scan2', o['status'])
        o = plan.forward(lat, g.scales, None, seed=7, scan2=True)
        vals = np.nan_to_num(g.ref_values, nan=1.0)[None]
        o = plan.forward(lat, g.scales, vals, value_mode=True, scan2=True, want_value_grad=True)
        assert (o['status'] == 0).all(), (name, 'scan2 value mode', o['status'])
os.environ['BA_Q_MIN'] = '1'; os.environ['BA_Q_DIV'] = '1e9'
g = Golden('congested6')
plan = he.EmuPlan(g.specs, len(g.global_codes), g.delta_t, g.max_t, False)
o = plan.forward({k: v[None] for k, v in g.latents.items()}, g.scales, g.noise[None], rerun_overflow=False)
assert (o['status'] == 0).all()
del os.environ['BA_Q_MIN'], os.environ['BA_Q_DIV']
spec = _burst_day(); N = spec.n_airports
plan = he.EmuPlan([spec], N, 0.2, 30.0, False)
lat = {"mean_turnaround": np.full((1, N), 0.5, np.float32), "mean_service": np.full((1, N), 0.01, np.float32), "travel": np.full((1, N, N), 2.0, np.float32)}
o = plan.forward(lat, (0.1, 0.1, 0.1), std_noise(1, 1, spec.n_flights, 8))
print('asan run finished: all statuses clean')
