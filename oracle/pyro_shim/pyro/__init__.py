"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.

A minimal stand-in for the ``pyro`` package surface that the reference's
``bayes_air`` package touches (pyro-ppl 1.9.0 is pinned by the reference's
``poetry.lock`` but is not installable in this image: no network, no wheel).
It exists so that the *unmodified* reference files under ``/root/reference``
can be imported and executed in this container to (a) generate the golden
vectors committed under ``tests/golden/`` and (b) validate the C restatement
in ``oracle/ba_oracle.c``.  It is put on ``sys.path`` only by
``oracle/reference_runner.py``.

Surface provided (everything the reference calls on the hot path):
  * ``pyro.param`` / ``pyro.sample`` / ``pyro.factor`` / ``pyro.markov`` /
    ``pyro.plate``                       (bayes_air/model.py:41-126)
  * ``pyro.poutine.trace / condition / scale`` and ``Trace.log_prob_sum``
                                        (scripts/wn/train.py:306-314)
  * ``pyro.distributions.{Gamma, Normal, Exponential,
    RelaxedBernoulliStraightThrough, constraints}``
                                        (bayes_air/network.py:102-109,
                                         bayes_air/types/airport.py:144-220)

One extension that real Pyro does not have: ``pyro.set_noise_source(fn)``.
When set, the standard-normal / standard-exponential draw behind every
reparameterised ``Normal`` / ``Exponential`` sample statement is taken from
``fn(site_name, kind)`` instead of the global torch RNG.  The draw is then
pushed through exactly the arithmetic torch uses (``loc + eps * scale`` and
``e / rate``), so that the CUDA path and the reference can be driven with
bit-identical per-flight noise.
"""
from __future__ import annotations

import collections

import torch
from torch.distributions import constraints as _constraints
from torch.distributions import transform_to as _transform_to

__version__ = "0.0-bayesair-oracle-shim"
__bayesair_oracle_shim__ = True

# --------------------------------------------------------------------------
# handler stack
# --------------------------------------------------------------------------
_STACK: list = []
_NOISE_SOURCE = None


def set_noise_source(fn) -> None:
    """Install (or clear with None) the per-site standard-noise provider."""
    global _NOISE_SOURCE
    _NOISE_SOURCE = fn


def _apply_stack(msg: dict) -> dict:
    # innermost handler sees the message first, like Pyro's messenger stack
    for handler in reversed(_STACK):
        handler.process(msg)
    if not msg["done"]:
        if msg["type"] == "sample" and msg["value"] is None:
            fn = msg["fn"]
            if _NOISE_SOURCE is not None and hasattr(fn, "draw_with"):
                msg["value"] = fn.draw_with(_NOISE_SOURCE, msg["name"])
            else:
                msg["value"] = fn()
        msg["done"] = True
    for handler in _STACK:
        handler.postprocess(msg)
    return msg


class _Handler:
    def __init__(self, fn=None):
        self.fn = fn

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        assert _STACK[-1] is self
        _STACK.pop()
        return False

    def __call__(self, *args, **kwargs):
        with self:
            return self.fn(*args, **kwargs)

    def process(self, msg):
        pass

    def postprocess(self, msg):
        pass


# --------------------------------------------------------------------------
# param store
# --------------------------------------------------------------------------
class _ParamStore:
    def __init__(self):
        self.unconstrained = collections.OrderedDict()
        self.constraints = {}

    def clear(self):
        self.unconstrained.clear()
        self.constraints.clear()

    def get(self, name, init=None, constraint=_constraints.real):
        if name not in self.unconstrained:
            if init is None:
                raise KeyError(name)
            with torch.no_grad():
                raw = _transform_to(constraint).inv(init.detach().clone())
            raw = raw.clone().requires_grad_(True)
            self.unconstrained[name] = raw
            self.constraints[name] = constraint
        raw = self.unconstrained[name]
        value = _transform_to(self.constraints[name])(raw)
        value.unconstrained = lambda: raw
        return value

    def named_parameters(self):
        return self.unconstrained.items()

    def __contains__(self, name):
        return name in self.unconstrained

    def save(self, path):
        torch.save({k: v.detach() for k, v in self.unconstrained.items()}, path)


_PARAMS = _ParamStore()


def get_param_store():
    return _PARAMS


def clear_param_store():
    _PARAMS.clear()


def set_rng_seed(seed: int):
    torch.manual_seed(seed)


def enable_validation(flag: bool = True):
    return None


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def param(name, init_tensor=None, constraint=_constraints.real, event_dim=None):
    value = _PARAMS.get(name, init_tensor, constraint)
    msg = {
        "type": "param", "name": name, "fn": None, "value": value,
        "is_observed": False, "scale": 1.0, "done": True,
    }
    _apply_stack(msg)
    return msg["value"]


def sample(name, fn, obs=None, infer=None):
    msg = {
        "type": "sample", "name": name, "fn": fn, "value": obs,
        "is_observed": obs is not None, "scale": 1.0, "done": False,
    }
    if not _STACK:
        if obs is not None:
            return obs
        if _NOISE_SOURCE is not None and hasattr(fn, "draw_with"):
            return fn.draw_with(_NOISE_SOURCE, name)
        return fn()
    _apply_stack(msg)
    return msg["value"]


class _Unit:
    """log-factor carrier, mirrors pyro.distributions.Unit for pyro.factor."""

    def __init__(self, log_factor):
        self.log_factor = log_factor

    def log_prob(self, value):
        return self.log_factor

    def __call__(self):
        return torch.zeros(())


def factor(name, log_factor, has_rsample=None):
    return sample(name, _Unit(log_factor), obs=torch.zeros(()))


def markov(iterable=None, history=1, keep=False):
    return iterable


class plate:  # noqa: N801  (Pyro's name)
    def __init__(self, name, size=None, dim=None):
        self.name, self.size = name, size

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        return iter(range(self.size))


from . import distributions  # noqa: E402
from . import poutine  # noqa: E402

__all__ = [
    "param", "sample", "factor", "markov", "plate", "poutine", "distributions",
    "get_param_store", "clear_param_store", "set_rng_seed", "enable_validation",
    "set_noise_source",
]
