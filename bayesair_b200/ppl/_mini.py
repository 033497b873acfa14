"""A small effect-handler runtime with Pyro's names (used only when ``pyro`` is
not installed).  Implements what the WN path of the reference and its drivers
call: ``sample`` / ``param`` / ``factor`` / ``plate`` / ``markov``, a global param
store with constraints, and the ``trace`` / ``condition`` / ``scale`` /
``block`` / ``replay`` handlers with ``Trace.log_prob_sum``.

Design: a stack of ``Messenger`` objects; a primitive builds a message dict and
walks the stack bottom-up (``_process_message``) then top-down
(``_postprocess_message``), exactly the protocol Pyro documents in
``pyro/poutine/runtime.py`` (re-implemented here from its public description,
not copied).
"""
from __future__ import annotations

import types
from collections import OrderedDict
from typing import Callable, Dict, Optional

import torch
import torch.distributions as _td
from torch.distributions import constraints, transform_to

__version__ = "bayesair_b200-mini"

_STACK = []


# ---------------------------------------------------------------------------
# distributions
# ---------------------------------------------------------------------------
class _Callable:
    """``dist()`` draws (reparameterised when possible), like Pyro's mixin."""

    def __call__(self, sample_shape=torch.Size()):
        return self.rsample(sample_shape) if self.has_rsample else self.sample(sample_shape)


class _Gamma(_Callable, _td.Gamma):
    pass


class _Normal(_Callable, _td.Normal):
    pass


class _Exponential(_Callable, _td.Exponential):
    pass


class _LogNormal(_Callable, _td.LogNormal):
    pass


class _MultivariateNormal(_Callable, _td.MultivariateNormal):
    pass


class _RelaxedBernoulliStraightThrough(_Callable, _td.RelaxedBernoulli):
    def rsample(self, sample_shape=torch.Size()):
        soft = _td.utils.clamp_probs(super().rsample(sample_shape))
        hard = (soft.round() - soft).detach() + soft
        hard._unquantize = soft
        return hard

    def log_prob(self, value):
        return super().log_prob(getattr(value, "_unquantize", value))


class _Unit:
    """Carrier of an arbitrary log-factor (``pyro.factor``)."""

    has_rsample = False

    def __init__(self, log_factor):
        self.log_factor = log_factor

    def log_prob(self, value):
        return self.log_factor

    def __call__(self, sample_shape=torch.Size()):
        return torch.zeros((), device=getattr(self.log_factor, "device", None))


class _Delta:
    has_rsample = True

    def __init__(self, v, log_density=0.0, event_dim=0):
        self.v, self.log_density, self.event_dim = v, log_density, event_dim

    def log_prob(self, value):
        shape = self.v.shape[:self.v.dim() - self.event_dim]
        return torch.zeros(shape, device=self.v.device, dtype=self.v.dtype) + self.log_density

    def __call__(self, sample_shape=torch.Size()):
        return self.v


distributions = types.SimpleNamespace(
    Gamma=_Gamma, Normal=_Normal, Exponential=_Exponential, LogNormal=_LogNormal,
    MultivariateNormal=_MultivariateNormal,
    RelaxedBernoulliStraightThrough=_RelaxedBernoulliStraightThrough,
    Unit=_Unit, Delta=_Delta, constraints=constraints,
)


# ---------------------------------------------------------------------------
# param store
# ---------------------------------------------------------------------------
class ParamStore:
    def __init__(self):
        self._raw: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self._constraint: Dict[str, object] = {}

    def clear(self):
        self._raw.clear()
        self._constraint.clear()

    def __contains__(self, name):
        return name in self._raw

    def setdefault(self, name, init, constraint):
        if name not in self._raw:
            value = init() if callable(init) else init
            with torch.no_grad():
                raw = transform_to(constraint).inv(value.detach()).clone()
            self._raw[name] = raw.requires_grad_(True)
            self._constraint[name] = constraint
        return self[name]

    def __getitem__(self, name):
        raw = self._raw[name]
        value = transform_to(self._constraint[name])(raw)
        value.unconstrained = lambda: raw
        return value

    def get_param(self, name):
        return self[name]

    def named_parameters(self):
        return self._raw.items()

    def keys(self):
        return self._raw.keys()

    def get_state(self):
        return {"params": {k: v.detach().clone() for k, v in self._raw.items()},
                "constraints": dict(self._constraint)}

    def set_state(self, state):
        self.clear()
        for k, v in state["params"].items():
            self._raw[k] = v.clone().requires_grad_(True)
            self._constraint[k] = state["constraints"][k]

    def save(self, path):
        torch.save(self.get_state(), path)

    def load(self, path, map_location=None):
        self.set_state(torch.load(path, map_location=map_location, weights_only=False))


_PARAM_STORE = ParamStore()


def get_param_store() -> ParamStore:
    return _PARAM_STORE


def clear_param_store():
    _PARAM_STORE.clear()


def set_rng_seed(seed: int):
    torch.manual_seed(seed)


def enable_validation(is_validate: bool = True):
    return None


# ---------------------------------------------------------------------------
# messengers
# ---------------------------------------------------------------------------
class Messenger:
    def __init__(self, fn: Optional[Callable] = None):
        self.fn = fn

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, exc_type, exc, tb):
        if _STACK and _STACK[-1] is self:
            _STACK.pop()
        else:  # exception unwound through nested handlers
            while _STACK and _STACK[-1] is not self:
                _STACK.pop()
            if _STACK:
                _STACK.pop()
        return False

    def __call__(self, *args, **kwargs):
        with self:
            return self.fn(*args, **kwargs)

    def _process_message(self, msg):
        pass

    def _postprocess_message(self, msg):
        pass


def _default_apply(msg):
    if msg["type"] == "sample" and msg["value"] is None:
        msg["value"] = msg["fn"]()


def apply_stack(msg):
    pointer = 0
    for pointer, handler in enumerate(reversed(_STACK)):
        handler._process_message(msg)
        if msg.get("stop"):
            break
    if not msg["done"]:
        _default_apply(msg)
        msg["done"] = True
    for handler in _STACK[-pointer - 1:] if _STACK else []:
        handler._postprocess_message(msg)
    return msg


def _new_msg(kind, name, fn, value, observed):
    return {"type": kind, "name": name, "fn": fn, "value": value, "is_observed": observed,
            "scale": 1.0, "mask": None, "done": False, "stop": False, "infer": {},
            "cond_indep_stack": ()}


def sample(name, fn, obs=None, infer=None):
    if not _STACK:
        return obs if obs is not None else fn()
    msg = _new_msg("sample", name, fn, obs, obs is not None)
    if infer:
        msg["infer"] = dict(infer)
    return apply_stack(msg)["value"]


def param(name, init_tensor=None, constraint=constraints.real, event_dim=None):
    if init_tensor is None and name not in _PARAM_STORE:
        raise KeyError(f"param '{name}' has no initial value")
    value = _PARAM_STORE.setdefault(name, init_tensor, constraint)
    if not _STACK:
        return value
    msg = _new_msg("param", name, None, value, False)
    msg["done"] = True
    return apply_stack(msg)["value"]


def factor(name, log_factor, *, has_rsample=None):
    unit = _Unit(log_factor)
    return sample(name, unit, obs=unit())


def deterministic(name, value, event_dim=None):
    return sample(name, _Delta(value), obs=value)


def markov(iterable=None, history=1, keep=False):
    return iterable


class plate(Messenger):  # noqa: N801
    """Conditional-independence annotation; the WN path needs no subsampling, so
    this only iterates / scopes."""

    def __init__(self, name, size=None, subsample_size=None, dim=None):
        super().__init__()
        self.name, self.size, self.dim = name, size, dim

    def __iter__(self):
        return iter(range(self.size))


class Trace:
    def __init__(self):
        self.nodes: "OrderedDict[str, dict]" = OrderedDict()

    def add_node(self, site_name, **site):
        if site_name in self.nodes and site.get("type") == "sample":
            raise RuntimeError(f"Multiple sample sites named '{site_name}'")
        self.nodes[site_name] = site

    def compute_log_prob(self, site_filter=lambda name, site: True):
        # Sites that share ONE distribution object (the drop-in model declares its 2N + N^2
        # global latents from three prior objects) are evaluated with one stacked log_prob
        # call per (object, value shape) instead of one tiny call per site.
        groups: Dict[tuple, list] = {}
        for name, site in self.nodes.items():
            if site["type"] == "sample" and site_filter(name, site) and "log_prob" not in site:
                fn, v = site["fn"], site["value"]
                if isinstance(fn, _td.Distribution) and fn.batch_shape == () and fn.event_shape == () \
                        and torch.is_tensor(v):
                    groups.setdefault((id(fn), tuple(v.shape), v.device, v.dtype), []).append(site)
        for sites in groups.values():
            if len(sites) < 8:
                continue
            lp_all = sites[0]["fn"].log_prob(torch.stack([s["value"] for s in sites]))
            for i, s in enumerate(sites):
                s["_lp"] = lp_all[i]
        for name, site in self.nodes.items():
            if site["type"] == "sample" and site_filter(name, site) and "log_prob" not in site:
                lp = site.pop("_lp") if "_lp" in site else site["fn"].log_prob(site["value"])
                sc = site["scale"]
                if not (isinstance(sc, (int, float)) and sc == 1.0):
                    lp = lp * sc
                site["unscaled_log_prob"] = lp
                site["log_prob"] = lp
                site["log_prob_sum"] = lp.sum()

    def log_prob_sum(self, site_filter=lambda name, site: True):
        self.compute_log_prob(site_filter)
        total = 0.0
        for name, site in self.nodes.items():
            if site["type"] == "sample" and site_filter(name, site):
                total = total + site["log_prob_sum"]
        return total

    @property
    def stochastic_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and not s["is_observed"]]

    @property
    def observation_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and s["is_observed"]]

    @property
    def param_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "param"]

    def iter_stochastic_nodes(self):
        for n in self.stochastic_nodes:
            yield n, self.nodes[n]


class _TraceMessenger(Messenger):
    def __init__(self, fn=None):
        super().__init__(fn)
        self.trace = Trace()

    def __enter__(self):
        self.trace = Trace()
        return super().__enter__()

    def _postprocess_message(self, msg):
        site = {k: msg[k] for k in ("type", "name", "fn", "value", "is_observed", "scale",
                                    "mask", "infer")}
        self.trace.add_node(msg["name"], **site)

    def get_trace(self, *args, **kwargs):
        self(*args, **kwargs)
        return self.trace


class _ConditionMessenger(Messenger):
    def __init__(self, fn=None, data=None):
        super().__init__(fn)
        self.data = {} if data is None else data

    def _process_message(self, msg):
        if msg["type"] == "sample" and not msg["is_observed"] and msg["name"] in self.data:
            msg["value"] = self.data[msg["name"]]
            msg["is_observed"] = True


class _ScaleMessenger(Messenger):
    def __init__(self, fn=None, scale=1.0):
        super().__init__(fn)
        self.scale = scale

    def _process_message(self, msg):
        if msg["type"] == "sample":
            msg["scale"] = msg["scale"] * self.scale


class _ReplayMessenger(Messenger):
    def __init__(self, fn=None, trace=None):
        super().__init__(fn)
        self.guide_trace = trace

    def _process_message(self, msg):
        if (msg["type"] == "sample" and not msg["is_observed"]
                and self.guide_trace is not None and msg["name"] in self.guide_trace.nodes):
            msg["value"] = self.guide_trace.nodes[msg["name"]]["value"]
            msg["done"] = True


class _BlockMessenger(Messenger):
    def __init__(self, fn=None, hide_fn=None, hide=None, expose=None):
        super().__init__(fn)
        if hide_fn is None:
            if hide is not None:
                hide_fn = lambda m: m["name"] in hide  # noqa: E731
            elif expose is not None:
                hide_fn = lambda m: m["name"] not in expose  # noqa: E731
            else:
                hide_fn = lambda m: True  # noqa: E731
        self.hide_fn = hide_fn

    def _process_message(self, msg):
        if self.hide_fn(msg):
            msg["stop"] = True


_PARTICLE_SHAPE = ()        # set by Trace_ELBO(vectorize_particles=True) around a step

poutine = types.SimpleNamespace(
    trace=_TraceMessenger, condition=_ConditionMessenger, scale=_ScaleMessenger,
    replay=_ReplayMessenger, block=_BlockMessenger, Trace=Trace, Messenger=Messenger,
)


from ._infer import infer, optim  # noqa: E402,F401  (pyro.infer / pyro.optim)
