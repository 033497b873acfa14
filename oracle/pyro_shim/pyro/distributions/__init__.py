"""ORACLE / TEST INFRASTRUCTURE ONLY.  See ``pyro_shim/pyro/__init__.py``.

The four distributions the reference's hot path uses, as thin subclasses of
the installed ``torch.distributions`` classes.  ``__call__`` mirrors Pyro's
``TorchDistributionMixin.__call__`` (``rsample`` when available).  ``draw_with``
re-implements torch's own reparameterisation arithmetic with externally
supplied standard noise (Normal: ``loc + eps * scale``; Exponential:
``e / rate``) -- see ``torch/distributions/normal.py::rsample`` and
``exponential.py::rsample`` of the installed torch.
"""
import torch
import torch.distributions as _td
from torch.distributions import constraints  # noqa: F401  (re-export)


class _CallMixin:
    def __call__(self, sample_shape=torch.Size()):
        if self.has_rsample:
            return self.rsample(sample_shape)
        return self.sample(sample_shape)


class Gamma(_CallMixin, _td.Gamma):
    pass


class Normal(_CallMixin, _td.Normal):
    def draw_with(self, source, name):
        eps = source(name, "normal")
        if eps is None:
            return self()
        eps = torch.as_tensor(eps, dtype=self.loc.dtype).reshape(self._extended_shape())
        return self.loc + eps * self.scale


class Exponential(_CallMixin, _td.Exponential):
    def draw_with(self, source, name):
        e = source(name, "exponential")
        if e is None:
            return self()
        e = torch.as_tensor(e, dtype=self.rate.dtype).reshape(self._extended_shape())
        return e / self.rate


class RelaxedBernoulliStraightThrough(_CallMixin, _td.RelaxedBernoulli):
    """Pyro's straight-through relaxed Bernoulli: hard 0/1 forward value whose
    gradient is that of the soft sample; ``log_prob`` scores the soft value
    when the argument carries one (``_unquantize``), else the argument itself
    (which is what happens for observed 0/1 values, bayes_air/network.py:102-109).
    """

    def rsample(self, sample_shape=torch.Size()):
        soft = super().rsample(sample_shape)
        soft = torch.distributions.utils.clamp_probs(soft)
        hard = soft.round()
        out = (hard - soft).detach() + soft
        out._unquantize = soft
        return out

    def log_prob(self, value):
        value = getattr(value, "_unquantize", value)
        return super().log_prob(value)

    def draw_with(self, source, name):
        """torch's LogitRelaxedBernoulli.rsample + SigmoidTransform with an externally
        supplied uniform (relaxed_bernoulli.py::rsample, transforms.py::SigmoidTransform)."""
        u = source(name, "uniform")
        if u is None:
            return self()
        base = self.base_dist
        shape = base._extended_shape()
        probs = torch.distributions.utils.clamp_probs(base.probs.expand(shape))
        uniforms = torch.distributions.utils.clamp_probs(
            torch.as_tensor(u, dtype=probs.dtype).reshape(shape))
        x = (uniforms.log() - (-uniforms).log1p() + probs.log() - (-probs).log1p()) / base.temperature
        finfo = torch.finfo(x.dtype)
        soft = torch.clamp(torch.sigmoid(x), min=finfo.tiny, max=1.0 - finfo.eps)
        soft = torch.distributions.utils.clamp_probs(soft)
        hard = soft.round()
        out = (hard - soft).detach() + soft
        out._unquantize = soft
        return out
