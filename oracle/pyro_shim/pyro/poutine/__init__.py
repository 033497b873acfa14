"""ORACLE / TEST INFRASTRUCTURE ONLY.  See ``pyro_shim/pyro/__init__.py``.

``trace`` / ``condition`` / ``scale`` effect handlers and a ``Trace`` with the
``log_prob_sum`` semantics of pyro-ppl 1.9.0 (per-site ``log_prob(value)``
multiplied by the site scale, ``.sum()``-ed, accumulated in site order starting
from the Python float 0.0) -- the call the reference makes at
scripts/wn/train.py:314.
"""
import collections

import pyro as _pyro


class Trace:
    def __init__(self):
        self.nodes = collections.OrderedDict()

    def add_node(self, site_name, **kw):
        if site_name in self.nodes and kw.get("type") == "sample":
            raise RuntimeError(f"Multiple sample sites named '{site_name}'")
        self.nodes[site_name] = kw

    def log_prob_sum(self, site_filter=lambda name, site: True):
        total = 0.0
        for name, site in self.nodes.items():
            if site["type"] != "sample" or not site_filter(name, site):
                continue
            lp = site["fn"].log_prob(site["value"])
            scale = site.get("scale", 1.0)
            if not (isinstance(scale, float) and scale == 1.0):
                lp = lp * scale
            total = total + lp.sum()
        return total

    def stochastic_nodes(self):
        return [n for n, s in self.nodes.items()
                if s["type"] == "sample" and not s["is_observed"]]

    def observation_nodes(self):
        return [n for n, s in self.nodes.items()
                if s["type"] == "sample" and s["is_observed"]]


class trace(_pyro._Handler):  # noqa: N801
    def __init__(self, fn=None):
        super().__init__(fn)
        self.trace = Trace()

    def postprocess(self, msg):
        self.trace.add_node(
            msg["name"], type=msg["type"], name=msg["name"], fn=msg["fn"],
            value=msg["value"], is_observed=msg["is_observed"], scale=msg["scale"],
        )

    def get_trace(self, *args, **kwargs):
        self.trace = Trace()
        self(*args, **kwargs)
        return self.trace


class condition(_pyro._Handler):  # noqa: N801
    def __init__(self, fn=None, data=None):
        super().__init__(fn)
        self.data = data or {}

    def process(self, msg):
        if msg["type"] == "sample" and not msg["is_observed"] and msg["name"] in self.data:
            msg["value"] = self.data[msg["name"]]
            msg["is_observed"] = True


class scale(_pyro._Handler):  # noqa: N801
    def __init__(self, fn=None, scale=1.0):
        super().__init__(fn)
        self.scale = scale

    def process(self, msg):
        if msg["type"] == "sample":
            msg["scale"] = msg["scale"] * self.scale
