"""Probabilistic-programming surface the host side is written against.

The reference model is a Pyro program (``import pyro`` at
``bayes_air/model.py:4``), and its callers drive it through Pyro effect handlers
(``poutine.trace`` / ``condition`` / ``scale``, ``scripts/wn/train.py:306-314``,
``scripts/wn/wn_network.py:285-302``).  ``pyro-ppl`` is not installed in this
image and cannot be (no network), so:

* if a real ``pyro`` is importable it is used unchanged -- the model in
  ``bayesair_b200.model`` then runs under genuine Pyro handlers, guides and SVI;
* otherwise ``bayesair_b200.ppl._mini`` provides the same names with the same
  semantics for the subset the WN path touches.

Either way ``from bayesair_b200.ppl import pyro, dist, poutine`` works.
"""
from __future__ import annotations

import importlib
import importlib.util


def _real_pyro():
    try:
        if importlib.util.find_spec("pyro") is None:
            return None
        mod = importlib.import_module("pyro")
    except Exception:
        return None
    # the oracle's shim (tests only) must never be picked up by the product
    if getattr(mod, "__bayesair_oracle_shim__", False):
        return None
    return mod


_pyro = _real_pyro()
if _pyro is not None:
    pyro = _pyro
    dist = importlib.import_module("pyro.distributions")
    poutine = importlib.import_module("pyro.poutine")
    USING_REAL_PYRO = True
else:
    from . import _mini as pyro  # noqa: F401
    from ._mini import distributions as dist  # noqa: F401
    from ._mini import poutine  # noqa: F401
    USING_REAL_PYRO = False

__all__ = ["pyro", "dist", "poutine", "USING_REAL_PYRO"]
