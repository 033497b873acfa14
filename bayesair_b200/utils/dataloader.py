"""Alias of :mod:`bayesair_b200.dataloader` under the reference's module path
(``bayes_air/utils/dataloader.py``)."""
from ..dataloader import *  # noqa: F401,F403
from ..dataloader import __all__  # noqa: F401
