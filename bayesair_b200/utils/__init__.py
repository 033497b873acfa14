"""Same module path as the reference's ``bayes_air.utils`` for the one utility on the
widened path: ``bayesair_b200.utils.dataloader``."""
