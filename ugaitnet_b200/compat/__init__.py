"""Keras-shaped surface of the reference for the hot path.

Putting ``ugaitnet_b200/compat`` at the front of ``sys.path`` makes ``import nets.mj_uwyhNets_ba``,
``nets.triplet_loss_all``, ``nets.mj_loss`` and ``nets.mj_metrics`` resolve to the B200 implementations
with the reference's names, argument order and return types."""
from .keras_shim import Average, History, Maximum, Model, optimizers, sign_max  # noqa: F401
