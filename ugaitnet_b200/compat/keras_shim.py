"""Minimal stand-ins for the Keras objects the reference's builder signatures mention
(optimizers.SGD/Adam instances, the Maximum / Average merge classes, sign_max, Model, History)."""
from __future__ import annotations

import numpy as np

from ..config import MERGE_AVG, MERGE_MAX, MERGE_SIGNMAX


class _Optimizer:
    def __init__(self, name, lr, **kw):
        self.name, self.lr, self.kw = name, float(lr), kw

    learning_rate = property(lambda self: self.lr)


class optimizers:  # noqa: N801  (mirrors `from tensorflow.keras import optimizers`)
    @staticmethod
    def SGD(learning_rate=0.01, momentum=0.0, decay=0.0, lr=None, **kw):
        return _Optimizer("sgd", lr if lr is not None else learning_rate, momentum=momentum, decay=decay)

    @staticmethod
    def Adam(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, lr=None, **kw):
        return _Optimizer("amsgrad" if amsgrad else "adam", lr if lr is not None else learning_rate, beta1=beta_1,
                          beta2=beta_2, eps=epsilon)

    @staticmethod
    def AdamW(weight_decay=0.0, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, lr=None, **kw):
        """tfa.optimizers.AdamW(learning_rate=lr, weight_decay=wd) (mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:236)."""
        return _Optimizer("adamw", lr if lr is not None else learning_rate, beta1=beta_1, beta2=beta_2, eps=epsilon,
                          weight_decay=weight_decay)


class _Merge:
    """`fMerge(name="fusion")([a, b, c])` in the reference (nets/mj_uwyhNets_ba.py:1189); here the class /
    factory is only a tag that selects the merge mode of the fused gate+merge+l2norm kernel."""
    merge_id = MERGE_MAX

    def __init__(self, name=None, **kw):
        self.name = name


class Maximum(_Merge):
    merge_id = MERGE_MAX


class Average(_Merge):
    merge_id = MERGE_AVG


def sign_max(**kwargs):
    """mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:169-178: per element keep the modality value with the
    largest |x| (ties -> lowest modality index)."""
    m = _Merge(**kwargs)
    m.merge_id = MERGE_SIGNMAX
    return m


sign_max.merge_id = MERGE_SIGNMAX


def merge_id_of(fMerge) -> int:
    mid = getattr(fMerge, "merge_id", None)
    if mid is None:
        name = getattr(fMerge, "__name__", str(fMerge))
        mid = {"Maximum": MERGE_MAX, "Average": MERGE_AVG, "sign_max": MERGE_SIGNMAX}.get(name)
    if mid is None:
        raise NotImplementedError(f"fMerge={fMerge!r}: only Maximum, Average and sign_max are implemented")
    return mid


class History:
    def __init__(self):
        self.epoch, self.history = [], {}

    def add(self, epoch, logs):
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(float(v))


class _Tag:
    """Symbolic handle returned by model.input / layer.output."""

    def __init__(self, model, name):
        self.model, self.name = model, name


def Model(inputs=None, outputs=None, **kw):
    """`Model(model.input, model.get_layer(codename).output)` of the test scripts
    (mains/mj_testUWYHGaitNet_open_tum.py:139-148): returns a predictor of that layer."""
    tags = outputs if isinstance(outputs, (list, tuple)) else [outputs]
    if not all(isinstance(t, _Tag) for t in tags):
        raise TypeError("compat.Model only builds sub-models from model.get_layer(name).output handles")
    return tags[0].model.submodel([t.name for t in tags], as_list=isinstance(outputs, (list, tuple)))
