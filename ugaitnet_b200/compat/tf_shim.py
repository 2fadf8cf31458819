"""Import shim that lets the reference's UNMODIFIED ``mains/`` scripts import against this package.

Every reference main starts with ``import tensorflow as tf`` followed by session / seed boilerplate, ``deepdish``,
``tensorflow_addons`` and ``from nets.mj_uwyhNets_ba import ...`` (/root/reference/mains/mj_trainUWYHGaitNet_DataGen_3mods.py:1-60,
mj_testUWYHGaitNet_open_tum.py:1-48).  ``install()`` registers

  * ``tensorflow`` (+ ``tensorflow.keras.*``, ``tensorflow.compat.v1``), ``tensorflow_addons``: permissive stand-ins --
    the names the mains touch at import time and in their training / test loops exist (optimizers map to
    ugaitnet_b200.compat.keras_shim, ``Model`` to its sub-model factory, ``Maximum`` / ``Average`` to the fusion tags,
    ``callbacks.Callback`` / ``ReduceLROnPlateau`` / ``ModelCheckpoint`` to small working classes, ``utils.Sequence`` to a
    plain base class); anything else resolves to an inert stub, never to arithmetic;
  * ``deepdish`` with ``io.load`` / ``io.save`` on ugaitnet_b200.samples (the reference's ``.h5`` sample / config files);
  * the package ``nets`` = ugaitnet_b200.compat.nets (mj_uwyhNets_ba, triplet_loss_all, mj_loss, mj_metrics), with the
    reference's own ``nets`` directory as a fallback for the modules outside the hot path (mj_utils, aux_loss).

Nothing here computes anything: the builders, losses and the k-NN the mains then call are the B200 ones.
"""
from __future__ import annotations

import importlib
import os
import sys
import types


class _Stub:
    """Inert object: any attribute, call, item or context-manager use yields another stub."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        s = _Stub()
        object.__setattr__(self, name, s)
        return s

    def __call__(self, *a, **k):
        return _Stub()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return False


class _StubModule(types.ModuleType):
    """Module whose unknown CamelCase attributes are fresh stub CLASSES (usable as base classes) and whose other
    unknown attributes are inert stubs."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        v = type(name, (_Stub,), {}) if name[:1].isupper() else _Stub()
        setattr(self, name, v)
        return v


class Sequence:                     # tensorflow.keras.utils.Sequence: the generators only need the base class
    def on_epoch_end(self):
        pass


class Callback:
    def __init__(self, *a, **k):
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        pass


class ReduceLROnPlateau(Callback):
    """tensorflow.keras.callbacks.ReduceLROnPlateau(monitor, factor, patience, min_lr) (mains/..._3mods.py:553-556):
    multiplies model.optimizer.lr by `factor` after `patience` epochs without improvement of `monitor` (min mode)."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4, cooldown=0,
                 min_lr=0, **kw):
        super().__init__()
        self.monitor, self.factor, self.patience, self.min_lr, self.min_delta = monitor, factor, patience, min_lr, min_delta
        self.maximize = mode == "max" or (mode == "auto" and "acc" in monitor)
        self.best, self.wait = None, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        better = self.best is None or (cur > self.best + self.min_delta if self.maximize else cur < self.best - self.min_delta)
        if better:
            self.best, self.wait = cur, 0
            return
        self.wait += 1
        if self.wait >= self.patience:
            new = max(float(self.model.optimizer.lr) * self.factor, self.min_lr)
            self.model.optimizer.lr = new
            self.wait = 0


class ModelCheckpoint(Callback):
    """tensorflow.keras.callbacks.ModelCheckpoint(filepath, save_weights_only, period) (mains/..._3mods.py:563-570)."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False, mode="auto",
                 save_freq="epoch", period=1, **kw):
        super().__init__()
        self.filepath, self.weights_only, self.period = filepath, save_weights_only, max(int(period), 1)

    def on_epoch_end(self, epoch, logs=None):
        if (epoch + 1) % self.period:
            return
        path = self.filepath.format(epoch=epoch + 1, **(logs or {}))
        (self.model.save_weights if self.weights_only else self.model.save)(path)


def _mod(name, cls=_StubModule, **attrs):
    m = cls(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def install(reference_root=None, gpu_knn=False):
    """Register the stand-in modules (idempotent).  reference_root: the checkout whose mains/ will be imported; its
    directory goes on sys.path (as the mains themselves do) and its nets/ becomes the fallback of the `nets` package.
    gpu_knn: `from sklearn.neighbors import KNeighborsClassifier` -- what the open-world test mains do INSIDE their
    evaluation functions (mains/mj_testUWYHGaitNet_open_tum.py:331, mj_testUWYHGaitNet_open_casiab.py) -- then yields
    ugaitnet_b200.knn.KNeighborsClassifier (same fit / predict / kneighbors, B200 search); install(gpu_knn=False)
    puts scikit-learn's class back."""
    from ugaitnet_b200.compat import keras_shim as ks
    here = os.path.dirname(os.path.abspath(__file__))
    import sklearn.neighbors as skn
    if not hasattr(skn, "_ugn_sklearn_knn"):
        skn._ugn_sklearn_knn = skn.KNeighborsClassifier
    if gpu_knn:
        from ugaitnet_b200.knn import KNeighborsClassifier as GpuKNN
        skn.KNeighborsClassifier = GpuKNN
    else:
        skn.KNeighborsClassifier = skn._ugn_sklearn_knn
    if "tensorflow" not in sys.modules or not isinstance(sys.modules["tensorflow"], _StubModule):
        tf = _mod("tensorflow", __version__="2.3.0-ugaitnet_b200-shim")
        tf.executing_eagerly = lambda: True
        _mod("tensorflow.random", set_seed=lambda s: None)
        compat = _mod("tensorflow.compat")
        _mod("tensorflow.compat.v1")
        keras = _mod("tensorflow.keras")
        _mod("tensorflow.keras.optimizers", SGD=ks.optimizers.SGD, Adam=ks.optimizers.Adam)
        _mod("tensorflow.keras.layers", Maximum=ks.Maximum, Average=ks.Average)
        _mod("tensorflow.keras.models", Model=ks.Model)
        keras.Model = ks.Model
        _mod("tensorflow.keras.callbacks", Callback=Callback, ReduceLROnPlateau=ReduceLROnPlateau,
             ModelCheckpoint=ModelCheckpoint)
        _mod("tensorflow.keras.utils", Sequence=Sequence)
        for sub in ("backend", "regularizers", "losses", "initializers", "metrics", "preprocessing"):
            _mod("tensorflow.keras." + sub)
        _mod("tensorflow.keras.preprocessing.image")
        tfa = _mod("tensorflow_addons")
        _mod("tensorflow_addons.losses")
        _mod("tensorflow_addons.optimizers", AdamW=ks.optimizers.AdamW)
    if "deepdish" not in sys.modules:
        from ugaitnet_b200 import samples
        dd = _mod("deepdish")
        _mod("deepdish.io", load=samples.load_sample, save=samples.save_sample)
    for name in ("tensorboard.plugins.hparams.api", "tensorboard.plugins.projector"):
        try:
            importlib.import_module(name)
        except Exception:
            parts = name.split(".")
            for i in range(1, len(parts) + 1):
                sub = ".".join(parts[:i])
                if sub not in sys.modules:
                    _mod(sub)
    # the `nets` package of the mains = the B200 drop-in (+ the reference's directory for modules outside the hot path)
    compat_dir = here
    if compat_dir not in sys.path:
        sys.path.insert(0, compat_dir)
    for k in [k for k in sys.modules if k == "nets" or k.startswith("nets.")]:
        del sys.modules[k]
    nets = importlib.import_module("nets")
    tfa_losses = sys.modules.get("tensorflow_addons.losses")
    if tfa_losses is not None and not hasattr(tfa_losses, "TripletHardLoss"):
        tfa_losses.TripletHardLoss = importlib.import_module("nets.mj_uwyhNets_ba").TripletHardLoss
    if reference_root:
        ref_nets = os.path.join(reference_root, "nets")
        if os.path.isdir(ref_nets) and ref_nets not in nets.__path__:
            nets.__path__.append(ref_nets)
        if reference_root not in sys.path:
            sys.path.append(reference_root)
    return nets
