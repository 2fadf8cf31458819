"""Drop-in for the model builders of nets/mj_uwyhNets_ba.py (batch-all version).

Same class / static-method names, positional order, defaults and return types as the reference
(UWYHSemiNet3Mods.build :1032-1037, build_or_load :1309-1316, UWYHSemiNet.build :669-672,
fit_generator :938-968, encode :971-999); the returned object follows the slice of the Keras Model
protocol the reference's mains/ scripts use (fit / predict / get_layer / save_weights / ...), and every
step runs on the B200 engine (ugaitnet_b200.net.UGaitEngine).

``gaitset=True`` (with a non-'relu' fActivation, as the reference requires, :1101) builds the GaitSet
branch type (build_gaitset_branch :420-484) on ugaitnet_b200.gaitset.GaitSetEngine: inputs
[B,25,60,60,c], signature [62,B,256], descriptor layer "flatten" (typecode 3).

``smoothlabels`` (label-smoothed cross-entropy, :1252-1262), ``normbfmerge`` (per-branch l2_normalize before
the gate, :1167-1168) and ``aux_losses`` (classprob_{of,gray,depth} heads on the gated branch outputs,
:1222-1251; stacked-CNN branches) are implemented.  Builder arguments that select graphs outside the hot path
raise NotImplementedError: use3D, postriplet == 2, init_branches / initnet weight surgery from Keras .hdf5
files, tfa TripletHardLoss (compile_hard), aux_losses together with gaitset.
"""
from __future__ import annotations

import os
import os.path as osp
from typing import Dict, List, Optional

import numpy as np
import torch

from ugaitnet_b200 import ops
from ugaitnet_b200.config import ACT_LEAKY, ACT_RELU, BRANCH_NAMES, GS_CONVS, MERGE_MAX, GaitSetConfig, NetConfig
from ugaitnet_b200.gaitset import GaitSetEngine
from ugaitnet_b200.net import UGaitEngine
from ugaitnet_b200.compat.keras_shim import Average, History, Maximum, _Tag, merge_id_of, optimizers  # noqa: F401
from ugaitnet_b200.compat.nets.triplet_loss_all import triplet_loss

MATH_MODE = os.environ.get("UGN_MATH_MODE", "bf16x3")


def mj_tensor_times_scalar(d):
    return d[0] * d[1]


def _unsupported(**flags):
    bad = [k for k, v in flags.items() if v]
    if bad:
        raise NotImplementedError(f"builder options outside the B200 hot path: {', '.join(bad)} "
                                  "(see DESIGN.md, 'out of scope')")


class _LayerProxy:
    def __init__(self, model, name, units=None, sublayers=()):
        self.model, self.name, self.units, self.trainable = model, name, units, True
        self.layers = list(sublayers)

    @property
    def output(self):
        return _Tag(self.model, self.name)

    def get_weights(self):
        return self.model._layer_weights(self.name)


class _OptimizerView:
    def __init__(self, model):
        self._m = model

    @property
    def lr(self):
        return self._m.engine.lr

    @lr.setter
    def lr(self, v):
        self._m.engine.lr = float(v)

    learning_rate = lr


class _SubModel:
    def __init__(self, model, names, as_list):
        self.model, self.names, self.as_list = model, names, as_list

    def predict(self, x, batch_size=None, verbose=0):
        outs = [self.model._predict_layer(x, n) for n in self.names]
        return outs if self.as_list else outs[0]

    __call__ = predict


class UGaitModel:
    """The compiled-model object the builders return."""
    dtype = "float32"

    def __init__(self, cfg: NetConfig, optimizer, losses, loss_weights, multimodal: bool):
        self.cfg, self.multimodal = cfg, multimodal
        opt = optimizer if optimizer is not None else optimizers.SGD(0.001, 0.9)
        kw = dict(getattr(opt, "kw", {}))
        self.gaitset = isinstance(cfg, GaitSetConfig)
        engine_cls = GaitSetEngine if self.gaitset else UGaitEngine
        self.engine = engine_cls(cfg, math_mode=MATH_MODE, optimizer=getattr(opt, "name", "sgd"),
                                  lr=getattr(opt, "lr", 0.001), momentum=kw.get("momentum", 0.9),
                                  beta1=kw.get("beta1", 0.9), beta2=kw.get("beta2", 0.999), eps=kw.get("eps", 1e-7),
                                  lr_decay=kw.get("decay", 0.0), decoupled_weight_decay=kw.get("weight_decay", 0.0),
                                  use_graph=os.environ.get("UGN_GRAPH", "1") == "1")
        self.loss, self.loss_weights = losses, loss_weights
        self.optimizer = _OptimizerView(self)
        self.stop_training = False
        names = []
        for m in range(cfg.nmods):
            if self.gaitset:
                sub = [_LayerProxy(self, f"{BRANCH_NAMES[m]}/{n[0]}") for n in GS_CONVS] + \
                      [_LayerProxy(self, f"{BRANCH_NAMES[m]}/matmul")]
            else:
                sub = [_LayerProxy(self, f"{BRANCH_NAMES[m]}/{n}") for n in
                       [f"conv{i}" for i in range(len(cfg.filters_numbers))] + ["ofFlat", "dense", "drop", "ofCode"]]
            names.append(_LayerProxy(self, BRANCH_NAMES[m], units=cfg.nd, sublayers=sub))
        for n in ("gate_of1", "gate_gray1", "gate_depth1")[:cfg.nmods]:
            names.append(_LayerProxy(self, n))
        names += [_LayerProxy(self, "fusion"), _LayerProxy(self, "signature", units=cfg.nd)]
        if cfg.nc > 0:
            names += [_LayerProxy(self, "code", units=cfg.nc), _LayerProxy(self, "dropcode")]
        if cfg.nclasses > 0:
            if self.gaitset:
                names.append(_LayerProxy(self, "flatten", units=62 * (cfg.nc or cfg.nd)))     # typecode 3 (:1213)
            names.append(_LayerProxy(self, "classprob", units=cfg.nclasses))
            if getattr(cfg, "aux_losses", False):
                names += [_LayerProxy(self, n, units=cfg.nclasses) for n in
                          ("classprob_of", "classprob_gray", "classprob_depth")[:cfg.nmods]]
        self.layers = names
        self.input = [_Tag(self, n) for n in ("ofinput1", "ofuse1", "grayinput1", "grayuse1", "depthinput1",
                                              "depthuse1")[:2 * cfg.nmods]] if multimodal else _Tag(self, "ofinput1")

    # -- Keras protocol -------------------------------------------------------------------------
    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}")

    def summary(self):
        e = self.engine
        print(f"UGaitNet (B200 engine, math={e.math_mode}): {self.cfg.nmods} branch(es), nd={self.cfg.nd}, "
              f"nc={self.cfg.nc}, classes={self.cfg.nclasses}")
        for s in e.seg_list:
            print(f"  {s.name:28s} {str(s.shape):24s} {s.n:>10d}")
        print(f"  total params: {sum(s.n for s in e.seg_list):,}")

    def submodel(self, names, as_list):
        return _SubModel(self, names, as_list)

    def _split_x(self, x):
        def cu(a):
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
            return t.to(self.engine.dev, dtype=torch.float32, non_blocking=True)
        if not self.multimodal:
            return [cu(x[0] if isinstance(x, (list, tuple)) else x)], None
        return [cu(x[2 * m]) for m in range(self.cfg.nmods)], [cu(x[2 * m + 1]) for m in range(self.cfg.nmods)]

    def _labels(self, y):
        lab = y[0] if isinstance(y, (list, tuple)) else y
        lab = lab if torch.is_tensor(lab) else torch.from_numpy(np.asarray(lab))
        return lab.reshape(-1).to(self.engine.dev)

    def _predict_layer(self, x, name, batch_size=None):
        ins, fl = self._split_x(x)
        layer = {"signature": "signature", "code": "code", "classprob": "classprob"}.get(name)
        if layer is None and self.gaitset and name == "flatten":
            layer = "flatten"
        if layer is None:
            raise NotImplementedError(f"descriptor layer {name!r} (typecode 3 'flatten' is the GaitSet layout)")
        out = self.engine.predict(ins, fl, layer=layer)
        if name == "classprob":
            out = torch.softmax(out, dim=1)     # Dense(softmax) output (:1214)
        return out.cpu().numpy()

    def predict(self, x, batch_size=None, verbose=0):
        sig = self._predict_layer(x, "signature")
        if self.cfg.nclasses > 0:
            return [sig, self._predict_layer(x, "classprob")]
        return sig

    def _logs(self, out, prefix=""):
        cfg = self.cfg
        logs = {}
        trip = float(out["triplet"])
        total = cfg.wver * trip
        if cfg.nclasses > 0:
            ce = float(out["ce"])
            logs[prefix + "signature_loss"], logs[prefix + "classprob_loss"] = trip, ce
            logs[prefix + "classprob_acc"] = float(out["acc"])
            total += cfg.wid * ce
            for m, v in enumerate(out.get("aux_ce", [])):           # classprob_{of,gray,depth} heads (:1222-1251)
                name = ("classprob_of", "classprob_gray", "classprob_depth")[m]
                logs[prefix + name + "_loss"], logs[prefix + name + "_acc"] = float(v), float(out["aux_acc"][m])
                total += cfg.waux * float(v)
        if "reg" in out:
            total += float(out["reg"])
        logs[prefix + "loss"] = total
        return logs

    def train_on_batch(self, x, y, **kw):
        ins, fl = self._split_x(x)
        return self._logs(self.engine.train_step(ins, fl, self._labels(y)))

    def test_on_batch(self, x, y, **kw):
        ins, fl = self._split_x(x)
        return self._logs(self.engine.eval_losses(ins, fl, self._labels(y)))

    def fit(self, x=None, validation_data=None, epochs=1, steps_per_epoch=None, callbacks=None,
            validation_steps=None, initial_epoch=0, verbose=2, **kw):
        gen, hist = x, History()
        callbacks = callbacks or []
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                try:
                    cb.set_model(self)
                except Exception:
                    pass
        for epoch in range(initial_epoch, epochs):
            n = steps_per_epoch or len(gen)
            acc: Dict[str, float] = {}
            for i in range(n):
                X, y = gen[i % len(gen)][:2]
                for k, v in self.train_on_batch(X, y).items():
                    acc[k] = acc.get(k, 0.0) + v
            logs = {k: v / n for k, v in acc.items()}
            if validation_data is not None:
                nv = validation_steps or len(validation_data)
                vacc: Dict[str, float] = {}
                for i in range(nv):
                    X, y = validation_data[i % len(validation_data)][:2]
                    for k, v in self._logs(self.engine.eval_losses(*self._split_x(X), self._labels(y)), "val_").items():
                        vacc[k] = vacc.get(k, 0.0) + v
                logs.update({k: v / nv for k, v in vacc.items()})
            logs["lr"] = self.engine.lr
            hist.add(epoch, logs)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()), flush=True)
            if hasattr(gen, "on_epoch_end"):
                gen.on_epoch_end()
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    try:
                        cb.on_epoch_end(epoch, logs)
                    except Exception:
                        pass
            if self.stop_training:
                break
        return hist

    # -- weights: Keras names / layouts ((kh,kw,cin,cout) conv kernels, (in,out) dense kernels) -------
    @staticmethod
    def _keras_name(seg_name):
        base, kind = seg_name.rsplit("/", 1)
        return f"{base}/{'kernel' if kind == 'w' else 'bias'}"

    def _layer_weights(self, name):
        P = self.engine.export_params()
        out = []
        for k in sorted(P):
            if k.startswith(name + "/") or k.rsplit("/", 1)[0] == name:
                out.append(self._to_keras(k, P[k]).cpu().numpy())
        return out

    @staticmethod
    def _to_keras(k, v):
        if v.dim() == 4:
            return v.permute(2, 3, 1, 0).contiguous()      # [Cout,Cin,kh,kw] -> (kh,kw,cin,cout)
        if v.dim() == 2:
            return v.t().contiguous()                      # [out,in] -> (in,out)
        return v

    @staticmethod
    def _from_keras(v):
        v = torch.as_tensor(v)
        if v.dim() == 4:
            return v.permute(3, 2, 0, 1).contiguous()
        if v.dim() == 2:
            return v.t().contiguous()
        return v

    def save_weights(self, path, **kw):
        P = self.engine.export_params()
        arrs = {self._keras_name(k): self._to_keras(k, v).cpu().numpy() for k, v in P.items()}
        with open(path, "wb") as f:          # .hdf5 in the reference; h5py is unavailable -> npz container
            np.savez(f, **arrs)

    def save(self, path, **kw):
        self.save_weights(path)
        e = self.engine
        with open(path + ".opt", "wb") as f:
            np.savez(f, m=e.m.cpu().numpy(), v=e.v.cpu().numpy(), t=e.t, lr=e.lr)

    def load_weights(self, path, by_name=True, skip_mismatch=False, **kw):
        z = np.load(path)
        mine = {self._keras_name(k): k for k in self.engine.segs}
        upd = {}
        for kn in z.files:
            k = mine.get(kn)
            if k is None:
                continue
            v = self._from_keras(z[kn])
            tshape = self.engine.oracle_shape(k)
            if tuple(v.shape) != tuple(tshape):
                if skip_mismatch:
                    continue
                raise ValueError(f"shape mismatch for {kn}: {tuple(v.shape)} vs {tuple(tshape)}")
            upd[k] = v
        self.engine.load_params(upd)
        if osp.exists(path + ".opt"):
            o = np.load(path + ".opt")
            if o["m"].shape[0] == self.engine.m.shape[0]:
                self.engine.m.copy_(torch.from_numpy(o["m"])); self.engine.v.copy_(torch.from_numpy(o["v"]))
                self.engine.t = int(o["t"])


def _cfg_from_args(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units,
                   weight_decay, dropout, margin, nclasses, loss_weights, fMerge, fActivation, alpha, single,
                   smoothlabels=0, normbfmerge=False, aux_losses=False):
    fs = [k[0] if isinstance(k, (tuple, list)) else int(k) for k in filters_size][:number_convolutional_layers]
    fn = list(filters_numbers if filters_numbers is not None else [64, 128, 512, 512])[:number_convolutional_layers]
    if isinstance(ndense_units, (list, tuple)):
        nd, nc = ndense_units[0], (ndense_units[1] if len(ndense_units) > 1 else 0)
    else:
        nd, nc = ndense_units, 0
    if isinstance(dropout, (list, tuple)):
        dropout = dropout[0]
    shapes = [input_shapes] if single else list(input_shapes)
    lw = list(loss_weights) if isinstance(loss_weights, (list, tuple)) else [loss_weights, loss_weights]
    return NetConfig(in_channels=tuple(int(s[0]) for s in shapes), filters_numbers=tuple(fn), filters_size=tuple(fs),
                     nd=int(nd), nc=int(nc) if not single else 0, nclasses=int(nclasses), weight_decay=float(weight_decay),
                     merge=merge_id_of(fMerge) if not single else MERGE_MAX,
                     act=ACT_RELU if fActivation == "relu" else ACT_LEAKY, alpha=float(alpha), margin=float(margin),
                     wver=float(lw[0]) if nclasses > 0 else 1.0, wid=float(lw[1]) if nclasses > 0 and len(lw) > 1 else 0.0,
                     hw=int(shapes[0][1]), dropout=float(dropout) if dropout > 0.001 else 0.0, single=single,
                     label_smoothing=float(smoothlabels), normbfmerge=bool(normbfmerge),
                     aux_losses=bool(aux_losses) and nclasses > 0 and not single,
                     waux=float(lw[-1]))         # loss_weights padded with its last entry (:1264-1268)


def _gs_cfg_from_args(input_shapes, ndense_units, dropout, margin, nclasses, loss_weights, fMerge, fActivation, alpha,
                      smoothlabels=0):
    """gaitset=True: input_shapes [(25,60,60,2), (25,60,60,1), ...] (mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:212-213)."""
    if fActivation == 'relu':
        raise ValueError("gaitset=True needs a non-'relu' fActivation: the reference only builds the GaitSet "
                         "branches in its LeakyReLU path (nets/mj_uwyhNets_ba.py:1101-1113)")
    shapes = list(input_shapes)
    if any(len(s) != 4 for s in shapes):
        raise ValueError("gaitset=True expects input shapes (frames, H, W, channels)")
    nc = ndense_units[1] if isinstance(ndense_units, (list, tuple)) and len(ndense_units) > 1 else 0
    if isinstance(dropout, (list, tuple)):
        dropout = dropout[-1]
    lw = list(loss_weights) if isinstance(loss_weights, (list, tuple)) else [loss_weights, loss_weights]
    return GaitSetConfig(in_channels=tuple(int(s[3]) for s in shapes), frames=int(shapes[0][0]), hw=int(shapes[0][1]),
                         nc=int(nc), nclasses=int(nclasses), merge=merge_id_of(fMerge), alpha=float(alpha),
                         margin=float(margin), wver=float(lw[0]) if nclasses > 0 else 1.0,
                         wid=float(lw[1]) if nclasses > 0 and len(lw) > 1 else 0.0,
                         dropout=float(dropout) if (dropout > 0.001 and nc) else 0.0, label_smoothing=float(smoothlabels))


class UWYHNet:
    @staticmethod
    def buildBranch(name, input_shape=(50, 60, 60), number_convolutional_layers=4, filters_size=None,
                    filters_numbers=None, ndense_units=512, weight_decay=1e-4, dropout=0.4, init_branch=None):
        raise NotImplementedError("stand-alone Keras Sequential branches are not exposed; branches are built "
                                  "inside UWYHSemiNet{,3Mods}.build")

    buildBranchLReLU = buildBranch


class UWYHSemiNet:
    def __init__(self):
        self.model = None

    @staticmethod
    def get_weights_filename(netpath):
        base, ext = osp.splitext(netpath)
        return base + "_weights" + ext

    @staticmethod
    def get_netconfig_filename(netpath):
        return osp.join(osp.dirname(netpath), "model-config.hdf5")

    @staticmethod
    def build(input_shapes, number_convolutional_layers, filters_size, filters_numbers,
              ndense_units=512, weight_decay=1e-4, dropout=0.4, optimizer=None, margin=0.2,
              nclasses=0, loss_weights=[1.0, 1.0], use3D=False, smoothlabels=0, postriplet=1, init_branches=None,
              freeze_branches=False, aux_losses=False, fMerge=Maximum, fActivation='relu', alpha=0.3, gaitset=False):
        _unsupported(use3D=use3D, postriplet_2=(postriplet == 2), freeze_branches=freeze_branches,
                     aux_losses_with_gaitset=(aux_losses and gaitset),
                     init_branches=bool(init_branches) and any(init_branches.values()))
        single = not isinstance(input_shapes, list)
        if gaitset:
            _unsupported(gaitset_single_modality=single)
            cfg = _gs_cfg_from_args(input_shapes, ndense_units, dropout, margin, nclasses, loss_weights, fMerge,
                                    fActivation, alpha, smoothlabels)
            losses = [triplet_loss(margin=margin), 'categorical_crossentropy'] if nclasses > 0 else triplet_loss(margin=margin)
            return UGaitModel(cfg, optimizer, losses, loss_weights if nclasses > 0 else 1.0, multimodal=True)
        cfg = _cfg_from_args(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units,
                             weight_decay, dropout, margin, nclasses, loss_weights, fMerge, fActivation, alpha, single,
                             smoothlabels=smoothlabels, aux_losses=aux_losses)
        losses = [triplet_loss(margin=margin), 'categorical_crossentropy'] if nclasses > 0 else triplet_loss(margin=margin)
        return UGaitModel(cfg, optimizer, losses, loss_weights if nclasses > 0 else 1.0, multimodal=not single)

    @staticmethod
    def build_or_load(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units=512,
                      weight_decay=1e-4, dropout=0.4, optimizer=None, margin=0.2, nclasses=0,
                      loss_weights=[1.0, 1.0], initnet="", freeze_convs=False, use3D=False, smoothlabels=0,
                      freeze_all=False, postriplet=1, init_branches=None, freeze_branches=False, aux_losses=False,
                      fMerge=Maximum, fActivation='relu', gaitset=False):
        # (the reference's 2-modality build_or_load has no `alpha` argument, :582-588: LeakyReLU keeps build()'s 0.3)
        _unsupported(freeze_convs=freeze_convs, freeze_all=freeze_all)
        if gaitset:
            fActivation = 'leaky'          # :588-589
        model = UWYHSemiNet.build(input_shapes, number_convolutional_layers, filters_size, filters_numbers,
                                  ndense_units, weight_decay, dropout, optimizer, margin, nclasses, loss_weights,
                                  use3D=use3D, smoothlabels=smoothlabels, postriplet=postriplet,
                                  init_branches=init_branches, freeze_branches=freeze_branches, aux_losses=aux_losses,
                                  fMerge=fMerge, fActivation=fActivation, gaitset=gaitset)
        if initnet != "":
            model.load_weights(UWYHSemiNet.get_weights_filename(initnet), by_name=True, skip_mismatch=True)
        return model

    @staticmethod
    def loadnet(netpath: str):
        raise NotImplementedError("loading Keras .hdf5 models needs h5py/TensorFlow; rebuild with build() and "
                                  "model.load_weights(<npz written by save_weights>)")

    @staticmethod
    def fit_generator(model, epochs, callbacks, training_generator, validation_generator, current_step, steps_per_epoch,
                      validation_steps, nworkers=0, new_lr=None):
        if new_lr is not None:
            model.optimizer.lr = new_lr
            print("INFO: learning rate has been changed to {}".format(new_lr))
        hist = model.fit(training_generator, validation_data=validation_generator, epochs=epochs,
                         steps_per_epoch=steps_per_epoch, callbacks=callbacks, validation_steps=validation_steps,
                         initial_epoch=current_step, verbose=2)
        return model, hist

    @staticmethod
    def encode(model, batch_data, use_data, gaitset=False):
        """:971-999 -- branch codes of the first TWO modalities, gated, ALWAYS Maximum (fMerge is ignored
        there), l2-normalised; returns numpy [B, nd]."""
        eng = model.engine
        dev = eng.dev
        if gaitset:
            return _encode_gaitset(model, batch_data, use_data)
        xs = [torch.as_tensor(np.asarray(b), dtype=torch.float32).to(dev) for b in batch_data[:2]]
        B = xs[0].shape[0]
        p = eng.plan(B, False)
        fl = [torch.as_tensor(np.asarray(u), dtype=torch.float32).reshape(-1, 1).to(dev).contiguous() for u in use_data[:2]]
        full_x = xs + [torch.zeros_like(p.br[m].x_in) for m in range(2, eng.cfg.nmods)]
        eng._set_inputs(p, full_x, fl + [torch.zeros(B, 1, device=dev)] * (eng.cfg.nmods - 2))
        eng._forward(p, False)
        sig = torch.zeros(B, eng.cfg.nd, device=dev)
        win = torch.zeros(B, eng.cfg.nd, dtype=torch.uint8, device=dev)
        inv = torch.zeros(B, 2, device=dev)
        ops.fuse_fwd(eng.ctx, [p.br[0].out, p.br[1].out], fl, sig, None, win, inv, MERGE_MAX, True)
        return sig.cpu().numpy()


def _encode_gaitset(model, batch_data, use_data):
    """encode(..., gaitset=True) (:973-979,:988-999): the per-branch [62,B,256] codes of the first two
    modalities (the reference reads them at the 'flatten' / 'flatten_1' layers), gated, Maximum, l2_normalize
    over axis 1; returns numpy in the layout the fusion kernel works in, [62,B,256]."""
    from ugaitnet_b200._ffi import TRef, check, lib, ptr_array, stream_ptr
    eng = model.engine
    dev = eng.dev
    xs = [torch.as_tensor(np.asarray(b), dtype=torch.float32).to(dev) for b in batch_data[:2]]
    B = xs[0].shape[0]
    p = eng.plan(B, False)
    fl = [torch.as_tensor(np.asarray(u), dtype=torch.float32).reshape(-1, 1).to(dev).contiguous() for u in use_data[:2]]
    full_x = xs + [torch.zeros_like(p.br[m].x_in) for m in range(2, eng.cfg.nmods)]
    eng._set_inputs(p, full_x, fl + [torch.zeros(B, 1, device=dev)] * (eng.cfg.nmods - 2))
    eng._forward(p, False)
    n, d = p.sig.shape[0], p.sig.shape[2]
    sig = torch.zeros(n, B, d, device=dev)
    win = torch.zeros(n, B, d, dtype=torch.uint8, device=dev)
    col = torch.zeros(n, d, 2, device=dev)
    Rb, Rf = [p.br[0].R["out"], p.br[1].R["out"]], [TRef(f) for f in fl]
    Rs, Rw, Rc = TRef(sig), TRef(win), TRef(col)
    check(lib.ugn_fuse3_fwd(eng.ctx.h, 2, ptr_array(Rb), ptr_array(Rf), Rs.ptr, Rw.ptr, Rc.ptr, MERGE_MAX, stream_ptr()))
    return sig.cpu().numpy()


class UWYHSemiNet3Mods(UWYHSemiNet):
    def __init__(self):
        super().__init__()

    @staticmethod
    def build(input_shapes, number_convolutional_layers, filters_size, filters_numbers,
              ndense_units=512, weight_decay=1e-4,
              dropout=0.4, optimizer=None, margin=0.2,
              nclasses=0, loss_weights=[1.0, 1.0], use3D=False, smoothlabels=0,
              postriplet=1, init_branches=None, freeze_branches=False, aux_losses=False, fMerge=Maximum,
              normbfmerge=False, fActivation='relu', alpha=0.3, gaitset=False):
        _unsupported(use3D=use3D, aux_losses_with_gaitset=(aux_losses and gaitset), freeze_branches=freeze_branches,
                     init_branches=bool(init_branches) and any(init_branches.values()))
        if gaitset:
            cfg = _gs_cfg_from_args(input_shapes, ndense_units, dropout, margin, nclasses, loss_weights, fMerge,
                                    fActivation, alpha, smoothlabels)       # (normbfmerge is ignored with gaitset, :1164)
            losses = [triplet_loss(margin=margin), 'categorical_crossentropy'] if nclasses > 0 else triplet_loss(margin=margin)
            return UGaitModel(cfg, optimizer, losses, loss_weights if nclasses > 0 else 1.0, multimodal=True)
        cfg = _cfg_from_args(list(input_shapes), number_convolutional_layers, filters_size, filters_numbers,
                             ndense_units, weight_decay, dropout, margin, nclasses, loss_weights, fMerge, fActivation,
                             alpha, single=False, smoothlabels=smoothlabels, normbfmerge=normbfmerge,
                             aux_losses=aux_losses)
        losses = [triplet_loss(margin=margin), 'categorical_crossentropy'] if nclasses > 0 else triplet_loss(margin=margin)
        return UGaitModel(cfg, optimizer, losses, loss_weights if nclasses > 0 else 1.0, multimodal=True)

    @staticmethod
    def compile_hard(model, optimizer, loss_weights, margin):
        raise NotImplementedError("tfa.losses.TripletHardLoss is outside the B200 hot path")

    @staticmethod
    def build_or_load(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units=512,
                      weight_decay=1e-4, dropout=0.4, optimizer=None, margin=0.2, nclasses=0,
                      loss_weights=[1.0, 1.0], initnet="", freeze_convs=False, use3D=False, smoothlabels=0,
                      freeze_all=False, postriplet=1, init_branches=None, freeze_branches=False, aux_losses=False,
                      fMerge=Maximum, normbfmerge=False, fActivation='relu', alpha=0.3, gaitset=False):
        _unsupported(freeze_convs=freeze_convs, freeze_all=freeze_all)
        if gaitset:
            fActivation = 'leaky'
        model = UWYHSemiNet3Mods.build(input_shapes, number_convolutional_layers, filters_size, filters_numbers,
                                       ndense_units, weight_decay, dropout, optimizer, margin, nclasses,
                                       loss_weights, use3D=use3D, smoothlabels=smoothlabels, postriplet=postriplet,
                                       init_branches=init_branches, freeze_branches=freeze_branches,
                                       aux_losses=aux_losses, fMerge=fMerge, normbfmerge=normbfmerge,
                                       fActivation=fActivation, alpha=alpha, gaitset=gaitset)
        if initnet != "":
            model.load_weights(UWYHSemiNet.get_weights_filename(initnet), by_name=True, skip_mismatch=True)
        return model
