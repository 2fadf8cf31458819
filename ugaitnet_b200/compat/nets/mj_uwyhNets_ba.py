"""Drop-in for the model builders of nets/mj_uwyhNets_ba.py (batch-all version).

Same class / static-method names, positional order, defaults and return types as the reference
(UWYHSemiNet3Mods.build :1032-1037, build_or_load :1309-1316, UWYHSemiNet.build :669-672,
fit_generator :938-968, encode :971-999); the returned object follows the slice of the Keras Model
protocol the reference's mains/ scripts use (fit / predict / get_layer / save_weights / ...), and every
step runs on the B200 engine (ugaitnet_b200.net.UGaitEngine).

``gaitset=True`` (with a non-'relu' fActivation, as the reference requires, :1101) builds the GaitSet
branch type (build_gaitset_branch :420-484) on ugaitnet_b200.gaitset.GaitSetEngine: inputs
[B,25,60,60,c], signature [62,B,256], descriptor layer "flatten" (typecode 3); on three, two or ONE input shape
(UWYHSemiNet.build_or_load(one shape, gaitset=True), :588-589 + :776-777 + :890-905: the branch output is the
signature -- what mains/mj_trainUWYHGaitNet_DataGen_CasiaB_1mod.py:329-333 builds) and with postriplet == 2 (:814-832).

``smoothlabels`` (label-smoothed cross-entropy, :1252-1262), ``normbfmerge`` (per-branch l2_normalize before
the gate, :1167-1168), ``aux_losses`` (classprob_{of,gray,depth} heads on the gated branch outputs, :1222-1251;
stacked-CNN branches), ``postriplet == 2`` (2-modality builder, :819-832), ``freeze_convs`` / ``freeze_all`` /
``freeze_branches`` and ``layer.trainable`` (:193, :1366-1391), ``initnet`` / ``init_branches`` / ``loadnet`` from Keras
HDF5 files (ugaitnet_b200.hdf5: pure-Python reader / writer, h5py is absent) are implemented.  Builder arguments that
select graphs that are ill-formed in the reference itself raise NotImplementedError: aux_losses together with gaitset,
use3D with a multi-modal gaitset graph, compile_hard with gaitset.  use3D builds the Conv3D branches
(:336-417) for every modality but a 50-channel optical flow; they run on the fp32 validation engine.  compile_hard (tfa TripletHardLoss) re-compiles a stacked-frame CNN model onto ugn_triplet_hard.
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

# the benchmarked, parity-gated math mode (tests/test_decisions_gpu.py): fp16 hi/lo split, 3-pass forward, 1-pass backward
MATH_MODE = os.environ.get("UGN_MATH_MODE", "f16mix")


class TripletHardLoss:
    """Stand-in for `tfa.losses.TripletHardLoss(margin)` in model.loss after compile_hard (:1303); callable on
    (y_true [m,1], y_pred [m,d]) like the Keras loss object, computed by ugn_triplet_hard."""
    name = "triplet_hard_loss"

    def __init__(self, margin=1.0, soft=False, distance_metric="L2", **kw):
        if soft or distance_metric != "L2":
            raise NotImplementedError("TripletHardLoss: only soft=False, distance_metric='L2' (what compile_hard uses)")
        self.margin = margin

    def __call__(self, y_true, y_pred):
        dev = torch.device("cuda", torch.cuda.current_device())
        emb = torch.as_tensor(np.asarray(y_pred) if not torch.is_tensor(y_pred) else y_pred,
                              dtype=torch.float32).to(dev).contiguous()
        lab = torch.as_tensor(np.asarray(y_true) if not torch.is_tensor(y_true) else y_true).to(dev)
        lab = lab.reshape(-1).to(torch.int32).contiguous()
        out = torch.zeros(2, device=dev)
        ws = torch.zeros(ops.triplet_workspace_bytes(1, emb.shape[0]) // 4 + 8, device=dev)
        ops.triplet_hard(ops.get_ctx(dev.index), emb, lab, float(self.margin), 1.0, out, None, ws)
        return out[0]


def mj_tensor_times_scalar(d):
    return d[0] * d[1]


def _unsupported(**flags):
    if flags.get("aux_losses_with_gaitset"):
        # :1222-1229 puts Dense(nclasses) on the GATED branch outputs, which are [62, B, 256] under gaitset (:1162): the
        # heads emit [62, B, nclasses] against one-hot targets [B, nclasses] -- Keras rejects the shapes at fit time
        raise NotImplementedError("aux_losses with gaitset: the reference graph itself is ill-formed (auxiliary heads on "
                                  "[62, B, 256] gated outputs vs [B, nclasses] targets, nets/mj_uwyhNets_ba.py:1222-1229)")
    bad = [k for k, v in flags.items() if v]
    if bad:
        raise NotImplementedError(f"builder options outside the B200 hot path: {', '.join(bad)} "
                                  "(see DESIGN.md, 'out of scope')")


class _LayerProxy:
    def __init__(self, model, name, units=None, sublayers=(), kind="layer"):
        self.model, self.name, self.units, self._trainable = model, name, units, True
        self.layers = list(sublayers)
        self.kind = kind                 # "conv" | "dense" | "layer": freeze_convs freezes the Conv2D sublayers only

    @property
    def trainable(self):
        return self._trainable

    @trainable.setter
    def trainable(self, flag):
        """layer.trainable = False (:1366-1391): the optimiser stops updating the layer's tensors."""
        self._trainable = bool(flag)
        for sub in self.layers:
            sub.trainable = flag
        self.model._set_trainable(self.name, bool(flag))

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
        has3d = bool(getattr(cfg, "branch3d", ()))       # use3D: Conv3D branches run on the fp32 validation engine
        self.engine = engine_cls(cfg, math_mode="fp32" if has3d else MATH_MODE, optimizer=getattr(opt, "name", "sgd"),
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
            elif cfg.is3d(m):
                sub = [_LayerProxy(self, f"{BRANCH_NAMES[m]}/conv{i}", kind="conv") for i in range(len(cfg.filters3d))] + \
                      [_LayerProxy(self, f"{BRANCH_NAMES[m]}/ofCode", kind="dense")]
            else:
                sub = [_LayerProxy(self, f"{BRANCH_NAMES[m]}/conv{i}", kind="conv") for i in range(len(cfg.filters_numbers))] + \
                      [_LayerProxy(self, f"{BRANCH_NAMES[m]}/{n}", kind="dense" if n in ("dense", "ofCode") else "layer")
                       for n in ["ofFlat", "dense", "drop", "ofCode"]]
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
        """model.predict(list): [signature, classprob] from ONE forward pass (single-copy input path)."""
        eng = self.engine
        xs = [x[2 * m] for m in range(self.cfg.nmods)] if self.multimodal else [x[0] if isinstance(x, (list, tuple)) else x]
        B = int(np.shape(xs[0])[0])
        if not hasattr(self, "_hbp"):
            self._hbp = {}
        hb = self._hbp.get(B)
        if hb is None:
            hb = self._hbp[B] = eng.host_batch(B, train=False)
        self._fill_and_send(hb, xs, x, train=False)
        sig = eng.predict_prefetched("embedding")      # the model's output 0 (postriplet == 2: the normalised "code")
        if self.cfg.nclasses > 0:
            p = eng.plan(B, False)
            prob = torch.softmax(p.logits, dim=1)
            return [sig.cpu().numpy(), prob.cpu().numpy()]
        return sig.cpu().numpy()

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
        if "act_reg" in out:            # activity regulariser of "code" (:1196): part of Keras' total loss
            total += float(out["act_reg"])
        logs[prefix + "loss"] = total
        return logs

    def train_on_batch(self, x, y, **kw):
        ins, fl = self._split_x(x)
        return self._logs(self.engine.train_step(ins, fl, self._labels(y)))

    # -- pipelined input path of fit(): generator batch (float64 numpy, data/mj_dataGeneratorMMUWYHsingle.py:664-823)
    #    -> multi-threaded cast straight into a pinned HostBatch -> ONE cudaMemcpyAsync -> step; the cast / copy of batch
    #    i+1 runs while the GPU works on batch i
    def _stage(self, X, y):
        eng = self.engine
        xs = [X[2 * m] for m in range(self.cfg.nmods)] if self.multimodal else [X[0] if isinstance(X, (list, tuple)) else X]
        B = int(np.shape(xs[0])[0])
        if not hasattr(self, "_hb"):
            self._hb, self._hb_k = {}, 0
        pair = self._hb.get(B)
        if pair is None:
            pair = self._hb[B] = [eng.host_batch(B), eng.host_batch(B)]
        self._hb_k ^= 1
        hb = pair[self._hb_k]
        lab = y[0] if isinstance(y, (list, tuple)) else y
        hb.labels[...] = np.asarray(lab).reshape(-1).astype(np.int32)
        self._fill_and_send(hb, xs, X, train=True)

    def _fill_and_send(self, hb, xs, X, train):
        """Header (flags; the caller has written the labels) first, then volume after volume: the H2D copy of modality m
        (UGaitEngine.prefetch_upto, a prefix of the ONE pinned block) runs underneath the f64 -> f32 cast of modality
        m + 1, so a batch is on the device one last-volume copy after the loader has finished with it."""
        eng, io = self.engine, hb.io
        if self.multimodal:
            for m in range(len(xs)):
                hb.flags[m][...] = np.asarray(X[2 * m + 1], dtype=np.float32).reshape(-1, 1)
        eng.prefetch_open(hb, train)
        for m, x in enumerate(xs):
            src = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
            hb.t["x"][m].copy_(src.reshape(hb.t["x"][m].shape))          # f64 -> f32 cast across the host cores
            eng.prefetch_upto(io.x_off[m + 1] if m + 1 < len(xs) else hb.nbytes)
        eng.prefetch_close()

    def _fit_epoch(self, gen, n):
        """One epoch through the single-copy path; returns the per-batch logs (each read back with ONE D2H copy)."""
        eng, logs = self.engine, []
        ln = len(gen)
        X, y = gen[0][:2]
        self._stage(X, y)
        pend = None
        for i in range(n):
            out = eng.train_step_prefetched()
            if i + 1 < n:
                X, y = gen[(i + 1) % ln][:2]
                self._stage(X, y)                        # overlaps the step that was just launched
            logs.append(self._logs_from_pack(out["losses"].cpu(), out))
        return logs

    def _logs_from_pack(self, pack, out):
        """The same dictionary as _logs, from the packed loss buffer [triplet, count, ce, acc, reg, act_reg]."""
        cfg = self.cfg
        trip, ce, acc, reg, areg = float(pack[0]), float(pack[2]), float(pack[3]), float(pack[4]), float(pack[5])
        logs, total = {}, cfg.wver * trip
        if cfg.nclasses > 0:
            logs["signature_loss"], logs["classprob_loss"], logs["classprob_acc"] = trip, ce, acc
            total += cfg.wid * ce
            for m, v in enumerate(out.get("aux_ce", [])):
                name = ("classprob_of", "classprob_gray", "classprob_depth")[m]
                logs[name + "_loss"], logs[name + "_acc"] = float(v), float(out["aux_acc"][m])
                total += cfg.waux * float(v)
        logs["loss"] = total + reg + (areg if cfg.nc > 0 else 0.0)
        return logs

    def test_on_batch(self, x, y, **kw):
        ins, fl = self._split_x(x)
        return self._logs(self.engine.eval_losses(ins, fl, self._labels(y)))

    def fit(self, x=None, validation_data=None, epochs=1, steps_per_epoch=None, callbacks=None,
            validation_steps=None, initial_epoch=0, verbose=2, **kw):
        gen, hist = x, History()
        callbacks = callbacks or []
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        for epoch in range(initial_epoch, epochs):
            n = steps_per_epoch or len(gen)
            acc: Dict[str, float] = {}
            for step_logs in self._fit_epoch(gen, n):
                for k, v in step_logs.items():
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
                    cb.on_epoch_end(epoch, logs)      # a failing checkpoint / LR callback must surface, as in Keras
            if self.stop_training:
                break
        return hist

    # -- weights: Keras names / layouts ((kh,kw,cin,cout) conv kernels, (in,out) dense kernels) -------
    def _set_trainable(self, name, flag):
        try:
            self.engine.set_trainable(name, flag)
        except KeyError:
            pass                          # a layer without parameters (gates, fusion, dropout ...)

    @staticmethod
    def _keras_name(seg_name):
        base, kind = seg_name.rsplit("/", 1)
        return f"{base}/{'kernel' if kind == 'w' else 'bias'}"

    def _keras_layout(self):
        """[(keras layer name, [(keras weight name, engine tensor name), ...]), ...] in the order Keras lists
        model.layers / layer.weights: a branch is ONE layer (a Sequential) whose weights are kernel, bias of every
        sublayer in graph order; the Conv2D sublayers carry Keras' auto-generated names (conv2d, conv2d_1, ...)."""
        e, out, conv_no, conv3_no = self.engine, [], 0, 0
        top: Dict[str, list] = {}
        b3 = {BRANCH_NAMES[m] for m in range(self.cfg.nmods) if getattr(self.cfg, "is3d", None) and self.cfg.is3d(m)}
        for sg in e.seg_list:
            parts = sg.name.split("/")
            kind = "kernel:0" if parts[-1] == "w" else "bias:0"
            if len(parts) == 3:                        # <branch>/<sublayer>/<w|b>
                sub = parts[1]
                if parts[0] in b3:                     # Conv3D branch (:346-363): conv3d, conv3d_1, ..., then "grayCode"
                    if parts[-1] == "w":
                        conv3_no += 1
                    sub = "grayCode" if sub == "ofCode" else ("conv3d" if conv3_no == 1 else f"conv3d_{conv3_no - 1}")
                elif sub.startswith("conv") and not self.gaitset:
                    if parts[-1] == "w":
                        conv_no += 1
                    sub = "conv2d" if conv_no == 1 else f"conv2d_{conv_no - 1}"
                top.setdefault(parts[0], []).append((f"{sub}/{kind}", sg.name))
            else:                                      # code / classprob / classprob_of ...
                top.setdefault(parts[0], []).append((f"{parts[0]}/{kind}", sg.name))
        for l in self.layers:
            out.append((l.name, top.get(l.name, [])))
        return out

    def _layer_weights(self, name):
        """layer.get_weights(): Keras order -- layers in graph order, kernel before bias
        (consumers index it: nets/mj_utils.py:159-160 ``filters = w[0]``)."""
        P = self.engine.export_params()
        out = []
        for sg in self.engine.seg_list:               # arena order = graph order, w before b
            k = sg.name
            if k.startswith(name + "/") or k.rsplit("/", 1)[0] == name:
                out.append(self._to_keras(k, P[k]).cpu().numpy())
        return out

    @staticmethod
    def _to_keras(k, v):
        if v.dim() == 5:
            return v.permute(2, 3, 4, 1, 0).contiguous()   # Conv3D [Cout,Cin,kt,kh,kw] -> (kt,kh,kw,cin,cout)
        if v.dim() == 4:
            return v.permute(2, 3, 1, 0).contiguous()      # [Cout,Cin,kh,kw] -> (kh,kw,cin,cout)
        if v.dim() == 2:
            return v.t().contiguous()                      # [out,in] -> (in,out)
        return v

    @staticmethod
    def _from_keras(v):
        v = torch.as_tensor(np.asarray(v))
        if v.dim() == 5 and tuple(v.shape[:3]) == (1, 1, 1):
            return v.reshape(v.shape[3], v.shape[4]).t().contiguous()     # "grayCode": Conv3D 1x1x1 == Dense
        if v.dim() == 5:
            return v.permute(4, 3, 0, 1, 2).contiguous()
        if v.dim() == 4:
            return v.permute(3, 2, 0, 1).contiguous()
        if v.dim() == 2:
            return v.t().contiguous()
        return v

    def _write_weight_groups(self, w, root: str):
        """Keras' hdf5_format.save_weights_to_hdf5_group layout: attrs layer_names / backend / keras_version on the
        group, one sub-group per layer with attr weight_names and one dataset per weight."""
        P = self.engine.export_params()
        g = w.group(root)
        layout = self._keras_layout()
        g.attrs["layer_names"] = [n.encode("utf8") for n, _ in layout]
        g.attrs["backend"] = b"tensorflow"
        g.attrs["keras_version"] = b"2.4.0"
        for lname, weights in layout:
            lg = w.group(f"{root}/{lname}")
            lg.attrs["weight_names"] = [wn.encode("utf8") for wn, _ in weights] if weights else np.zeros((0,), dtype="S1")
            for wn, seg in weights:
                arr = self._to_keras(seg, P[seg]).cpu().numpy()
                if wn == "grayCode/kernel:0":            # Conv3D(nd, (1,1,1)) of a use3D branch: (1,1,1,in,out)
                    arr = arr.reshape((1, 1, 1) + arr.shape)
                w.dataset(f"{root}/{lname}/{wn}", arr)

    def save_weights(self, path, **kw):
        """model.save_weights(path): a regular HDF5 file in Keras' layout (ugaitnet_b200.hdf5 writer)."""
        from ugaitnet_b200 import hdf5
        w = hdf5.Writer()
        self._write_weight_groups(w, "/")
        w.save(path)

    def save(self, path, **kw):
        """model.save(path): Keras full-model layout -- weights under /model_weights, the builder arguments as a JSON
        attribute (what loadnet() rebuilds the graph from) and the optimiser state under /optimizer_weights."""
        import json
        from ugaitnet_b200 import hdf5
        w, e = hdf5.Writer(), self.engine
        self._write_weight_groups(w, "/model_weights")
        w.root.attrs["backend"] = b"tensorflow"
        w.root.attrs["keras_version"] = b"2.4.0"
        w.root.attrs["ugn_builder_config"] = json.dumps(getattr(self, "builder_config", {})).encode("utf8")
        og = w.group("/optimizer_weights")
        og.attrs["optimizer"] = e.optimizer.encode("utf8")
        og.attrs["iterations"] = np.int64(e.t)
        og.attrs["lr"] = np.float64(e.lr)
        w.dataset("/optimizer_weights/m", e.m.cpu().numpy())
        w.dataset("/optimizer_weights/v", e.v.cpu().numpy())
        if getattr(e, "vhat", None) is not None:
            w.dataset("/optimizer_weights/vhat", e.vhat.cpu().numpy())
        w.save(path)

    def load_branch(self, path, bname):
        return _load_branch(self, path, bname)

    def save_branch(self, path, bname):
        """Stand-alone branch file in the layout of Keras' Sequential.save: one layer group per sublayer."""
        from ugaitnet_b200 import hdf5
        P = self.engine.export_params()
        w = hdf5.Writer()
        weights = dict(self._keras_layout())[bname]
        subs = []
        for wn, seg in weights:
            sub = wn.split("/")[0]
            if sub not in subs:
                subs.append(sub)
            w.dataset(f"/model_weights/{sub}/{wn}", self._to_keras(seg, P[seg]).cpu().numpy())
        g = w.group("/model_weights")
        g.attrs["layer_names"] = [s.encode("utf8") for s in subs]
        for sub in subs:
            w.group(f"/model_weights/{sub}").attrs["weight_names"] = [wn.encode("utf8") for wn, _ in weights
                                                                      if wn.split("/")[0] == sub]
        w.save(path)

    def _load_npz(self, path, skip_mismatch):
        z = np.load(path)
        mine = {self._keras_name(k): k for k in self.engine.segs}
        upd = {}
        for kn in z.files:
            k = mine.get(kn)
            if k is None:
                continue
            v = self._from_keras(z[kn])
            if tuple(v.shape) != tuple(self.engine.oracle_shape(k)):
                if skip_mismatch:
                    continue
                raise ValueError(f"shape mismatch for {kn}: {tuple(v.shape)} vs {tuple(self.engine.oracle_shape(k))}")
            upd[k] = v
        self.engine.load_params(upd)

    def load_weights(self, path, by_name=True, skip_mismatch=False, **kw):
        """Keras load_weights on an HDF5 weights file (or the /model_weights group of a full-model file).
        by_name=True: layers are matched by name and their weights by position inside the layer (Keras'
        load_weights_from_hdf5_group_by_name); by_name=False: all weights in file order against all tensors in graph
        order (topological loading).  skip_mismatch skips layers whose weight count or shapes differ."""
        from ugaitnet_b200 import hdf5
        if not hdf5.is_hdf5(path):
            with open(path, "rb") as fh:
                magic = fh.read(4)
            if magic[:2] == b"PK":              # round-1 checkpoints: npz container under the .hdf5 name
                return self._load_npz(path, skip_mismatch)
            raise ValueError(f"{path}: neither an HDF5 file nor a ugaitnet_b200 npz checkpoint")
        f = hdf5.File(path)
        g = f["model_weights"] if "model_weights" in f.keys() else f
        names = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.atleast_1d(g.attrs.get("layer_names", []))]
        file_layers = []
        for ln in names:
            lg = g[ln]
            wn = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.atleast_1d(lg.attrs.get("weight_names", []))]
            file_layers.append((ln, [lg[n].value for n in wn if n]))
        layout = dict(self._keras_layout())
        upd = {}

        def assign(segs, vals, what):
            if len(segs) != len(vals):
                if skip_mismatch:
                    return
                raise ValueError(f"layer {what}: {len(vals)} weights in the file, {len(segs)} in the model")
            tmp = {}
            for seg, v in zip(segs, vals):
                t = self._from_keras(v)
                if tuple(t.shape) != tuple(self.engine.oracle_shape(seg)):
                    if skip_mismatch:
                        return
                    raise ValueError(f"shape mismatch for {seg}: {tuple(t.shape)} vs {tuple(self.engine.oracle_shape(seg))}")
                tmp[seg] = t
            upd.update(tmp)
        if by_name:
            for ln, vals in file_layers:
                if ln in layout and (vals or layout[ln]):
                    assign([s for _, s in layout[ln]], vals, ln)
        else:
            assign([s for _, ws in self._keras_layout() for _, s in ws], [v for _, vs in file_layers for v in vs], "(all)")
        self.engine.load_params(upd)
        if "optimizer_weights" in f.keys():
            og, e = f["optimizer_weights"], self.engine
            m = og["m"].value
            if m.shape[0] == e.m.shape[0]:
                e.m.copy_(torch.from_numpy(m))
                e.v.copy_(torch.from_numpy(og["v"].value))
                if "vhat" in og.keys():
                    from ugaitnet_b200._ffi import TRef
                    e.vhat = torch.from_numpy(og["vhat"].value).to(e.dev)
                    e.R["vhat"] = TRef(e.vhat)
                e.t = int(og.attrs.get("iterations", 0))
                e.lr = float(og.attrs.get("lr", e.lr))
        return self


class PairModel(UGaitModel):
    """Model object of UWYHNet.build (:154-245): nine inputs, one output (the verification loss)."""
    dtype = "float32"

    def __init__(self, cfg, optimizer):
        opt = optimizer if optimizer is not None else optimizers.SGD(0.001, 0.9)
        kw = dict(getattr(opt, "kw", {}))
        self.cfg, self.multimodal, self.gaitset = cfg, True, False
        self.engine = UGaitEngine(cfg, math_mode=MATH_MODE, optimizer=getattr(opt, "name", "sgd"),
                                  lr=getattr(opt, "lr", 0.001), momentum=kw.get("momentum", 0.9),
                                  beta1=kw.get("beta1", 0.9), beta2=kw.get("beta2", 0.999), eps=kw.get("eps", 1e-7),
                                  lr_decay=kw.get("decay", 0.0), decoupled_weight_decay=kw.get("weight_decay", 0.0),
                                  use_graph=os.environ.get("UGN_GRAPH", "1") == "1")
        self.optimizer = _OptimizerView(self)
        self.loss, self.loss_weights, self.stop_training = None, [1.0], False
        sub = lambda bn: [_LayerProxy(self, f"{bn}/conv{i}", kind="conv") for i in range(len(cfg.filters_numbers))] + \
                         [_LayerProxy(self, f"{bn}/{n}", kind="dense" if n in ("dense", "ofCode") else "layer")
                          for n in ["ofFlat", "dense", "drop", "ofCode"]]
        self.layers = [_LayerProxy(self, BRANCH_NAMES[m], units=cfg.nd, sublayers=sub(BRANCH_NAMES[m])) for m in range(2)]
        self.layers += [_LayerProxy(self, n) for n in ("gate_of1", "gate_gray1", "agg1", "embedL2_1", "gate_of2",
                                                       "gate_gray2", "agg2", "embedL2_2", "loss_pair")]
        self.input = [_Tag(self, n) for n in ("ofinput1", "ofuse1", "grayinput1", "grayuse1", "ofinput2", "ofuse2",
                                              "grayinput2", "grayuse2", "label")]

    def _stack(self, x):
        cu = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda(non_blocking=True)
        if len(x) != 9:
            raise ValueError(f"the pair network takes 9 inputs (:232), got {len(x)}")
        ins = [torch.cat([cu(x[0]), cu(x[4])]), torch.cat([cu(x[2]), cu(x[6])])]
        fl = [torch.cat([cu(x[1]).reshape(-1, 1), cu(x[5]).reshape(-1, 1)]),
              torch.cat([cu(x[3]).reshape(-1, 1), cu(x[7]).reshape(-1, 1)])]
        lab = torch.as_tensor(np.asarray(x[8])).reshape(-1).to(torch.int32).cuda()
        return ins, fl, lab

    def train_on_batch(self, x, y=None, **kw):
        out = self.engine.train_step(*self._stack(x))
        return float(out["triplet"]) * self.cfg.wver + float(out["reg"])

    def test_on_batch(self, x, y=None, **kw):
        out = self.engine.eval_losses(*self._stack(x))
        return float(out["triplet"])

    def predict(self, x, batch_size=None, verbose=0):
        """The model's output tensor is the pair loss of the batch (:230-235)."""
        return np.float32(self.test_on_batch(x))

    def embed(self, x4):
        """Normalised signature of ONE side ([of, ofuse, gray, grayuse]) -- the `embedL2_1` layer."""
        cu = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()
        return self.engine.predict([cu(x4[0]), cu(x4[2])], [cu(x4[1]).reshape(-1, 1), cu(x4[3]).reshape(-1, 1)]).cpu().numpy()

    def fit(self, x=None, epochs=1, steps_per_epoch=None, callbacks=None, initial_epoch=0, verbose=2, **kw):
        hist = History()
        for epoch in range(initial_epoch, epochs):
            n = steps_per_epoch or len(x)
            tot = 0.0
            for i in range(n):
                item = x[i % len(x)]
                tot += self.train_on_batch(item[0] if isinstance(item, tuple) else item)
            hist.epoch.append(epoch)
            hist.history.setdefault("loss", []).append(tot / n)
            hist.history.setdefault("lr", []).append(float(self.engine.lr))
            if hasattr(x, "on_epoch_end"):
                x.on_epoch_end()
        return hist

    def summary(self):
        print(f"UWYHNet pair network (B200 engine, math={self.engine.math_mode}): 2 shared branches, nd={self.cfg.nd}")



def _file_layers(g):
    names = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.atleast_1d(g.attrs.get("layer_names", []))]
    out = []
    for ln in names:
        lg = g[ln]
        wn = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.atleast_1d(lg.attrs.get("weight_names", []))]
        out.append((ln, [lg[n].value for n in wn if n]))
    return out


def _load_branch(model, path, bname):
    """fc_loadBranch (:57-66): initialise branch `bname` from a saved model -- a stand-alone branch file (Keras
    Sequential.save: its sublayers are the file's layers) or a full model that contains a layer of that name."""
    from ugaitnet_b200 import hdf5
    f = hdf5.File(path)
    g = f["model_weights"] if "model_weights" in f.keys() else f
    layers = _file_layers(g)
    byname = dict(layers)
    vals = byname[bname] if bname in byname and byname[bname] else [v for _, vs in layers for v in vs]
    segs = [s for _, s in dict(model._keras_layout())[bname]]
    if len(vals) != len(segs):
        raise ValueError(f"{path}: {len(vals)} weights for branch {bname}, the graph has {len(segs)}")
    upd = {}
    for seg, v in zip(segs, vals):
        t = model._from_keras(v)
        if tuple(t.shape) != tuple(model.engine.oracle_shape(seg)):
            raise ValueError(f"{path}: shape mismatch for {seg}: {tuple(t.shape)} vs {tuple(model.engine.oracle_shape(seg))}")
        upd[seg] = t
    model.engine.load_params(upd)


def _cfg_from_args(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units,
                   weight_decay, dropout, margin, nclasses, loss_weights, fMerge, fActivation, alpha, single,
                   smoothlabels=0, normbfmerge=False, aux_losses=False, postriplet=1, use3D=False):
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
    # use3D (:721-745, :1077-1099): every modality but a 50-channel optical flow gets a Conv3D branch on [25,60,60,1]
    b3 = tuple(bool(use3D) and int(s[0]) != 50 for s in shapes)
    if any(b3) and (normbfmerge or aux_losses):
        _unsupported(use3D_with_normbfmerge_or_aux_losses=True)
    return NetConfig(branch3d=b3 if any(b3) else (), in_channels=tuple(int(s[0]) for s in shapes), filters_numbers=tuple(fn), filters_size=tuple(fs),
                     nd=int(nd), nc=int(nc) if not single else 0, nclasses=int(nclasses), weight_decay=float(weight_decay),
                     merge=merge_id_of(fMerge) if not single else MERGE_MAX,
                     act=ACT_RELU if fActivation == "relu" else ACT_LEAKY, alpha=float(alpha), margin=float(margin),
                     wver=float(lw[0]) if nclasses > 0 else 1.0, wid=float(lw[1]) if nclasses > 0 and len(lw) > 1 else 0.0,
                     hw=int(shapes[0][1]), dropout=float(dropout) if dropout > 0.001 else 0.0, single=single,
                     label_smoothing=float(smoothlabels), normbfmerge=bool(normbfmerge),
                     aux_losses=bool(aux_losses) and nclasses > 0 and not single,
                     waux=float(lw[-1]),         # loss_weights padded with its last entry (:1264-1268)
                     postriplet=int(postriplet) if (nc and not single) else 1)


def _gs_cfg_from_args(input_shapes, ndense_units, dropout, margin, nclasses, loss_weights, fMerge, fActivation, alpha,
                      smoothlabels=0, single=False, postriplet=1):
    """gaitset=True: input_shapes [(25,60,60,2), (25,60,60,1), ...] (mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:212-213);
    single: ONE shape (25,60,60,c) -> the 1-modality graph (UWYHSemiNet.build :776-777, :890-905: the branch output is the
    signature; no gate / fusion / l2_normalize, no FC1 (`if False and add_extra_dense`), plain cross-entropy)."""
    if fActivation == 'relu':
        raise ValueError("gaitset=True needs a non-'relu' fActivation: the reference only builds the GaitSet "
                         "branches in its LeakyReLU path (nets/mj_uwyhNets_ba.py:1101-1113)")
    shapes = [tuple(input_shapes)] if single else list(input_shapes)
    if any(len(s) != 4 for s in shapes):
        raise ValueError("gaitset=True expects input shapes (frames, H, W, channels)")
    nc = ndense_units[1] if isinstance(ndense_units, (list, tuple)) and len(ndense_units) > 1 else 0
    if isinstance(dropout, (list, tuple)):
        dropout = dropout[-1]
    lw = list(loss_weights) if isinstance(loss_weights, (list, tuple)) else [loss_weights, loss_weights]
    if single:
        return GaitSetConfig(in_channels=(int(shapes[0][3]),), frames=int(shapes[0][0]), hw=int(shapes[0][1]), nc=0,
                             nclasses=int(nclasses), alpha=float(alpha), margin=float(margin),
                             wver=float(lw[0]) if nclasses > 0 else 1.0,
                             wid=float(lw[1]) if nclasses > 0 and len(lw) > 1 else 0.0, dropout=0.0, single=True)
    return GaitSetConfig(in_channels=tuple(int(s[3]) for s in shapes), frames=int(shapes[0][0]), hw=int(shapes[0][1]),
                         nc=int(nc), nclasses=int(nclasses), merge=merge_id_of(fMerge), alpha=float(alpha),
                         margin=float(margin), wver=float(lw[0]) if nclasses > 0 else 1.0,
                         wid=float(lw[1]) if nclasses > 0 and len(lw) > 1 else 0.0,
                         dropout=float(dropout) if (dropout > 0.001 and nc) else 0.0, label_smoothing=float(smoothlabels),
                         postriplet=int(postriplet) if nc else 1)


def _jsonable(v):
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if callable(v) or hasattr(v, "merge_id"):
        return {"__merge__": {0: "Maximum", 1: "Average", 2: "sign_max"}[merge_id_of(v)]}
    if hasattr(v, "name") and hasattr(v, "kw"):          # optimizer stand-in
        return {"__optimizer__": v.name, "lr": v.lr, "kw": {k: float(x) for k, x in v.kw.items()}}
    return v


def _unjson(v):
    from ugaitnet_b200.compat.keras_shim import _Optimizer, sign_max
    if isinstance(v, dict) and "__merge__" in v:
        return {"Maximum": Maximum, "Average": Average, "sign_max": sign_max}[v["__merge__"]]
    if isinstance(v, dict) and "__optimizer__" in v:
        return _Optimizer(v["__optimizer__"], v["lr"], **v["kw"])
    if isinstance(v, list):
        return [_unjson(x) for x in v]
    return v


def _remember(model, cls_name, **kwargs):
    """The builder call that made `model`, JSON-able: model.save() stores it, loadnet() rebuilds from it (the role of
    Keras' model_config attribute / the reference's model-config.hdf5)."""
    model.builder_config = {"class": cls_name, "kwargs": {k: _jsonable(v) for k, v in kwargs.items()}}
    return model


def _apply_init_branches(model, init_branches):
    """init_branches = {'of': path, 'gray': path, 'depth': path} (:57-66 fc_loadBranch, :69-75): every given branch
    is initialised from a saved model file -- here: the tensors of the same-named branch layer of that file."""
    if not init_branches:
        return
    for key, bname in (("of", "ofBranch"), ("gray", "grayBranch"), ("depth", "depthBranch")):
        path = init_branches.get(key, "")
        if path:
            model.load_branch(path, bname)


def _freeze(model, freeze_convs=False, freeze_all=False, freeze_branches=False):
    """:193 (freeze_branches: whole branches), :1366-1391 (freeze_convs: the Conv2D layers of every branch; freeze_all:
    every layer of every branch -- with gaitset every model layer but the last)."""
    if not (freeze_convs or freeze_all or freeze_branches):
        return
    if model.gaitset:
        if freeze_all:
            for l in model.layers[:-1]:
                l.trainable = False
        elif freeze_branches:
            for l in model.layers:
                if l.name in BRANCH_NAMES:
                    l.trainable = False
        return
    for l in model.layers:
        if l.name not in BRANCH_NAMES:
            continue
        if freeze_all or freeze_branches:
            l.trainable = False
        else:
            for sub in l.layers:
                if sub.kind == "conv":
                    sub.trainable = False


class MatMul:
    """Stand-in for the Keras layer of the GaitSet branch (:23-48): out[n] = x[n] . kernel[n] over the 2 * bin_num = 62
    parts, kernel [62, 128, hidden_dim].  The arithmetic lives in the engine (ugn_bmm_f32, tensor "<branch>/matmul/w");
    this class keeps the constructor / get_config protocol (it is named in custom_objects dictionaries, :58) and can
    multiply host arrays for inspection."""

    def __init__(self, bin_num=31, hidden_dim=256, **kwargs):
        self.bin_num, self.hidden_dim = bin_num, hidden_dim
        self.name = kwargs.get("name", "mat_mul")
        self.kernel = None                      # set_kernel(): e.g. model.get_layer("ofBranch").layers[-1].get_weights()[0]

    def set_kernel(self, kernel):
        k = np.asarray(kernel, dtype=np.float32)
        if k.shape != (2 * self.bin_num, 128, self.hidden_dim):
            raise ValueError(f"MatMul kernel must be {(2 * self.bin_num, 128, self.hidden_dim)}, got {k.shape}")
        self.kernel = k
        return self

    def call(self, x):
        if self.kernel is None:
            raise ValueError("MatMul: no kernel set (set_kernel)")
        return np.matmul(np.asarray(x, dtype=np.float32), self.kernel)

    __call__ = call

    def get_config(self):
        return {"name": self.name, "bin_num": self.bin_num, "hidden_dim": self.hidden_dim}


def fc_loadBranch(init_branch):
    """:57-62 -- the branch of a saved model as a stand-alone branch model: the saved file names its builder arguments
    (model.save of this package); the first branch of that graph is rebuilt alone and initialised from the file."""
    saved = _loadnet(init_branch)
    bname = BRANCH_NAMES[0]
    if saved.gaitset:
        import dataclasses
        cfg = dataclasses.replace(saved.cfg, in_channels=saved.cfg.in_channels[:1], nc=0, nclasses=0, single=True, postriplet=1)
        branch = UGaitModel(cfg, None, triplet_loss(margin=cfg.margin), 1.0, multimodal=False)
    else:
        c = saved.cfg
        act = "relu" if c.act == ACT_RELU else "leaky"
        if c.is3d(0):
            branch = UWYHSemiNet._branch3d(bname, (c.in_channels[0], c.hw, c.hw, 1), c.nd, "", act, c.alpha)
        else:
            branch = UWYHNet.buildBranch(bname, (c.in_channels[0], c.hw, c.hw), len(c.filters_numbers),
                                         [(k, k) for k in c.filters_size], list(c.filters_numbers), c.nd, c.weight_decay,
                                         c.dropout, None, _activation=act, alpha=c.alpha)
    P = saved.engine.export_params()
    branch.engine.load_params({k: v for k, v in P.items() if k.startswith(bname + "/")})
    branch.name = bname
    return branch


def mj_buildnet_by_config(netconfig, buildfun):
    """:299-330 -- rebuild a network from the dictionary the mains store next to their checkpoints, through ANY of the
    builders (`buildfun(input_shape, nlayers, filters_size, filters_numbers, ndense_units, weight_decay, dropout, ...)`)."""
    g = netconfig.get
    fn = netconfig["filters_numbers"]
    return buildfun(netconfig["input_shape"], len(fn), netconfig["filters_size"], fn, netconfig["ndense_units"],
                    netconfig["weight_decay"], netconfig["dropout"], nclasses=g("nclasses", 150),
                    loss_weights=g("loss_weights", [1.0, 0.1]), optimizer=netconfig["optimizer"],
                    margin=netconfig["margin"], use3D=g("use3D", False))


class UWYHNet:
    @staticmethod
    def buildBranch(name, input_shape=(50, 60, 60), number_convolutional_layers=4, filters_size=None,
                    filters_numbers=None, ndense_units=512, weight_decay=1e-4, dropout=0.4, init_branch=None,
                    _activation='relu', alpha=0.3):
        """:67-107 -- ONE stand-alone branch (the Keras Sequential conv stack -> ofFlat -> dense [-> drop] -> ofCode):
        a single-modality model without heads whose predict() is the branch output [B, ndense_units].  init_branch:
        a saved branch / model file to initialise from (fc_loadBranch, :57-66)."""
        if filters_size is None:
            filters_size = [(7, 7), (5, 5), (3, 3), (2, 2)]
        cfg = _cfg_from_args(tuple(input_shape), number_convolutional_layers, filters_size, filters_numbers, ndense_units,
                             weight_decay, dropout, 0.2, 0, [1.0, 1.0], Maximum, _activation, alpha, single=True)
        model = UGaitModel(cfg, None, triplet_loss(margin=0.2), 1.0, multimodal=False)
        model.name = name
        if init_branch:
            _load_branch(model, init_branch, BRANCH_NAMES[0])
        return model

    @staticmethod
    def buildBranchLReLU(name, input_shape=(50, 60, 60), number_convolutional_layers=4, filters_size=None,
                         filters_numbers=None, ndense_units=512, weight_decay=1e-4, dropout=0.4, init_branch=None,
                         alpha=0.3):
        """:110-152 -- the same stack with LeakyReLU(alpha) after every convolution."""
        return UWYHNet.buildBranch(name, input_shape, number_convolutional_layers, filters_size, filters_numbers,
                                   ndense_units, weight_decay, dropout, init_branch, _activation='leaky', alpha=alpha)

    @staticmethod
    def build(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units=512,
              weight_decay=1e-4, dropout=0.4, optimizer=None, margin=0.2, init_branches=None, freeze_branches=False,
              activation_fn='relu', alpha=0.3):
        """:154-245 -- the Siamese PAIR network: two weight-sharing (of, gray) towers, gated, Maximum-merged and
        l2-normalised, joined by VerifLossLayer(margin) on the pair label; the model's output IS the loss
        (model.compile(optimizer) only).  Inputs [ofinput1, ofuse1, grayinput1, grayuse1, ofinput2, ofuse2, grayinput2,
        grayuse2, label].  Here: both sides run as ONE 2B-row batch through the 2-modality engine
        (NetConfig.pair_loss -> ugn_pair_verif_loss)."""
        import dataclasses
        cfg = _cfg_from_args(list(input_shapes)[:2], number_convolutional_layers, filters_size, filters_numbers,
                             ndense_units, weight_decay, dropout, margin, 0, [1.0, 1.0], Maximum, activation_fn, alpha,
                             single=False)
        cfg = dataclasses.replace(cfg, pair_loss=True, nc=0)
        model = PairModel(cfg, optimizer)
        _apply_init_branches(model, init_branches)
        _freeze(model, freeze_branches=freeze_branches)
        return model

    @staticmethod
    def fit_generator(model, epochs, callbacks, training_generator, validation_generator, current_step, steps_per_epoch,
                      validation_steps, nworkers=0, new_lr=None):
        """:248-272 -- the same driver as UWYHSemiNet.fit_generator."""
        return UWYHSemiNet.fit_generator(model, epochs, callbacks, training_generator, validation_generator, current_step,
                                         steps_per_epoch, validation_steps, nworkers, new_lr)

    @staticmethod
    def encode(model, batch_data, use_data):
        """:275-295 -- gated, Maximum-merged, l2-normalised (of, gray) codes of a batch as numpy [B, nd]."""
        return UWYHSemiNet.encode(model, batch_data, use_data)


class UWYHSemiNet:
    def __init__(self):
        self.model = None

    @staticmethod
    def _branch3d(name, input_shape, ndense_units, init_branch, activation, alpha):
        cfg = _cfg_from_args(tuple(input_shape), 4, [(7, 7), (5, 5), (3, 3), (2, 2)], None, ndense_units, 0.0, 0.0, 0.2, 0,
                             [1.0, 1.0], Maximum, activation, alpha, single=True, use3D=True)
        model = UGaitModel(cfg, None, triplet_loss(margin=0.2), 1.0, multimodal=False)
        model.name = name
        if init_branch:
            _load_branch(model, init_branch, BRANCH_NAMES[0])
        return model

    @staticmethod
    def build_3Dbranch(name, input_shape=(25, 60, 60, 1), ndense_units=512, init_branch=""):
        """:336-372 -- ONE stand-alone Conv3D branch (six strided 'valid' Conv3D + ReLU, Conv3D(ndense_units, 1x1x1)
        "grayCode", Flatten) as a single-modality model whose predict() is the branch output [B, ndense_units]."""
        return UWYHSemiNet._branch3d(name, input_shape, ndense_units, init_branch, 'relu', 0.3)

    @staticmethod
    def build_3DbranchLReLU(name, input_shape=(25, 60, 60, 1), ndense_units=512, init_branch="", alpha=0.3):
        """:375-417 -- the same stack with LeakyReLU(alpha) after every convolution."""
        return UWYHSemiNet._branch3d(name, input_shape, ndense_units, init_branch, 'leaky', alpha)

    @staticmethod
    def build_gaitset_branch(name, input_layer, input_shape=(25, 60, 60, 1), ndense_units=512, init_branch="", norm=True):
        """:420-484 -- ONE stand-alone GaitSet branch: predict() is its [62, B, 256] output.  The reference applies the
        branch to `input_layer` (a Keras tensor) and returns the output tensor; here the handle is the model itself
        (`input_layer`, `ndense_units` and `norm` do not shape the branch in the reference either, :427-482)."""
        cfg = GaitSetConfig(in_channels=(int(input_shape[3]),), frames=int(input_shape[0]), hw=int(input_shape[1]), nc=0,
                            nclasses=0, single=True)
        model = UGaitModel(cfg, None, triplet_loss(margin=cfg.margin), 1.0, multimodal=False)
        model.name = name
        if init_branch:
            _load_branch(model, init_branch, BRANCH_NAMES[0])
        return model

    @staticmethod
    def get_weights_filename(modelpath):
        """:537-545 -- <dir>/<basename without extension>_weights.hdf5"""
        return osp.join(osp.dirname(modelpath), osp.splitext(osp.basename(modelpath))[0] + "_weights.hdf5")

    @staticmethod
    def get_netconfig_filename(modelpath):
        return osp.join(osp.dirname(modelpath), "model-config.hdf5")

    @staticmethod
    def build(input_shapes, number_convolutional_layers, filters_size, filters_numbers,
              ndense_units=512, weight_decay=1e-4, dropout=0.4, optimizer=None, margin=0.2,
              nclasses=0, loss_weights=[1.0, 1.0], use3D=False, smoothlabels=0, postriplet=1, init_branches=None,
              freeze_branches=False, aux_losses=False, fMerge=Maximum, fActivation='relu', alpha=0.3, gaitset=False):
        single = not isinstance(input_shapes, list)
        # use3D + gaitset: the 1-modality builder takes the GaitSet branch whatever use3D says (:776-784); with two
        # modalities the graph would fuse a [62,B,256] GaitSet output with a [B,nd] Conv3D output (:764-773) -- ill-formed
        _unsupported(use3D_with_gaitset=(use3D and gaitset and not single), aux_losses_with_gaitset=(aux_losses and gaitset))
        kwargs = dict(input_shapes=input_shapes, number_convolutional_layers=number_convolutional_layers,
                      filters_size=filters_size, filters_numbers=filters_numbers, ndense_units=ndense_units,
                      weight_decay=weight_decay, dropout=dropout, optimizer=optimizer, margin=margin, nclasses=nclasses,
                      loss_weights=loss_weights, smoothlabels=smoothlabels, postriplet=postriplet, aux_losses=aux_losses,
                      fMerge=fMerge, fActivation=fActivation, alpha=alpha, gaitset=gaitset, use3D=use3D)
        losses = [triplet_loss(margin=margin), 'categorical_crossentropy'] if nclasses > 0 else triplet_loss(margin=margin)
        if gaitset:
            cfg = _gs_cfg_from_args(input_shapes, ndense_units, dropout, margin, nclasses, loss_weights, fMerge,
                                    fActivation, alpha, smoothlabels, single=single, postriplet=postriplet)
            model = UGaitModel(cfg, optimizer, losses, loss_weights if nclasses > 0 else 1.0, multimodal=not single)
        else:
            cfg = _cfg_from_args(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units,
                                 weight_decay, dropout, margin, nclasses, loss_weights, fMerge, fActivation, alpha, single,
                                 smoothlabels=smoothlabels, aux_losses=aux_losses, postriplet=postriplet, use3D=use3D)
            model = UGaitModel(cfg, optimizer, losses, loss_weights if nclasses > 0 else 1.0, multimodal=not single)
        _apply_init_branches(model, init_branches)
        _freeze(model, freeze_branches=freeze_branches)
        return _remember(model, "UWYHSemiNet", **kwargs)

    @staticmethod
    def build_by_config(netconfig):
        """:487-534 -- rebuild from the dict the mains store next to their checkpoints (model-config.hdf5)."""
        g = netconfig.get
        fn = netconfig["filters_numbers"]
        return UWYHSemiNet.build(netconfig["input_shape"], len(fn), netconfig["filters_size"], fn,
                                 netconfig["ndense_units"], netconfig["weight_decay"], netconfig["dropout"],
                                 nclasses=g("nclasses", 150), loss_weights=g("loss_weights", [1.0, 0.1]),
                                 optimizer=netconfig.get("optimizer"), margin=netconfig["margin"], use3D=g("use3D", False),
                                 postriplet=g("postriplet", 1), fMerge=g("fMerge", Maximum),
                                 fActivation=g("fActivation", "relu"))

    @staticmethod
    def build_or_load(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units=512,
                      weight_decay=1e-4, dropout=0.4, optimizer=None, margin=0.2, nclasses=0,
                      loss_weights=[1.0, 1.0], initnet="", freeze_convs=False, use3D=False, smoothlabels=0,
                      freeze_all=False, postriplet=1, init_branches=None, freeze_branches=False, aux_losses=False,
                      fMerge=Maximum, fActivation='relu', gaitset=False):
        # (the reference's 2-modality build_or_load has no `alpha` argument, :582-588: LeakyReLU keeps build()'s 0.3)
        if gaitset:
            fActivation = 'leaky'          # :588-589
        model = UWYHSemiNet.build(input_shapes, number_convolutional_layers, filters_size, filters_numbers,
                                  ndense_units, weight_decay, dropout, optimizer, margin, nclasses, loss_weights,
                                  use3D=use3D, smoothlabels=smoothlabels, postriplet=postriplet,
                                  init_branches=init_branches, freeze_branches=freeze_branches, aux_losses=aux_losses,
                                  fMerge=fMerge, fActivation=fActivation, gaitset=gaitset)
        if initnet != "":
            # :598-664 -- the reference loads the saved model and, when the classifier width differs ("surgery"),
            # rebuilds and loads the compatible weights by name; both paths end in exactly this state
            model.load_weights(_weights_file_of(initnet), by_name=True, skip_mismatch=True)
            _freeze(model, freeze_convs=freeze_convs, freeze_all=freeze_all)
        return model

    @staticmethod
    def loadnet(netpath: str):
        """:554-579 / :1008-1029 -- load_model(netpath, compile=False): rebuild the graph from the builder arguments
        stored in the file (this wrapper's model.save) or, for a checkpoint directory written by the reference's mains,
        from `ugn_builder_config.json` next to it, then load the weights."""
        return _loadnet(netpath)

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
        # (postriplet is accepted and ignored by the reference's 3-modality graph: "TODO implement use of 'postriplet'", :1054)
        _unsupported(use3D_with_gaitset=(use3D and gaitset), aux_losses_with_gaitset=(aux_losses and gaitset))
        kwargs = dict(input_shapes=list(input_shapes), number_convolutional_layers=number_convolutional_layers,
                      filters_size=filters_size, filters_numbers=filters_numbers, ndense_units=ndense_units,
                      weight_decay=weight_decay, dropout=dropout, optimizer=optimizer, margin=margin, nclasses=nclasses,
                      loss_weights=loss_weights, smoothlabels=smoothlabels, aux_losses=aux_losses, fMerge=fMerge,
                      normbfmerge=normbfmerge, fActivation=fActivation, alpha=alpha, gaitset=gaitset, use3D=use3D)
        losses = [triplet_loss(margin=margin), 'categorical_crossentropy'] if nclasses > 0 else triplet_loss(margin=margin)
        if gaitset:
            cfg = _gs_cfg_from_args(input_shapes, ndense_units, dropout, margin, nclasses, loss_weights, fMerge,
                                    fActivation, alpha, smoothlabels)       # (normbfmerge is ignored with gaitset, :1164)
        else:
            cfg = _cfg_from_args(list(input_shapes), number_convolutional_layers, filters_size, filters_numbers,
                                 ndense_units, weight_decay, dropout, margin, nclasses, loss_weights, fMerge, fActivation,
                                 alpha, single=False, smoothlabels=smoothlabels, normbfmerge=normbfmerge,
                                 aux_losses=aux_losses, use3D=use3D)
        model = UGaitModel(cfg, optimizer, losses, loss_weights if nclasses > 0 else 1.0, multimodal=True)
        _apply_init_branches(model, init_branches)
        _freeze(model, freeze_branches=freeze_branches)
        return _remember(model, "UWYHSemiNet3Mods", **kwargs)

    @staticmethod
    def loadnet(netpath: str):
        """:1008-1029 -- as UWYHSemiNet.loadnet (the file names the builder that made it)."""
        return _loadnet(netpath)

    @staticmethod
    def compile_hard(model, optimizer, loss_weights, margin):
        """:1302-1306 -- re-compile with `tfa.losses.TripletHardLoss(margin)` (batch-hard: farthest positive, nearest
        negative per anchor; CUDA: ugn_triplet_hard) in place of the batch-all loss; the CE loss, the weights of the net
        and the metrics stay, the optimiser starts afresh as under model.compile."""
        if model.gaitset:
            raise NotImplementedError("tfa.losses.TripletHardLoss takes rank-2 embeddings; the GaitSet signature is "
                                      "[62, B, 256] (the reference would fail inside tfa as well)")
        opt = optimizer if optimizer is not None else optimizers.SGD(0.001, 0.9)
        kw = dict(getattr(opt, "kw", {}))
        lw = list(loss_weights) if loss_weights is not None else [1.0, 1.0]
        model.engine.recompile(optimizer=getattr(opt, "name", "sgd"), lr=getattr(opt, "lr", 0.001), margin=float(margin),
                               wver=float(lw[0]), wid=float(lw[1]) if len(lw) > 1 else None, triplet_hard=True,
                               momentum=kw.get("momentum"), beta1=kw.get("beta1"), beta2=kw.get("beta2"),
                               eps=kw.get("eps"), lr_decay=kw.get("decay"),
                               decoupled_weight_decay=kw.get("weight_decay"))
        model.cfg = model.engine.cfg
        model.loss = [TripletHardLoss(margin=margin)] + list(model.loss[1:])
        model.loss_weights = lw
        return model

    @staticmethod
    def build_or_load(input_shapes, number_convolutional_layers, filters_size, filters_numbers, ndense_units=512,
                      weight_decay=1e-4, dropout=0.4, optimizer=None, margin=0.2, nclasses=0,
                      loss_weights=[1.0, 1.0], initnet="", freeze_convs=False, use3D=False, smoothlabels=0,
                      freeze_all=False, postriplet=1, init_branches=None, freeze_branches=False, aux_losses=False,
                      fMerge=Maximum, normbfmerge=False, fActivation='relu', alpha=0.3, gaitset=False):
        if gaitset:
            fActivation = 'leaky'
        model = UWYHSemiNet3Mods.build(input_shapes, number_convolutional_layers, filters_size, filters_numbers,
                                       ndense_units, weight_decay, dropout, optimizer, margin, nclasses,
                                       loss_weights, use3D=use3D, smoothlabels=smoothlabels, postriplet=postriplet,
                                       init_branches=init_branches, freeze_branches=freeze_branches,
                                       aux_losses=aux_losses, fMerge=fMerge, normbfmerge=normbfmerge,
                                       fActivation=fActivation, alpha=alpha, gaitset=gaitset)
        if initnet != "":
            # :1325-1391 -- load (with classifier "surgery" = by-name loading that skips the mismatching classprob), then
            # freeze the convolutions / every branch layer
            model.load_weights(_weights_file_of(initnet), by_name=True, skip_mismatch=True)
            _freeze(model, freeze_convs=freeze_convs, freeze_all=freeze_all)
        return model


def _weights_file_of(initnet: str) -> str:
    """initnet names the full-model file; the mains keep the weights next to it as <name>_weights.hdf5
    (:536-544).  Fall back to the model file itself (its /model_weights group) when only that exists."""
    wf = UWYHSemiNet.get_weights_filename(initnet)
    return wf if osp.exists(wf) else initnet


def _loadnet(netpath: str):
    import json
    from ugaitnet_b200 import hdf5
    conf = None
    if osp.exists(netpath) and hdf5.is_hdf5(netpath):
        f = hdf5.File(netpath)
        raw = f.attrs.get("ugn_builder_config")
        if raw is not None:
            conf = json.loads(raw.decode("utf8") if isinstance(raw, bytes) else raw)
    if conf is None:
        side = osp.join(osp.dirname(netpath), "ugn_builder_config.json")
        if osp.exists(side):
            with open(side) as fh:
                conf = json.load(fh)
    if conf is None or "kwargs" not in conf:
        raise ValueError(f"{netpath}: no builder configuration found (files written by tf.keras carry a Keras "
                         "model_config; rebuild the graph with build() / build_by_config() and call "
                         "model.load_weights(path, by_name=True), which reads Keras HDF5 weight files)")
    cls = {"UWYHSemiNet": UWYHSemiNet, "UWYHSemiNet3Mods": UWYHSemiNet3Mods}[conf["class"]]
    kw = {k: _unjson(v) for k, v in conf["kwargs"].items()}
    shapes = kw.pop("input_shapes")
    shapes = [tuple(s) for s in shapes] if isinstance(shapes[0], (list, tuple)) else tuple(shapes)
    kw["filters_size"] = [tuple(k) if isinstance(k, list) else k for k in kw["filters_size"]]
    model = cls.build(shapes, kw.pop("number_convolutional_layers"), kw.pop("filters_size"), kw.pop("filters_numbers"), **kw)
    model.load_weights(netpath if "model_weights" in hdf5.File(netpath).keys() else _weights_file_of(netpath), by_name=True)
    return model
