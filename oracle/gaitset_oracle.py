"""CPU oracle for the GaitSet branch type of UGaitNet.  TEST INFRASTRUCTURE ONLY.

PyTorch-CPU restatement (fp64 "truth" / fp32 "TF-like") of
``UWYHSemiNet.build_gaitset_branch`` + ``MatMul`` (/root/reference/nets/mj_uwyhNets_ba.py:420-484, :23-48)
and of the 3-modality graph built around it with ``gaitset=True`` (:1110-1214): gate, fusion,
``l2_normalize(axis=1)`` -- which in this layout ``[62, B, 256]`` normalises over the BATCH axis, kept
literally --, FC1 "code", transpose + Flatten + FC2 "classprob", batch-all triplet over the 62 parts.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it.

PARITY STATUS: **parity unpinned** -- the arithmetic runs inside TensorFlow 2.3 / Keras in the reference,
TensorFlow cannot be installed here and the reference holds no test or golden vector for this path.  The
restatement follows the call sites cited per function and documented Keras/TF semantics
(``padding='same'`` zero padding, ``LeakyReLU()`` alpha 0.3, ``reduce_max`` gradient split evenly among
ties, ``MaxPooling2D`` gradient to the first maximum, GlorotUniform fans of a rank-3 kernel,
activity regulariser divided by ``shape(output)[0]``).  The ARITHMETIC of the restated graph is pinned by a second,
independent restatement (tests/test_oracle.py: layer-by-layer channels-last numpy, forward to 1e-12, gradients against
central differences of that forward); what stays unpinned is whether the graph is the one TF 2.3 builds.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from .ugait_oracle import BRANCH_NAMES, MERGE_SIGNMAX, l2_normalize, merge_modalities, softmax_ce, triplet_loss_all

# (name, cin, cout, k) of build_gaitset_branch, in graph order; cin None = per-frame input channels
GS_CONVS = (("a1", None, 32, 5), ("a2", 32, 32, 3), ("b1", 32, 64, 3), ("b2", 64, 64, 3), ("a3", 32, 64, 3),
            ("a4", 64, 64, 3), ("b3", 64, 128, 3), ("b4", 128, 128, 3), ("a5", 64, 128, 3), ("a6", 128, 128, 3))
HPP_BINS = (1, 2, 4, 8, 16)
N_PARTS = 2 * sum(HPP_BINS)          # 62
GS_ALPHA = 0.3                       # layers.LeakyReLU() default


@dataclass
class GaitSetConfig:
    """Builder arguments that shape the gaitset=True graph (nets/mj_uwyhNets_ba.py:1032-1037)."""
    in_channels: Sequence[int] = (2, 1, 1)      # per-frame channels: OF (x,y), gray, depth
    frames: int = 25
    hw: int = 60
    hidden: int = 256                            # MatMul hidden_dim (:24)
    nc: int = 0                                  # ndense_units[1] (FC1 "code"), 0 = absent
    nclasses: int = 150
    merge: int = MERGE_SIGNMAX
    alpha: float = 0.3                           # LeakyReLU after "code" (fActivation != 'relu' is required)
    margin: float = 0.2
    wver: float = 1.0
    wid: float = 0.1
    label_smoothing: float = 0.0                 # smoothlabels (:1252-1262)
    postriplet: int = 1                          # 2 (with nc > 0; 2-modality builder :814-832): the fusion is NOT normalised,
    #                                              FC1 is the layer "signature", its l2_normalize(axis=1) "code" is the
    #                                              embedding the triplet loss and the classifier see
    single: bool = False                         # UWYHSemiNet.build on ONE input shape (:890-905): the branch output IS
    #                                              the signature -- no use-flag gate, no fusion, no l2_normalize, no FC1

    @property
    def nmods(self):
        return len(self.in_channels)


def init_params(cfg: GaitSetConfig, seed: int = 232323, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Keras defaults: glorot_uniform conv kernels (no bias, :428-466), MatMul GlorotUniform on the rank-3
    shape (62,128,256) (:33-34: receptive field 62 -> fan_in 128*62, fan_out 256*62), glorot Dense."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, limit):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * limit).to(dtype)

    P = {}
    for m in range(cfg.nmods):
        bn = BRANCH_NAMES[m]
        for name, cin, co, k in GS_CONVS:
            cin = cfg.in_channels[m] if cin is None else cin
            P[f"{bn}/{name}/w"] = uni((co, cin, k, k), math.sqrt(6.0 / (cin * k * k + co * k * k)))
        P[f"{bn}/matmul/w"] = uni((N_PARTS, 128, cfg.hidden), math.sqrt(6.0 / (N_PARTS * (128 + cfg.hidden))))
    feat = cfg.hidden
    if cfg.nc > 0:
        P["code/w"] = uni((cfg.nc, cfg.hidden), math.sqrt(6.0 / (cfg.nc + cfg.hidden)))
        P["code/b"] = torch.zeros(cfg.nc, dtype=dtype)
        feat = cfg.nc
    if cfg.nclasses > 0:
        P["classprob/w"] = uni((cfg.nclasses, N_PARTS * feat), math.sqrt(6.0 / (cfg.nclasses + N_PARTS * feat)))
        P["classprob/b"] = torch.zeros(cfg.nclasses, dtype=dtype)
    return P


def _conv(x, w, pool=False):
    """Conv2D(padding='same', use_bias=False) + LeakyReLU() [+ MaxPooling2D 2x2] (:428-433)."""
    y = F.leaky_relu(F.conv2d(x, w, padding=w.shape[-1] // 2), GS_ALPHA)
    return F.max_pool2d(y, 2) if pool else y


def _hpp(x):
    """:468-479 -- Reshape((num_bin, -1, c)) of the NHWC map: consecutive (h,w) positions in row-major
    order are grouped into num_bin strips; mean + max per strip.  x NCHW [B,C,H,W] -> list of [B,nb,C]."""
    B, C = x.shape[:2]
    flat = x.permute(0, 2, 3, 1).reshape(B, -1, C)
    return [flat.reshape(B, nb, -1, C).mean(2) + flat.reshape(B, nb, -1, C).amax(2) for nb in HPP_BINS]


def gaitset_branch_forward(x, P, bn, cfg: GaitSetConfig, return_acts: bool = False):
    """build_gaitset_branch (:420-484).  x [B,T,H,W,c] (the Keras input layout) -> [62,B,hidden]."""
    B, T, H, W, c = x.shape
    acts = {}
    f = x.permute(0, 1, 4, 2, 3).reshape(B * T, c, H, W)
    f = F.pad(f, (2, 2, 2, 2))                                      # TimeDistributed(ZeroPadding2D(2))

    def setmax(a):                                                  # Lambda reduce_max(axis=1)
        return a.reshape(B, T, *a.shape[1:]).amax(1)

    a = _conv(f, P[f"{bn}/a1/w"])
    a = _conv(a, P[f"{bn}/a2/w"], pool=True)
    acts["a2"] = a
    b = setmax(a)
    b = _conv(b, P[f"{bn}/b1/w"])
    b = _conv(b, P[f"{bn}/b2/w"], pool=True)
    a = _conv(a, P[f"{bn}/a3/w"])
    a = _conv(a, P[f"{bn}/a4/w"], pool=True)
    acts["a4"] = a
    b = b + setmax(a)
    b = _conv(b, P[f"{bn}/b3/w"])
    b = _conv(b, P[f"{bn}/b4/w"])
    a = _conv(a, P[f"{bn}/a5/w"])
    a = _conv(a, P[f"{bn}/a6/w"])
    a = setmax(a)
    b = b + a
    acts["a_set"], acts["b_set"] = a, b
    feats = []
    for fa, fb in zip(_hpp(a), _hpp(b)):
        feats += [fa, fb]
    feat = torch.cat(feats, 1).permute(1, 0, 2)                     # [62,B,128]
    acts["hpp"] = feat
    out = torch.matmul(feat, P[f"{bn}/matmul/w"])                   # MatMul (:41-46)
    return (out, acts) if return_acts else out


def model_forward(inputs, flags, P, cfg: GaitSetConfig, return_all: bool = False):
    """UWYHSemiNet3Mods.build with gaitset=True (:1163-1214); cfg.single: the 1-modality graph of UWYHSemiNet.build
    (:890-905: ``ofout1 = ofBranch; outsignature = ofout1``, then transpose + Flatten + "classprob")."""
    outs = {}
    if cfg.single:
        assert cfg.nmods == 1 and cfg.nc == 0
        sig = gaitset_branch_forward(inputs[0], P, BRANCH_NAMES[0], cfg)
        outs["branch0"] = outs["signature"] = sig
        if cfg.nclasses > 0:
            outs["logits"] = F.linear(sig.permute(1, 0, 2).flatten(1), P["classprob/w"], P["classprob/b"])
        return outs if return_all else (outs["signature"], outs.get("logits"))
    gated = []
    for m in range(cfg.nmods):
        b = gaitset_branch_forward(inputs[m], P, BRANCH_NAMES[m], cfg)
        outs[f"branch{m}"] = b
        gated.append(b * flags[m].reshape(1, -1, 1))                # mj_tensor_times_scalar broadcast (:51-54)
    fused = merge_modalities(gated, cfg.merge)
    outs["fusion"] = fused
    if cfg.postriplet == 2 and cfg.nc > 0:                          # :819-832 (LeakyReLU path :825-827)
        lin = F.linear(fused, P["code/w"], P["code/b"])             # Dense(nc, activation=None, activity_regularizer, name="signature")
        outs["code_lin"] = outs["signature_layer"] = lin
        sig = l2_normalize(F.leaky_relu(lin, cfg.alpha), 1)         # Lambda(l2_normalize(axis=1), name="code") = outsignature
        outs["signature"] = outs["code"] = sig
        if cfg.nclasses > 0:                                        # Dropout("dropcode") -> transpose + Flatten -> "classprob"
            outs["logits"] = F.linear(sig.permute(1, 0, 2).flatten(1), P["classprob/w"], P["classprob/b"])
        return outs if return_all else (outs["signature"], outs.get("logits"))
    sig = l2_normalize(fused, 1)                                    # axis=1 == the batch axis here (:1191)
    outs["signature"] = sig
    feat = sig
    if cfg.nc > 0:
        lin = F.linear(sig, P["code/w"], P["code/b"])               # Dense(activation=None, activity_regularizer)
        outs["code_lin"] = lin
        feat = F.leaky_relu(lin, cfg.alpha)                         # (:1199-1201)
        outs["code"] = feat
    if cfg.nclasses > 0:
        flat = feat.permute(1, 0, 2).flatten(1)                     # transpose [1,0,2] + Flatten (:1211-1213)
        outs["logits"] = F.linear(flat, P["classprob/w"], P["classprob/b"])
    return outs if return_all else (outs["signature"], outs.get("logits"))


def total_loss(inputs, flags, labels, P, cfg: GaitSetConfig):
    outs = model_forward(inputs, flags, P, cfg, return_all=True)
    res = {}
    trip, cnt = triplet_loss_all(labels, outs["signature"], cfg.margin)
    res["triplet"], res["count"] = trip, cnt
    loss = cfg.wver * trip
    if cfg.nclasses > 0:
        onehot = F.one_hot(labels.reshape(-1).long(), cfg.nclasses).to(trip.dtype)
        ce, acc = softmax_ce(outs["logits"], onehot)
        if cfg.label_smoothing > 0:
            ce, _ = softmax_ce(outs["logits"], onehot * (1.0 - cfg.label_smoothing) + cfg.label_smoothing / cfg.nclasses)
        res["ce"], res["acc"] = ce, acc
        loss = loss + cfg.wid * ce
    reg = torch.zeros((), dtype=trip.dtype)
    if cfg.nc > 0:
        # activity regulariser of "code": Keras divides by shape(output)[0], which is 62 in this layout
        reg = reg + 1e-3 * (outs["code_lin"] ** 2).sum() / outs["code_lin"].shape[0]
    res["reg"] = reg
    res["loss"] = loss + reg
    res["signature"], res["logits"] = outs["signature"], outs.get("logits")
    return res


def loss_and_grads(inputs, flags, labels, P, cfg):
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    res = total_loss(inputs, flags, labels, Pg, cfg)
    res["loss"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in res.items()}, grads


def synth_batch(cfg: GaitSetConfig, ids: int, per_id: int, seed: int = 232323, dtype=torch.float32,
                missing: bool = True):
    """Synthetic gaitset batch: [B,T,hw,hw,c] per modality (value ranges of SURVEY 8d), flags [B,1] with the
    never-all-missing rule, labels [B] (ids x per_id)."""
    g = torch.Generator().manual_seed(seed)
    B = ids * per_id
    xs = []
    for c in cfg.in_channels:
        if c == 2:
            x = (torch.randn(B, cfg.frames, cfg.hw, cfg.hw, c, generator=g, dtype=torch.float64) * 0.3).clamp(-3.3, 3.3)
        else:
            x = torch.rand(B, cfg.frames, cfg.hw, cfg.hw, c, generator=g, dtype=torch.float64) - 0.5
        xs.append(x.to(dtype))
    flags = torch.ones(cfg.nmods, B, 1, dtype=dtype)
    if missing and cfg.nmods > 1:
        for b in range(B):
            if b % 3 == 1:
                flags[int(torch.randint(cfg.nmods, (1,), generator=g)), b, 0] = 0.0
            elif b % 3 == 2:
                keep = int(torch.randint(cfg.nmods, (1,), generator=g))
                for m in range(cfg.nmods):
                    flags[m, b, 0] = 1.0 if m == keep else 0.0
    for m in range(cfg.nmods):
        off = flags[m].reshape(B) == 0
        xs[m][off] = 1e-9                                           # the generator's `noise` fill
    labels = (torch.arange(ids) * 2 + 1).repeat_interleave(per_id)
    assert cfg.nclasses == 0 or int(labels.max()) < cfg.nclasses
    return xs, [flags[m] for m in range(cfg.nmods)], labels
