"""CPU oracle for the UGaitNet hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (PyTorch-CPU / numpy, fp64 "truth" and fp32 "TF-like") of
the arithmetic the reference executes for one training / descriptor-extraction step and for
the open-world k-NN test.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``ugaitnet_b200``) never imports anything from ``oracle/``.

PARITY STATUS
-------------
* ``knn_predict`` is PINNED against the real reference call
  (``sklearn.neighbors.KNeighborsClassifier(n_neighbors=k).fit(G, y).predict(Q)``,
  /root/reference/mains/mj_testUWYHGaitNet_open_tum.py:331-341) -- sklearn is installed in
  the build container and the golden vectors in ``tests/golden/knn_*.npz`` were produced by
  it (``tests/golden/make_golden.py``).
* ``eer_verif_dist`` is PINNED against the reference's own module
  (/root/reference/nets/mj_metrics.py:10-24) and its demo known-answer (EER 0.25, thr 0.07).
* Everything that executes inside TensorFlow 2.3 / Keras in the reference (conv branches,
  fusion, losses, Adam) is **parity unpinned**: TensorFlow is not installable here (no
  network) and the reference ships no tests / golden vectors for the path.  Those functions
  restate the documented Keras/TF 2.3 semantics of the call sites cited next to each one.

All citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MERGE_MAX, MERGE_AVG, MERGE_SIGNMAX = 0, 1, 2
ACT_LINEAR, ACT_RELU, ACT_LEAKY = 0, 1, 2


# --------------------------------------------------------------------------------------
# configuration / parameters
# --------------------------------------------------------------------------------------
@dataclass
class NetConfig:
    """Mirror of the builder arguments (nets/mj_uwyhNets_ba.py:1032-1037, :669-672)."""
    in_channels: Sequence[int] = (50, 25, 25)          # one entry per modality
    filters_numbers: Sequence[int] = (96, 192, 512, 512)  # mains/..._3mods.py:232-237
    filters_size: Sequence[int] = (7, 5, 3, 2)
    nd: int = 2048                                      # ndense_units[0]
    nc: int = 0                                         # ndense_units[1] (FC1 "code"), 0 = absent
    nclasses: int = 150
    weight_decay: float = 5e-5                          # mains/..._3mods.py:239
    merge: int = MERGE_SIGNMAX
    act: int = ACT_RELU                                 # fActivation ('relu' | leaky)
    alpha: float = 0.3
    margin: float = 0.2
    wver: float = 1.0
    wid: float = 0.1
    hw: int = 60
    single: bool = False      # UWYHSemiNet.build with a non-list input_shapes (:900-915): no gate,
    #                           no fusion, NO l2_normalize on the signature.
    label_smoothing: float = 0.0   # smoothlabels: tf.losses.CategoricalCrossentropy(label_smoothing) (:1252-1262)
    normbfmerge: bool = False      # per-branch l2_normalize before the gate (:1167-1168)
    aux_losses: bool = False       # classprob_{of,gray,depth} heads on the gated branch outputs (:1222-1251)
    waux: float = 1.0              # loss_weights[-1] (:1264-1268)
    triplet_hard: bool = False     # compile_hard: tfa.losses.TripletHardLoss instead of the batch-all loss (:1302-1306)
    branch3d: tuple = ()           # use3D: per modality True -> Conv3D branch (build_3Dbranch{,LReLU}, :336-417) on [B,25,60,60,1]
    filters3d: tuple = (64, 128, 256, 512, 512, 512)   # its six activated Conv3D layers (geometry fixed, LAYERS3D below)
    pair_loss: bool = False        # UWYHNet.build (:154-245): rows [0,B) / [B,2B) are the two sides of B pairs, VerifLossLayer
    postriplet: int = 1            # 2: fusion -> Dense "signature" (activity-regularised) -> l2_normalize "code" = the
    #                                embedding of the triplet loss and of the classifier (:814-832, 2-modality builder)

    @property
    def nmods(self):
        return len(self.in_channels)

    def spatial(self):
        """[(H_in, H_conv, H_out_after_pool)] per conv layer; pool after all but the last
        (nets/mj_uwyhNets_ba.py:85-92)."""
        s, out = self.hw, []
        for i, k in enumerate(self.filters_size):
            c = s - k + 1
            p = c // 2 if i != len(self.filters_size) - 1 else c
            out.append((s, c, p))
            s = p
        return out

    @property
    def flat(self):
        return self.filters_numbers[-1] * self.spatial()[-1][2] ** 2


BRANCH_NAMES = ("ofBranch", "grayBranch", "depthBranch")
# build_3Dbranch (:346-363): (kernel (kt,kh,kw), strides) of the six activated Conv3D layers, 'valid', channels_last;
# then Conv3D(ndense_units, (1,1,1)) "grayCode" (linear, he_uniform, L2 1e-3) and Flatten.  [25,60,60,1] ends at 1x1x1.
LAYERS3D = (((3, 5, 5), (1, 2, 2)), ((3, 3, 3), (1, 2, 2)), ((3, 3, 3), (2, 2, 2)), ((3, 3, 3), (2, 2, 2)),
            ((3, 2, 2), (1, 1, 1)), ((2, 1, 1), (1, 1, 1)))


def is3d(cfg, m):
    return bool(len(getattr(cfg, "branch3d", ())) > m and cfg.branch3d[m])
AUX_NAMES = ("classprob_of", "classprob_gray", "classprob_depth")


def init_params(cfg: NetConfig, seed: int = 232323, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Keras default initialisers (glorot_uniform kernels, zero biases; ``ofCode`` is
    he_uniform -- nets/mj_uwyhNets_ba.py:82-105).  Layouts: conv ``[Cout,Cin,kh,kw]``,
    dense ``[out,in]`` (the transposes of Keras' (kh,kw,cin,cout) / (in,out))."""
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}

    def uni(shape, limit):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * limit).to(dtype)

    for m in range(cfg.nmods):
        bn = BRANCH_NAMES[m]
        cin = cfg.in_channels[m]
        if is3d(cfg, m):          # Keras defaults: glorot_uniform Conv3D kernels, zero biases; grayCode he_uniform
            cin = 1
            for li, (co, (k, _)) in enumerate(zip(cfg.filters3d, LAYERS3D)):
                rf = k[0] * k[1] * k[2]
                P[f"{bn}/conv{li}/w"] = uni((co, cin) + tuple(k), math.sqrt(6.0 / (cin * rf + co * rf)))
                P[f"{bn}/conv{li}/b"] = torch.zeros(co, dtype=dtype)
                cin = co
            P[f"{bn}/ofCode/w"] = uni((cfg.nd, cin), math.sqrt(6.0 / cin))
            P[f"{bn}/ofCode/b"] = torch.zeros(cfg.nd, dtype=dtype)
            continue
        for li, (co, k) in enumerate(zip(cfg.filters_numbers, cfg.filters_size)):
            fan_in, fan_out = cin * k * k, co * k * k
            P[f"{bn}/conv{li}/w"] = uni((co, cin, k, k), math.sqrt(6.0 / (fan_in + fan_out)))
            P[f"{bn}/conv{li}/b"] = torch.zeros(co, dtype=dtype)
            cin = co
        P[f"{bn}/dense/w"] = uni((2 * cfg.nd, cfg.flat), math.sqrt(6.0 / (cfg.flat + 2 * cfg.nd)))
        P[f"{bn}/dense/b"] = torch.zeros(2 * cfg.nd, dtype=dtype)
        P[f"{bn}/ofCode/w"] = uni((cfg.nd, 2 * cfg.nd), math.sqrt(6.0 / (2 * cfg.nd)))
        P[f"{bn}/ofCode/b"] = torch.zeros(cfg.nd, dtype=dtype)
    feat = cfg.nd
    if cfg.nc > 0:
        P["code/w"] = uni((cfg.nc, cfg.nd), math.sqrt(6.0 / (cfg.nd + cfg.nc)))
        P["code/b"] = torch.zeros(cfg.nc, dtype=dtype)
        feat = cfg.nc
    if cfg.nclasses > 0:
        P["classprob/w"] = uni((cfg.nclasses, feat), math.sqrt(6.0 / (feat + cfg.nclasses)))
        P["classprob/b"] = torch.zeros(cfg.nclasses, dtype=dtype)
    if cfg.nclasses > 0 and getattr(cfg, "aux_losses", False):
        for m in range(cfg.nmods):
            P[f"{AUX_NAMES[m]}/w"] = uni((cfg.nclasses, cfg.nd), math.sqrt(6.0 / (cfg.nd + cfg.nclasses)))
            P[f"{AUX_NAMES[m]}/b"] = torch.zeros(cfg.nclasses, dtype=dtype)
    return P


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------
def _act(x, act, alpha):
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_LEAKY:
        return F.leaky_relu(x, alpha)
    return x


def _windows2x2(h):
    """[B,C,H,W] -> [B,C,H//2,W//2,4], window position = dy*2+dx (floor pooling: odd last row/column dropped)."""
    B, C, H, W = h.shape
    Hp, Wp = H // 2, W // 2
    return h[:, :, :2 * Hp, :2 * Wp].reshape(B, C, Hp, 2, Wp, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, Hp, Wp, 4)


def branch3d_forward(x, P, bn, cfg: NetConfig):
    """UWYHSemiNet.build_3Dbranch / build_3DbranchLReLU (nets/mj_uwyhNets_ba.py:336-417): x [B,25,60,60,1] (the
    generator's np.expand_dims(x, 3), data/mj_dataGeneratorMMUWYHsingle.py:431-432) -> six strided 'valid' Conv3D with
    ReLU / LeakyReLU(alpha) -> Conv3D(nd, 1x1x1) "grayCode" (linear) -> Flatten."""
    h = x.reshape(x.shape[0], 1, x.shape[1], x.shape[2], x.shape[3])          # channels_last C = 1 -> NCDHW
    for li, (_, st) in enumerate(LAYERS3D):
        h = _act(F.conv3d(h, P[f"{bn}/conv{li}/w"], P[f"{bn}/conv{li}/b"], stride=st), cfg.act, cfg.alpha)
    assert tuple(h.shape[2:]) == (1, 1, 1), "build_3Dbranch expects the [25,60,60,1] volume"
    return F.linear(h.flatten(1), P[f"{bn}/ofCode/w"], P[f"{bn}/ofCode/b"])


def branch_forward(x, P, bn, cfg: NetConfig, drop_mask=None, return_acts=False, decisions=None, record=None):
    """UWYHNet.buildBranch / buildBranchLReLU (nets/mj_uwyhNets_ba.py:67-107, :110-152).
    x: [B,Cin,60,60] NCHW.  drop_mask: already-scaled inverted-dropout mask [B,2nd] or None.

    Discrete-decision bookkeeping for the gradient-parity tests (tests/test_decisions_gpu.py):
      record    -- dict that receives, per conv layer li, the decisions this run takes: ``pool{li}`` arg-max position
                   (dy*2+dx, first maximum wins) of every 2x2 window, ``gap{li}`` = best minus second best activated
                   value of the window, ``act{li}`` = (selected pre-activation > 0), ``mag{li}`` = |selected pre-activation|;
      decisions -- dict with ``pool{li}`` / ``act{li}`` taken by ANOTHER implementation: the forward pass then gathers
                   exactly those window positions and applies exactly those activation branches, so that autograd
                   routes the gradient the way that implementation did (the graph is linear once the decisions are fixed)."""
    acts = {}
    h = x
    nl = len(cfg.filters_numbers)
    for li in range(nl):
        z = F.conv2d(h, P[f"{bn}/conv{li}/w"], P[f"{bn}/conv{li}/b"])      # valid, stride 1
        pooled = li != nl - 1
        if decisions is None and record is None:
            h = _act(z, cfg.act, cfg.alpha)
            if pooled:
                h = F.max_pool2d(h, 2)                                     # floor pooling
        else:
            if pooled:
                win = _windows2x2(z)                                       # activation is monotonic: pool(act(z)) = act(pool(z))
                if decisions is not None:
                    idx = decisions[f"pool{li}"].long()
                else:
                    # first maximum of the window.  The activation is monotonic, so this is the arg-max of the activated
                    # window wherever its maximum is positive; an all-<=0 ReLU window is a 4-way tie of zeros that routes
                    # no gradient, and recording its largest pre-activation tells how far it is from switching on
                    idx = win.detach().argmax(dim=4)
                sel = torch.gather(win, 4, idx.unsqueeze(4)).squeeze(4)
                if record is not None:
                    a = _act(win.detach(), cfg.act, cfg.alpha)
                    top2 = a.topk(2, dim=4).values
                    record[f"pool{li}"] = idx
                    record[f"gap{li}"] = top2[..., 0] - top2[..., 1]
            else:
                sel = z
            mask = decisions[f"act{li}"].bool() if decisions is not None else (sel.detach() > 0)
            if record is not None:
                record[f"act{li}"] = mask
                record[f"mag{li}"] = sel.detach().abs()
            if cfg.act == ACT_RELU:
                h = torch.where(mask, sel, torch.zeros_like(sel))
            elif cfg.act == ACT_LEAKY:
                h = torch.where(mask, sel, cfg.alpha * sel)
            else:
                h = sel
        acts[f"conv{li}"] = h
    h = h.flatten(1)                                                       # (C,H,W) order
    h = F.linear(h, P[f"{bn}/dense/w"], P[f"{bn}/dense/b"])
    if drop_mask is not None:
        h = h * drop_mask
    acts["dense"] = h
    h = F.linear(h, P[f"{bn}/ofCode/w"], P[f"{bn}/ofCode/b"])
    acts["ofCode"] = h
    return (h, acts) if return_acts else h


def l2_normalize(x, axis=1, eps=1e-12):
    """tf.math.l2_normalize (nets/mj_uwyhNets_ba.py:1191): x * rsqrt(max(sum x^2, eps))."""
    ss = (x * x).sum(dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=eps))


def merge_modalities(gated: List[torch.Tensor], merge: int, winner=None, record=None):
    """fMerge(name="fusion") (nets/mj_uwyhNets_ba.py:1189).
    MAX:     keras Maximum = left fold of tf.maximum (ties -> gradient to the earlier input).
    AVG:     keras Average (gated zeros included in the mean).
    SIGNMAX: mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:169-178 -- argmax(|x|) over the
             modality axis (ties -> lowest modality index), gather the signed value.
    winner: forced winner index [B,d] (another implementation's decision); record: receives ``winner`` and ``wgap``
    (best minus second best score) of this run."""
    st = torch.stack(gated, 0)
    if merge == MERGE_AVG:
        return st.mean(0)
    score = st if merge == MERGE_MAX else st.abs()
    # explicit first-winner index so autograd routes ties to the earlier input
    idx = _first_argmax(score.detach()) if winner is None else winner.long()
    if record is not None:
        record["winner"] = idx
        if st.shape[0] > 1:
            t2 = score.detach().topk(2, dim=0).values
            record["wgap"] = t2[0] - t2[1]
    return torch.gather(st, 0, idx.unsqueeze(0)).squeeze(0)


def _first_argmax(st):
    """argmax over dim 0 with ties -> lowest index (tf.math.argmax semantics)."""
    best = st[0]
    idx = torch.zeros(st.shape[1:], dtype=torch.long)
    for m in range(1, st.shape[0]):
        take = st[m] > best
        idx = torch.where(take, torch.full_like(idx, m), idx)
        best = torch.where(take, st[m], best)
    return idx


def batch_dist(x):
    """nets/triplet_loss_all.py:70-77, x: [n,m,d]."""
    x2 = (x * x).sum(2)
    d = x2.unsqueeze(2) + x2.unsqueeze(1) - 2.0 * torch.matmul(x, x.transpose(1, 2))
    d = torch.clamp(d, min=0.0)
    err = d <= 0.0
    d = torch.sqrt(d + err.to(d.dtype) * 1e-16)
    return d * (~err).to(d.dtype)


def triplet_loss_all(labels, emb, margin):
    """General-mask form of nets/triplet_loss_all.py:8-67.  labels [m] (any dtype), emb
    [n,m,d] (or [m,d] -> n=1).  For each part: sum over (a,p,neg) with lab_p==lab_a
    (p==a INCLUDED, :40) and lab_neg!=lab_a of max(margin + d_ap - d_an, 0), divided by the
    number of strictly positive terms (0 if none, :55-59); mean over parts (:61).
    Returns (loss, count_per_part)."""
    if emb.dim() == 2:
        emb = emb.unsqueeze(0)
    lab = labels.reshape(-1)
    same = lab.unsqueeze(0) == lab.unsqueeze(1)                 # [a,b]
    dist = batch_dist(emb)                                      # [n,a,b]
    t = margin + dist.unsqueeze(3) - dist.unsqueeze(2)          # [n,a,p,neg]
    valid = (same.unsqueeze(2) & (~same).unsqueeze(1)).unsqueeze(0)
    t = torch.where(valid, torch.clamp(t, min=0.0), torch.zeros_like(t))
    ssum = t.flatten(1).sum(1)
    cnt = (t > 0).flatten(1).sum(1).to(emb.dtype)
    mean = torch.where(cnt != 0, ssum / torch.clamp(cnt, min=1.0), torch.zeros_like(ssum))
    return mean.mean(0), cnt


def pair_verif_loss(labels, emb, margin):
    """VerifLossLayer(margin) of the Siamese builder UWYHNet.build (nets/mj_uwyhNets_ba.py:230-231 ->
    nets/mj_loss.py:71-91): emb [2B,d] = the two signatures stacked (first elements, then second elements), labels [B]
    (1 same / 0 different):  0.5 * sum_{pos rows} (a-b)^2 + 0.5 * max(0, m - sqrt(sum_{neg rows} (a-b)^2))^2."""
    B = emb.shape[0] // 2
    res2 = (emb[:B] - emb[B:]) ** 2
    lab = labels.reshape(-1)[:B]
    xpos = 0.5 * res2[lab == 1].sum()
    xneg = 0.5 * torch.clamp(margin - torch.sqrt(res2[lab == 0].sum()), min=0.0) ** 2
    return xpos + xneg


def triplet_hard_loss(labels, emb, margin):
    """`tfa.losses.TripletHardLoss(margin)` as compiled by UWYHSemiNet3Mods.compile_hard
    (nets/mj_uwyhNets_ba.py:1302-1306): soft=False, distance_metric="L2".  tensorflow_addons is an un-vendored,
    unpinned dependency (absent here): this restates its published algorithm (tfa/losses/triplet.py `triplet_hard_loss`,
    tfa/losses/metric_learning.py `pairwise_distance`), op by op:
      pdist   = sqrt(max(|a|^2 + |b|^2 - 2ab, 0)), entries <= 0 forced to 0 with zero gradient, diagonal zeroed
      hard_n  = masked_minimum(pdist, labels differ) = min_j((pdist - rowmax) * mask) + rowmax
      hard_p  = masked_maximum(pdist, labels equal minus the diagonal) = max_j((pdist - rowmin) * mask) + rowmin
      loss    = mean_a max(hard_p - hard_n + margin, 0)
    reduce_max / reduce_min split the gradient evenly among tied entries (torch amax / amin do the same).
    emb [m,d] (tfa accepts rank-2 embeddings only).  Returns (loss, number of anchors with a positive term)."""
    lab = labels.reshape(-1, 1)
    m = emb.shape[0]
    pd = batch_dist(emb.unsqueeze(0))[0]
    eye = torch.eye(m, dtype=emb.dtype)
    pd = pd * (1.0 - eye)
    adj = lab == lab.t()
    adj_not = (~adj).to(emb.dtype)
    rowmax = pd.amax(dim=1, keepdim=True)
    hard_n = ((pd - rowmax) * adj_not).amin(dim=1, keepdim=True) + rowmax
    mask_p = adj.to(emb.dtype) - eye
    rowmin = pd.amin(dim=1, keepdim=True)
    hard_p = ((pd - rowmin) * mask_p).amax(dim=1, keepdim=True) + rowmin
    t = torch.clamp(hard_p - hard_n + margin, min=0.0)
    return t.mean(), (t > 0).sum().to(emb.dtype)


def triplet_loss_all_literal_np(labels, emb, margin):
    """Op-by-op numpy restatement of nets/triplet_loss_all.py:33-61 (boolean_mask + reshape,
    which REQUIRES equal #positives / #negatives per anchor).  fp64.  Used to pin the
    general-mask form on balanced batches."""
    emb = np.asarray(emb, dtype=np.float64)
    if emb.ndim == 2:
        emb = emb[None]
    n, m, _ = emb.shape
    lab = np.asarray(labels).reshape(m, 1).T                     # tf.transpose(labels,[1,0])
    lab = np.repeat(lab, n, axis=0)                              # [n,m]
    hp = (lab[:, None, :] == lab[:, :, None]).reshape(-1)
    hn = (lab[:, None, :] != lab[:, :, None]).reshape(-1)
    x2 = (emb ** 2).sum(2)
    d = x2[:, :, None] + x2[:, None, :] - 2.0 * emb @ emb.transpose(0, 2, 1)
    d = np.maximum(d, 0.0)
    err = d <= 0.0
    d = np.sqrt(d + err * 1e-16) * (~err)
    d = d.reshape(-1)
    full_hp = d[hp].reshape(n, m, -1, 1)
    full_hn = d[hn].reshape(n, m, 1, -1)
    metric = np.maximum(margin + (full_hp - full_hn), 0.0).reshape(n, -1)
    s = metric.sum(1)
    c = (metric > 0).sum(1).astype(np.float64)
    mean = np.where(c != 0, s / np.where(c != 0, c, 1.0), 0.0)
    return float(mean.mean()), c


def softmax_ce(logits, onehot):
    """Dense(softmax) + 'categorical_crossentropy' (nets/mj_uwyhNets_ba.py:1214,1243):
    mean over the batch of -sum(onehot * log_softmax(logits)).  Returns (loss, acc)."""
    lsm = F.log_softmax(logits, dim=1)
    loss = -(onehot * lsm).sum(1).mean()
    acc = (logits.argmax(1) == onehot.argmax(1)).to(logits.dtype).mean()
    return loss, acc


def model_forward(inputs, flags, P, cfg: NetConfig, drop_masks=None, code_drop_mask=None,
                  return_all=False, decisions=None, record=None):
    """UWYHSemiNet3Mods.build graph (nets/mj_uwyhNets_ba.py:1163-1214) for 2-D CNN branches;
    with cfg.single the 1-modality graph of UWYHSemiNet.build (:900-915)."""
    outs = {}
    gated = []
    for m in range(cfg.nmods):
        dm = None if drop_masks is None else drop_masks[m]
        rec_m = None
        if record is not None:
            rec_m = record.setdefault(m, {})
        if is3d(cfg, m):
            b = branch3d_forward(inputs[m], P, BRANCH_NAMES[m], cfg)
        else:
            b = branch_forward(inputs[m], P, BRANCH_NAMES[m], cfg, dm, decisions=None if decisions is None else decisions[m],
                               record=rec_m)
        outs[f"branch{m}"] = b
        if cfg.single:
            gated.append(b)
        else:
            if cfg.normbfmerge:
                b = l2_normalize(b, 1)                                  # "nrmbfl2*" Lambda (:1167-1168)
            gated.append(b * flags[m])                                  # :51-54
            if cfg.aux_losses and cfg.nclasses > 0:                     # classprob_{of,gray,depth} (:1222-1225)
                outs[f"aux_logits{m}"] = F.linear(gated[-1], P[f"{AUX_NAMES[m]}/w"], P[f"{AUX_NAMES[m]}/b"])
    if cfg.single:
        sig = gated[0]                                                  # :904 (no normalise)
    else:
        fused = merge_modalities(gated, cfg.merge, None if decisions is None else decisions.get("winner"), record)
        outs["fusion"] = fused
        sig = l2_normalize(fused, 1)
    if getattr(cfg, "postriplet", 1) == 2 and cfg.nc > 0 and not cfg.single:
        # :819-832 -- no normalisation of the fusion; Dense(nc) named "signature" (+ LeakyReLU), its l2_normalize named
        # "code" is `outsignature`; Dropout("dropcode") feeds the classifier
        z = F.linear(fused, P["code/w"], P["code/b"])
        outs["code_reg"] = F.relu(z) if cfg.act == ACT_RELU else z
        x = l2_normalize(_act(z, cfg.act, cfg.alpha), 1)
        outs["signature"] = outs["code"] = x
        feat = x if code_drop_mask is None else x * code_drop_mask
        if cfg.nclasses > 0:
            outs["logits"] = F.linear(feat, P["classprob/w"], P["classprob/b"])
        return outs if return_all else (outs["signature"], outs.get("logits"))
    outs["signature"] = sig
    feat = sig
    if cfg.nc > 0:
        code = F.linear(sig, P["code/w"], P["code/b"])
        # relu: Dense(activation='relu', activity_regularizer) regularises the activated output (:1195-1196);
        # otherwise Dense(activation=None, activity_regularizer) + a separate LeakyReLU: the LINEAR output (:1198-1201)
        outs["code_reg"] = F.relu(code) if cfg.act == ACT_RELU else code
        code = _act(code, cfg.act, cfg.alpha)
        outs["code"] = code
        feat = code if code_drop_mask is None else code * code_drop_mask
    if cfg.nclasses > 0:
        outs["logits"] = F.linear(feat, P["classprob/w"], P["classprob/b"])
    return outs if return_all else (outs["signature"], outs.get("logits"))


def total_loss(inputs, flags, labels, P, cfg: NetConfig, drop_masks=None, code_drop_mask=None, decisions=None,
               record=None):
    """model.compile(loss=[triplet, CE], loss_weights=[wver, wid]) + Keras regularisers
    (nets/mj_uwyhNets_ba.py:1297; :79,:104 kernel L2 = wd*sum(w^2) without 1/2; activity
    L2 on "code" = 1e-3*sum(code^2)/batch).  Returns dict of scalars."""
    outs = model_forward(inputs, flags, P, cfg, drop_masks, code_drop_mask, return_all=True, decisions=decisions,
                         record=record)
    res = {}
    if getattr(cfg, "pair_loss", False):         # UWYHNet.build (:154-245): the loss layer IS the model output
        trip, cnt = pair_verif_loss(labels, outs["signature"], cfg.margin), torch.zeros(())
    elif getattr(cfg, "triplet_hard", False):      # compile_hard (:1302-1306)
        trip, cnt = triplet_hard_loss(labels, outs["signature"], cfg.margin)
    else:
        trip, cnt = triplet_loss_all(labels, outs["signature"], cfg.margin)
    res["triplet"], res["count"] = trip, cnt
    loss = cfg.wver * trip
    if cfg.nclasses > 0:
        onehot = F.one_hot(labels.reshape(-1).long(), cfg.nclasses).to(trip.dtype)
        ce, acc = softmax_ce(outs["logits"], onehot)
        if cfg.label_smoothing > 0:
            # Keras: y_true * (1 - e) + e / num_classes, then the same categorical cross-entropy
            ce, _ = softmax_ce(outs["logits"], onehot * (1.0 - cfg.label_smoothing) + cfg.label_smoothing / cfg.nclasses)
        res["ce"], res["acc"] = ce, acc
        loss = loss + cfg.wid * ce
        if cfg.aux_losses:
            # one more CE per head, all weighted by loss_weights[-1] (:1245-1251, :1264-1268)
            res["aux_ce"] = []
            for m in range(cfg.nmods):
                tgt = onehot * (1.0 - cfg.label_smoothing) + cfg.label_smoothing / cfg.nclasses
                ce_m, _ = softmax_ce(outs[f"aux_logits{m}"], tgt)
                res["aux_ce"].append(ce_m)
                loss = loss + cfg.waux * ce_m
    reg = torch.zeros((), dtype=trip.dtype)
    for m in range(cfg.nmods):
        bn = BRANCH_NAMES[m]
        if not is3d(cfg, m):        # the Conv3D layers carry no kernel regulariser (:346-363), only "grayCode" does
            for li in range(len(cfg.filters_numbers)):
                reg = reg + cfg.weight_decay * (P[f"{bn}/conv{li}/w"] ** 2).sum()
        reg = reg + 1e-3 * (P[f"{bn}/ofCode/w"] ** 2).sum()
    if cfg.nc > 0:
        reg = reg + 1e-3 * (outs["code_reg"] ** 2).sum() / outs["code_reg"].shape[0]
    res["reg"] = reg
    res["loss"] = loss + reg
    res["signature"] = outs["signature"]
    res["logits"] = outs.get("logits")
    return res


def loss_and_grads(inputs, flags, labels, P, cfg, drop_masks=None, code_drop_mask=None, decisions=None, record=None):
    """Autograd gradients of the TOTAL loss (regularisers included) w.r.t. every parameter.  decisions / record: see
    branch_forward -- run with another implementation's discrete decisions injected / export this run's."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    res = total_loss(inputs, flags, labels, Pg, cfg, drop_masks, code_drop_mask, decisions, record)
    res["loss"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    det = lambda v: v.detach() if torch.is_tensor(v) else ([x.detach() for x in v] if isinstance(v, list) else v)
    return {k: det(v) for k, v in res.items()}, grads


def adam_step(P, G, M, V, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-7, Vhat=None, weight_decay=0.0):
    """Keras Adam (mains/mj_trainUWYHGaitNet_DataGen_3mods.py:242): t is the 1-based step.
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t*m/(sqrt(v)+eps).
    Vhat (dict) -> optimizers.Adam(amsgrad=True) (mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:234): the denominator
    uses vhat = max(vhat, v).  weight_decay -> tfa.optimizers.AdamW (:236): decoupled var -= weight_decay * var on the
    pre-update weights, not scaled by the learning rate, then the Adam update."""
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for k in P:
        M[k].mul_(b1).add_(G[k], alpha=1 - b1)
        V[k].mul_(b2).addcmul_(G[k], G[k], value=1 - b2)
        den = V[k]
        if Vhat is not None:
            Vhat[k] = torch.maximum(Vhat[k], V[k])
            den = Vhat[k]
        P[k].sub_(weight_decay * P[k] + lr_t * M[k] / (den.sqrt() + eps))


def sgd_step(P, G, V, t, lr=1e-3, momentum=0.9, decay=0.0):
    """Keras SGD(lr, momentum, decay) (mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:232): t is the 1-based step;
    lr_t = lr / (1 + decay * (t - 1)); v = momentum*v - lr_t*g; w += v."""
    lr_t = lr / (1.0 + decay * (t - 1))
    for k in P:
        V[k].mul_(momentum).sub_(G[k], alpha=lr_t)
        P[k].add_(V[k])


# --------------------------------------------------------------------------------------
# k-NN (sklearn KNeighborsClassifier semantics) and EER metric
# --------------------------------------------------------------------------------------
def knn_search(gallery, queries, k):
    """Exact brute-force k nearest neighbours, Euclidean, fp64 direct sum((q-g)^2),
    ordering (distance, then lower gallery index).  Returns (dist2 [Q,k] f64, idx [Q,k])."""
    G = np.asarray(gallery, dtype=np.float64)
    Q = np.asarray(queries, dtype=np.float64)
    out_d = np.empty((Q.shape[0], k))
    out_i = np.empty((Q.shape[0], k), dtype=np.int64)
    for s in range(0, Q.shape[0], 64):
        q = Q[s:s + 64]
        d2 = np.empty((q.shape[0], G.shape[0]))
        for j in range(q.shape[0]):
            diff = G - q[j]
            d2[j] = np.einsum("ij,ij->i", diff, diff)
        idx = np.lexsort((np.broadcast_to(np.arange(G.shape[0]), d2.shape), d2), axis=1)[:, :k]
        out_i[s:s + 64] = idx
        out_d[s:s + 64] = np.take_along_axis(d2, idx, 1)
    return out_d, out_i


def knn_vote(neigh_labels):
    """weights='uniform' vote of KNeighborsClassifier.predict: most frequent label, ties ->
    smallest label (argmax over the sorted classes_)."""
    out = np.empty(neigh_labels.shape[0], dtype=neigh_labels.dtype)
    for i, row in enumerate(neigh_labels):
        vals, cnts = np.unique(row, return_counts=True)
        out[i] = vals[np.argmax(cnts)]
    return out


def knn_predict(gallery, gallery_labels, queries, k):
    """mains/mj_testUWYHGaitNet_open_tum.py:331-341."""
    _, idx = knn_search(gallery, queries, k)
    return knn_vote(np.asarray(gallery_labels)[idx]), idx


def eer_verif_dist(gt_labels, distances):
    """nets/mj_metrics.py:10-24 restated without sklearn: ROC over score = -distance,
    EER = fpr at argmin|fnr-fpr|, returns (EER, -threshold).  Follows sklearn.roc_curve
    including its default drop_intermediate=True."""
    y = np.asarray(gt_labels).astype(np.float64)
    s = -np.asarray(distances, dtype=np.float64)
    order = np.argsort(-s, kind="mergesort")
    s, y = s[order], y[order]
    distinct = np.where(np.diff(s))[0]
    thr_idx = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y)[thr_idx]
    fps = 1 + thr_idx - tps
    if tps.size > 2:   # roc_curve(drop_intermediate=True): drop collinear ROC points
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        tps, fps, thr_idx = tps[keep], fps[keep], thr_idx[keep]
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    thr = np.r_[np.inf, s[thr_idx]]
    fpr = fps / fps[-1]
    tpr = tps / tps[-1]
    fnr = 1 - tpr
    i = np.nanargmin(np.abs(fnr - fpr))
    return fpr[i], -thr[i]


# --------------------------------------------------------------------------------------
# synthetic step inputs (data/mj_dataGeneratorMMUWYHsingle.py:664-823)
# --------------------------------------------------------------------------------------
NOISE = 1e-9   # data/mj_dataGeneratorMMUWYHsingle.py:102


def synth_batch(cfg: NetConfig, base_rows: int, expand: int, seed: int = 232323, ids_per: int = 2,
                kinds: Sequence[str] = ("of", "gray", "depth"), dtype=np.float32):
    """Synthetic TUM-GAID-shaped step input with the reference's missing-modality expansion.
    Row i*E = all modalities; rows i*E+1.. = copies with modalities disabled:
    even i: min(ex+1, M-1) draws (with replacement) of a modality to disable;
    odd i: only modality (i+ex)%3 enabled (:791-803).  Disabled -> volume = 1e-9, flag 0."""
    rng = np.random.default_rng(seed)
    import random as _r
    pr = _r.Random(seed)
    E = max(expand, 1)
    B = base_rows * E
    M = cfg.nmods
    xs = [np.empty((B, c, cfg.hw, cfg.hw), dtype=dtype) for c in cfg.in_channels]
    fl = [np.ones((B, 1), dtype=dtype) for _ in range(M)]
    labels = np.empty((B, 1), dtype=dtype)
    nids = max(base_rows // ids_per, 1)
    ids = rng.choice(max(cfg.nclasses, nids), size=nids, replace=False)
    for i in range(base_rows):
        lb = ids[(i // ids_per) % nids]
        for m in range(M):
            shp = xs[m].shape[1:]
            kind = kinds[m] if m < len(kinds) else "gray"
            if kind == "of":
                v = np.clip(rng.normal(0, 0.3, shp), -3.3, 3.3)
            elif kind == "sil":
                v = (rng.random(shp) < 0.3).astype(np.float64)
            else:
                v = rng.uniform(-0.5, 0.5, shp)
            xs[m][i * E] = v
        labels[i * E:(i + 1) * E] = lb
        for ex in range(E - 1):
            if i % 2 == 0:
                nd_ = min(ex + 1, M - 1) if E > 2 else pr.randrange(1, M, 1)
                l_dis = [1] * M
                for _ in range(nd_):
                    l_dis[pr.randrange(0, M, 1)] = 0
            else:
                l_dis = [0] * M
                l_dis[(i + ex) % 3 % M] = 1
            r = i * E + ex + 1
            for m in range(M):
                if l_dis[m] == 0:
                    xs[m][r] = NOISE
                    fl[m][r] = 0.0
                else:
                    xs[m][r] = xs[m][i * E]
    return xs, fl, labels


def video_level(codes, labels, vids, preds=None, use_avg=True):
    """mains/mj_testUWYHGaitNet_open_tum.py:355-420 restated loop for loop: per unique video id the
    descriptors are averaged (or max-pooled), labels / predictions voted with statistics.mode (first
    encountered mode on ties, Python >= 3.8; first element if mode() raises)."""
    import statistics
    codes, labels, vids = np.asarray(codes), np.asarray(labels).reshape(-1), np.asarray(vids).reshape(-1)
    uvids = np.unique(vids)
    out_codes, out_labs, out_preds = [], [], []
    for vix in uvids:
        idx = np.where(vids == vix)[0]
        vid_logits = codes[idx, ]
        out_codes.append(vid_logits.mean(axis=0) if use_avg else vid_logits.max(axis=0))
        try:
            out_labs.append(statistics.mode(labels[idx]))
        except Exception:
            out_labs.append(labels[idx][0])
        if preds is not None:
            try:
                out_preds.append(statistics.mode(np.asarray(preds)[idx]))
            except Exception:
                out_preds.append(np.asarray(preds)[idx][0])
    return uvids, np.vstack(out_codes), np.asarray(out_labs), (np.asarray(out_preds) if preds is not None else None)
