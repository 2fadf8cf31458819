"""Second, INDEPENDENT restatement of the TensorFlow-executed part of the step.  TEST INFRASTRUCTURE ONLY.

``oracle/ugait_oracle.py`` restates the step with PyTorch operators and lets autograd derive the backward pass.
Nothing of the real reference can execute here or on the GPU box (TensorFlow / Keras / h5py are absent from the image
and from /opt/wheelhouse, probed in round 2: gpurun_out/r02a/probe.log), so the torch restatement is cross-checked by a
second one that shares NO code and NO library operator with it: plain numpy fp64, every layer written out from its
definition (sliding windows + einsum), and every backward formula derived by hand.  tests/test_oracle.py asserts

    torch oracle (autograd)  ==  this file (hand-derived)  ==  central finite differences of this file's forward

on losses, descriptors, every gradient tensor and an Adam / SGD update.  Two independent derivations agreeing to
1e-9 pin the ARITHMETIC of the restated graph; what stays unpinned is only whether the graph itself is the one
TensorFlow 2.3 would build from the call sites cited below.

Citations (relative to /root/reference/):
  branch      nets/mj_uwyhNets_ba.py:67-107 (ReLU), :110-152 (LeakyReLU)
  gate        nets/mj_uwyhNets_ba.py:51-54, :1164-1185
  fusion      nets/mj_uwyhNets_ba.py:1189; sign_max mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:169-178
  signature   nets/mj_uwyhNets_ba.py:1191 (tf.math.l2_normalize, eps 1e-12)
  FC1 / FC2   nets/mj_uwyhNets_ba.py:1194-1214; compile :1239-1297
  triplet     nets/triplet_loss_all.py:8-77
  optimiser   mains/mj_trainUWYHGaitNet_DataGen_3mods.py:242-251
"""
from __future__ import annotations

import math

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

MERGE_MAX, MERGE_AVG, MERGE_SIGNMAX = 0, 1, 2
ACT_LINEAR, ACT_RELU, ACT_LEAKY = 0, 1, 2
BRANCH_NAMES = ("ofBranch", "grayBranch", "depthBranch")


# ------------------------------------------------------------------------------------------------ layers, forward
def conv_valid(x, w, b):
    """Conv2D(padding='valid', strides 1, channels_first): y[n,o,i,j] = b[o] + sum_{c,u,v} x[n,c,i+u,j+v] w[o,c,u,v]."""
    k = w.shape[2]
    win = sliding_window_view(x, (k, k), axis=(2, 3))            # [n,c,i,j,u,v]
    return np.einsum("ncijuv,ocuv->noij", win, w, optimize=True) + b[None, :, None, None]


def conv_valid_bwd(x, w, dy):
    k = w.shape[2]
    win = sliding_window_view(x, (k, k), axis=(2, 3))
    dw = np.einsum("ncijuv,noij->ocuv", win, dy, optimize=True)
    db = dy.sum(axis=(0, 2, 3))
    # dx[n,c,p,q] = sum_{o,u,v} dy[n,o,p-u,q-v] w[o,c,u,v]: 'full' correlation of dy with the flipped kernel
    dyp = np.pad(dy, ((0, 0), (0, 0), (k - 1, k - 1), (k - 1, k - 1)))
    winy = sliding_window_view(dyp, (k, k), axis=(2, 3))         # [n,o,p,q,u',v'] = dyp[p+u', q+v'], dy index p+u'-(k-1)
    dx = np.einsum("nopquv,ocuv->ncpq", winy, w[:, :, ::-1, ::-1], optimize=True)
    return dx, dw, db


def act_fwd(z, act, alpha):
    if act == ACT_RELU:
        return np.where(z > 0, z, 0.0)
    if act == ACT_LEAKY:
        return np.where(z > 0, z, alpha * z)
    return z


def act_bwd(z, dy, act, alpha):
    if act == ACT_RELU:
        return np.where(z > 0, dy, 0.0)
    if act == ACT_LEAKY:
        return np.where(z > 0, dy, alpha * dy)
    return dy


def pool2_fwd(a):
    """MaxPooling2D(2) ('valid': an odd trailing row / column is dropped).  Gradient goes to the FIRST maximum of
    the window in (dy, dx) order (TF MaxPoolGrad)."""
    n, c, H, W = a.shape
    hp, wp = H // 2, W // 2
    win = a[:, :, :2 * hp, :2 * wp].reshape(n, c, hp, 2, wp, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, hp, wp, 4)
    idx = np.argmax(win, axis=4)                                 # numpy argmax: first occurrence
    return np.take_along_axis(win, idx[..., None], 4)[..., 0], idx


def pool2_bwd(shape, idx, dy):
    n, c, H, W = shape
    hp, wp = H // 2, W // 2
    dwin = np.zeros((n, c, hp, wp, 4))
    np.put_along_axis(dwin, idx[..., None], dy[..., None], 4)
    da = np.zeros(shape)
    da[:, :, :2 * hp, :2 * wp] = dwin.reshape(n, c, hp, wp, 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, 2 * hp, 2 * wp)
    return da


def branch(x, P, bn, cfg, drop_mask):
    """Forward of one modality branch; returns the output and the tape the hand-written backward needs."""
    tape = []
    h = x
    nl = len(cfg.filters_numbers)
    for li in range(nl):
        w, b = P[f"{bn}/conv{li}/w"], P[f"{bn}/conv{li}/b"]
        z = conv_valid(h, w, b)
        a = act_fwd(z, cfg.act, cfg.alpha)
        if li != nl - 1:
            y, idx = pool2_fwd(a)
        else:
            y, idx = a, None
        tape.append((h, z, idx))
        h = y
    shp = h.shape
    flat = h.reshape(h.shape[0], -1)                             # Flatten of a channels_first tensor: (C,H,W) order
    h1 = flat @ P[f"{bn}/dense/w"].T + P[f"{bn}/dense/b"]
    h1d = h1 * drop_mask if drop_mask is not None else h1        # Dropout: inverted mask (already scaled)
    out = h1d @ P[f"{bn}/ofCode/w"].T + P[f"{bn}/ofCode/b"]
    return out, (tape, shp, flat, h1d)


def branch_bwd(dout, P, bn, cfg, drop_mask, saved, G):
    tape, shp, flat, h1d = saved
    G[f"{bn}/ofCode/w"] = dout.T @ h1d
    G[f"{bn}/ofCode/b"] = dout.sum(0)
    dh1 = dout @ P[f"{bn}/ofCode/w"]
    if drop_mask is not None:
        dh1 = dh1 * drop_mask
    G[f"{bn}/dense/w"] = dh1.T @ flat
    G[f"{bn}/dense/b"] = dh1.sum(0)
    dh = (dh1 @ P[f"{bn}/dense/w"]).reshape(shp)
    for li in range(len(tape) - 1, -1, -1):
        x, z, idx = tape[li]
        da = pool2_bwd(z.shape, idx, dh) if idx is not None else dh
        dz = act_bwd(z, da, cfg.act, cfg.alpha)
        dh, G[f"{bn}/conv{li}/w"], G[f"{bn}/conv{li}/b"] = conv_valid_bwd(x, P[f"{bn}/conv{li}/w"], dz)


# ------------------------------------------------------------------------------------------------ losses
def batch_dist(x):
    """nets/triplet_loss_all.py:70-77 on one part: x [m,d] -> (dist [m,m], error_mask)."""
    x2 = (x * x).sum(1)
    d2 = np.maximum(x2[:, None] + x2[None, :] - 2.0 * x @ x.T, 0.0)
    err = d2 <= 0.0
    d = np.sqrt(d2 + err * 1e-16) * (~err)
    return d, err


def triplet_fwd_bwd(labels, emb, margin):
    """Batch-all triplet on emb [m,d] (one part).  Loss = sum of positive hinge terms / their count (:55-59).
    Backward by hand: the count is a constant; dL/dd_ap = +#active(neg)/c, dL/dd_an = -#active(p)/c;
    d(d_ab)/dx_a = (x_a - x_b)/d_ab, zero where the error mask is set."""
    lab = np.asarray(labels).reshape(-1)
    m = lab.shape[0]
    same = lab[:, None] == lab[None, :]
    dist, err = batch_dist(emb)
    loss_sum, count = 0.0, 0
    dD = np.zeros((m, m))
    for a in range(m):
        pos = np.where(same[a])[0]                               # includes p == a (:40)
        neg = np.where(~same[a])[0]
        t = margin + dist[a, pos][:, None] - dist[a, neg][None, :]
        act = t > 0
        loss_sum += t[act].sum()
        count += int(act.sum())
        dD[a, pos] += act.sum(1)
        dD[a, neg] -= act.sum(0)
    if count == 0:
        return 0.0, 0, np.zeros_like(emb)
    dD /= count
    with np.errstate(divide="ignore", invalid="ignore"):
        coef = np.where(err, 0.0, dD / np.where(err, 1.0, dist))  # [a,b]: dL/dd_ab / d_ab
    # dL/dx_a = sum_b coef[a,b](x_a-x_b) + sum_b coef[b,a](x_a-x_b)   ((a,b) and (b,a) are distinct terms)
    S = coef + coef.T
    demb = S.sum(1)[:, None] * emb - S @ emb
    return loss_sum / count, count, demb


def softmax_ce_fwd_bwd(logits, labels, nclasses, smoothing=0.0):
    """Dense(softmax) + categorical_crossentropy, mean over the batch; label smoothing y(1-e) + e/C."""
    B = logits.shape[0]
    y = np.zeros((B, nclasses))
    y[np.arange(B), np.asarray(labels).reshape(-1).astype(int)] = 1.0
    yt = y * (1.0 - smoothing) + smoothing / nclasses if smoothing > 0 else y
    zmax = logits.max(1, keepdims=True)
    lse = zmax + np.log(np.exp(logits - zmax).sum(1, keepdims=True))
    p = np.exp(logits - lse)
    loss = float(-(yt * (logits - lse)).sum(1).mean())
    acc = float((logits.argmax(1) == y.argmax(1)).mean())
    return loss, acc, (p * yt.sum(1, keepdims=True) - yt) / B


# ------------------------------------------------------------------------------------------------ whole step
def step(inputs, flags, labels, P, cfg, drop_masks=None, code_drop_mask=None, want_grads=True):
    """Total loss of model.compile(loss=[triplet, CE], loss_weights=[wver, wid]) + Keras regularisers and its
    gradient w.r.t. every parameter, by hand.  Returns (dict of scalars + signature, grads or None)."""
    M = cfg.nmods
    outs, saved = [], []
    for m in range(M):
        o, s = branch(inputs[m], P, BRANCH_NAMES[m], cfg, None if drop_masks is None else drop_masks[m])
        outs.append(o)
        saved.append(s)
    if cfg.single:
        sig = outs[0]
    else:
        gated = np.stack([outs[m] * flags[m] for m in range(M)], 0)              # [M,B,d]
        if cfg.merge == MERGE_AVG:
            fused, win = gated.mean(0), None
        else:
            score = gated if cfg.merge == MERGE_MAX else np.abs(gated)
            win = np.argmax(score, axis=0)                                      # first maximum: lowest modality index
            fused = np.take_along_axis(gated, win[None], 0)[0]
        ss = np.maximum((fused * fused).sum(1, keepdims=True), 1e-12)
        inv = 1.0 / np.sqrt(ss)
        sig = fused * inv
    res = {"signature": sig}
    trip, cnt, dsig = triplet_fwd_bwd(labels, sig, cfg.margin)
    res["triplet"], res["count"] = trip, cnt
    dsig = cfg.wver * dsig
    loss = cfg.wver * trip
    G = {}
    B = sig.shape[0]
    feat = sig
    if cfg.nc > 0:
        zc = sig @ P["code/w"].T + P["code/b"]
        code = act_fwd(zc, cfg.act, cfg.alpha)
        reg_on = code if cfg.act == ACT_RELU else zc                            # :1195-1196 vs :1198-1201
        feat = code * code_drop_mask if code_drop_mask is not None else code
    if cfg.nclasses > 0:
        logits = feat @ P["classprob/w"].T + P["classprob/b"]
        ce, acc, dlog = softmax_ce_fwd_bwd(logits, labels, cfg.nclasses, cfg.label_smoothing)
        res["ce"], res["acc"], res["logits"] = ce, acc, logits
        loss += cfg.wid * ce
        dlog = cfg.wid * dlog
        G["classprob/w"], G["classprob/b"] = dlog.T @ feat, dlog.sum(0)
        dfeat = dlog @ P["classprob/w"]
    else:
        dfeat = np.zeros_like(feat)
    reg = 0.0
    if cfg.nc > 0:
        reg += 1e-3 * (reg_on ** 2).sum() / B                                   # activity_regularizer l2(1e-3) / batch
        dcode = dfeat * code_drop_mask if code_drop_mask is not None else dfeat
        if cfg.act == ACT_RELU:
            dcode = dcode + 2e-3 * code / B
        dzc = act_bwd(zc, dcode, cfg.act, cfg.alpha)
        if cfg.act != ACT_RELU:
            dzc = dzc + 2e-3 * zc / B
        G["code/w"], G["code/b"] = dzc.T @ sig, dzc.sum(0)
        dsig = dsig + dzc @ P["code/w"]
    else:
        dsig = dsig + dfeat
    for m in range(M):
        bn = BRANCH_NAMES[m]
        for li in range(len(cfg.filters_numbers)):
            reg += cfg.weight_decay * (P[f"{bn}/conv{li}/w"] ** 2).sum()        # l2(wd): wd * sum w^2, no 1/2
        reg += 1e-3 * (P[f"{bn}/ofCode/w"] ** 2).sum()
    res["reg"], res["loss"] = reg, loss + reg
    if not want_grads:
        return res, None
    if cfg.single:
        douts = [dsig]
    else:
        # l2_normalize backward: y = x * inv, inv = rsqrt(max(ss, eps)); below eps inv is a constant
        live = ((fused * fused).sum(1, keepdims=True) > 1e-12)
        dfused = inv * dsig - np.where(live, (inv ** 3) * fused * (fused * dsig).sum(1, keepdims=True), 0.0)
        if cfg.merge == MERGE_AVG:
            douts = [dfused / M * flags[m] for m in range(M)]
        else:
            douts = [np.where(win == m, dfused, 0.0) * flags[m] for m in range(M)]
    for m in range(M):
        branch_bwd(douts[m], P, BRANCH_NAMES[m], cfg, None if drop_masks is None else drop_masks[m], saved[m], G)
        bn = BRANCH_NAMES[m]
        for li in range(len(cfg.filters_numbers)):
            G[f"{bn}/conv{li}/w"] = G[f"{bn}/conv{li}/w"] + 2 * cfg.weight_decay * P[f"{bn}/conv{li}/w"]
        G[f"{bn}/ofCode/w"] = G[f"{bn}/ofCode/w"] + 2e-3 * P[f"{bn}/ofCode/w"]
    return res, G


def adam_update(P, G, M, V, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-7):
    """Keras Adam: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps)."""
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for k in P:
        M[k] = b1 * M[k] + (1 - b1) * G[k]
        V[k] = b2 * V[k] + (1 - b2) * G[k] * G[k]
        P[k] = P[k] - lr_t * M[k] / (np.sqrt(V[k]) + eps)
