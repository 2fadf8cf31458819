"""Python-level operators: one function per C-ABI entry point (include/ugaitnet_b200.h).

Every function takes torch CUDA tensors (or pre-exported ``TRef``s), launches on torch's
current stream and returns nothing / the output tensors it was given.  No function here
computes anything in PyTorch: torch only owns the memory and the stream.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _ffi
from ._ffi import TRef, check, lib, stream_ptr

ACT_LINEAR, ACT_RELU, ACT_LEAKY = 0, 1, 2
MERGE_MAX, MERGE_AVG, MERGE_SIGNMAX = 0, 1, 2

_ctx_cache = {}


def get_ctx(device: Optional[int] = None) -> _ffi.Ctx:
    if device is None:
        device = torch.cuda.current_device()
    c = _ctx_cache.get(device)
    if c is None:
        c = _ctx_cache[device] = _ffi.Ctx(device)
    return c


def _r(x):
    if x is None or isinstance(x, TRef):
        return x
    return TRef(x)


def _p(x):
    return None if x is None else x.ptr


def pack_input(ctx, x_nchw, x_nhwc):
    a, b = _r(x_nchw), _r(x_nhwc)
    check(lib.ugn_pack_input(ctx.h, a.ptr, b.ptr, stream_ptr()))


def pack_input_expand(ctx, x_base, src_row, enable, mirror, x_nhwc, noise=1e-9):
    rs = [_r(t) for t in (x_base, src_row, enable, mirror)]
    o = _r(x_nhwc)
    check(lib.ugn_pack_input_expand(ctx.h, *[_p(r) for r in rs], float(noise), o.ptr, stream_ptr()))


def pack_weight(ctx, w_master, w_packed):
    a, b = _r(w_master), _r(w_packed)
    check(lib.ugn_pack_weight(ctx.h, a.ptr, b.ptr, stream_ptr()))


def split_bf16(ctx, src, dst):
    a, b = _r(src), _r(dst)
    check(lib.ugn_split_bf16(ctx.h, a.ptr, b.ptr, stream_ptr()))


def conv2d_fwd(ctx, x, w, bias, y, pool_idx, act=ACT_RELU, alpha=0.3, pool=True):
    rs = [_r(t) for t in (x, w, bias, y, pool_idx)]
    check(lib.ugn_conv2d_fwd(ctx.h, *[_p(r) for r in rs], int(act), float(alpha), int(bool(pool)), stream_ptr()))


def conv2d_bwd_act(ctx, dy, y, pool_idx, dz, act=ACT_RELU, alpha=0.3, pool=True, db=None):
    rs = [_r(t) for t in (dy, y, pool_idx, dz, db)]
    check(lib.ugn_conv2d_bwd_act(ctx.h, *[_p(r) for r in rs], int(act), float(alpha), int(bool(pool)), stream_ptr()))


def conv2d_dgrad(ctx, dz, w, dx):
    rs = [_r(t) for t in (dz, w, dx)]
    check(lib.ugn_conv2d_dgrad(ctx.h, *[_p(r) for r in rs], stream_ptr()))


def conv2d_wgrad(ctx, x, dz, dw, db):
    rs = [_r(t) for t in (x, dz, dw, db)]
    check(lib.ugn_conv2d_wgrad(ctx.h, *[_p(r) for r in rs], stream_ptr()))


def flatten_chw(ctx, y, flat):
    a, b = _r(y), _r(flat)
    check(lib.ugn_flatten_chw(ctx.h, a.ptr, b.ptr, stream_ptr()))


def unflatten_chw(ctx, dflat, dy):
    a, b = _r(dflat), _r(dy)
    check(lib.ugn_unflatten_chw(ctx.h, a.ptr, b.ptr, stream_ptr()))


def linear_fwd(ctx, x, w, bias, drop_mask, y, y16=None, act=ACT_LINEAR, alpha=0.3):
    rs = [_r(t) for t in (x, w, bias, drop_mask, y, y16)]
    check(lib.ugn_linear_fwd(ctx.h, *[_p(r) for r in rs], int(act), float(alpha), stream_ptr()))


def act_mask_bwd(ctx, dy, y, drop_mask, dz, dz16=None, act=ACT_LINEAR, alpha=0.3):
    rs = [_r(t) for t in (dy, y, drop_mask, dz, dz16)]
    check(lib.ugn_act_mask_bwd(ctx.h, *[_p(r) for r in rs], int(act), float(alpha), stream_ptr()))


def linear_bwd(ctx, x, w, dz, dx, dw, db):
    rs = [_r(t) for t in (x, w, dz, dx, dw, db)]
    check(lib.ugn_linear_bwd(ctx.h, *[_p(r) for r in rs], stream_ptr()))


def fuse_fwd(ctx, br: Sequence, flags: Sequence, sig, sig16, winner, inv_norm, merge, normalize=True):
    rb = [_r(t) for t in br]
    rf = [_r(t) for t in flags]
    rs = [_r(t) for t in (sig, sig16, winner, inv_norm)]
    check(lib.ugn_fuse_fwd(ctx.h, len(rb), _ffi.ptr_array(rb), _ffi.ptr_array(rf), *[_p(r) for r in rs],
                           int(merge), int(bool(normalize)), stream_ptr()))


def fuse_bwd(ctx, dsig, sig, winner, inv_norm, flags: Sequence, dbr: Sequence, merge, normalize=True):
    rf = [_r(t) for t in flags]
    rd = [_r(t) for t in dbr]
    rs = [_r(t) for t in (dsig, sig, winner, inv_norm)]
    check(lib.ugn_fuse_bwd(ctx.h, len(rd), *[_p(r) for r in rs], _ffi.ptr_array(rf), _ffi.ptr_array(rd),
                           int(merge), int(bool(normalize)), stream_ptr()))


def softmax_ce(ctx, logits, labels, loss_acc, dlogits, scale=1.0):
    rs = [_r(t) for t in (logits, labels, loss_acc, dlogits)]
    check(lib.ugn_softmax_ce(ctx.h, *[_p(r) for r in rs], float(scale), stream_ptr()))


def triplet_workspace_bytes(n: int, B: int) -> int:
    return int(lib.ugn_triplet_workspace_bytes(int(n), int(B)))


def triplet_all(ctx, emb, labels, margin, scale, out, demb, workspace):
    rs = [_r(t) for t in (emb, labels)]
    ro = [_r(t) for t in (out, demb, workspace)]
    check(lib.ugn_triplet_all(ctx.h, rs[0].ptr, rs[1].ptr, float(margin), float(scale), *[_p(r) for r in ro],
                              stream_ptr()))


def triplet_hard(ctx, emb, labels, margin, scale, out, demb, workspace, emb16=None):
    """tfa.losses.TripletHardLoss (compile_hard, nets/mj_uwyhNets_ba.py:1302-1306); emb [B,d]."""
    rs = [_r(t) for t in (emb, emb16, labels)]
    ro = [_r(t) for t in (out, demb, workspace)]
    check(lib.ugn_triplet_hard(ctx.h, rs[0].ptr, _p(rs[1]), rs[2].ptr, float(margin), float(scale),
                               *[_p(r) for r in ro], stream_ptr()))


def adam_step(ctx, w, g, m, v, seg_off, seg_l2, lr_t, beta1=0.9, beta2=0.999, eps=1e-7, gscale=1.0,
              reg_out=None, lr_dev=None, pack_table=None, pack_planes=0, pack_f16=0):
    rs = [_r(t) for t in (w, g, m, v, seg_off, seg_l2)]
    ro = [_r(reg_out), _r(lr_dev), _r(pack_table)]
    check(lib.ugn_adam_step(ctx.h, *[_p(r) for r in rs], float(lr_t), float(beta1), float(beta2), float(eps),
                            float(gscale), _p(ro[0]), _p(ro[1]), _p(ro[2]), int(pack_planes), int(pack_f16),
                            stream_ptr()))


def sgd_step(ctx, w, g, v, seg_off, seg_l2, lr, momentum=0.9, gscale=1.0, reg_out=None, lr_dev=None,
             pack_table=None, pack_planes=0, pack_f16=0):
    rs = [_r(t) for t in (w, g, v, seg_off, seg_l2)]
    ro = [_r(reg_out), _r(lr_dev), _r(pack_table)]
    check(lib.ugn_sgd_step(ctx.h, *[_p(r) for r in rs], float(lr), float(momentum), float(gscale), _p(ro[0]),
                           _p(ro[1]), _p(ro[2]), int(pack_planes), int(pack_f16), stream_ptr()))


def knn_workspace_bytes(Q, N, D, k) -> int:
    return int(lib.ugn_knn_workspace_bytes(int(Q), int(N), int(D), int(k)))


def knn_gallery_norms(ctx, gallery, g2, gmax2=None):
    a, b, c = _r(gallery), _r(g2), _r(gmax2)
    check(lib.ugn_knn_gallery_norms(ctx.h, a.ptr, b.ptr, _p(c), stream_ptr()))


def knn_pack(ctx, x, x16):
    a, b = _r(x), _r(x16)
    check(lib.ugn_knn_pack(ctx.h, a.ptr, b.ptr, stream_ptr()))


def knn_topk_tc(ctx, queries, q16, gallery, g16, g2, gmax2, labels, k, idx_base, out_d2, out_idx, out_lab, flags,
                workspace):
    rs = [_r(t) for t in (queries, q16, gallery, g16, g2, gmax2, labels)]
    ro = [_r(t) for t in (out_d2, out_idx, out_lab, flags, workspace)]
    check(lib.ugn_knn_topk_tc(ctx.h, *[_p(r) for r in rs], int(k), int(idx_base), *[_p(r) for r in ro],
                              stream_ptr()))


def knn_topk(ctx, queries, gallery, g2, labels, k, idx_base, out_d2, out_idx, out_lab, workspace):
    rs = [_r(t) for t in (queries, gallery, g2, labels)]
    ro = [_r(t) for t in (out_d2, out_idx, out_lab, workspace)]
    check(lib.ugn_knn_topk(ctx.h, *[_p(r) for r in rs], int(k), int(idx_base), *[_p(r) for r in ro], stream_ptr()))


def knn_merge_vote(ctx, d2, idx, lab, k, out_d2, out_idx, out_lab, pred):
    rs = [_r(t) for t in (d2, idx, lab)]
    ro = [_r(t) for t in (out_d2, out_idx, out_lab, pred)]
    check(lib.ugn_knn_merge_vote(ctx.h, *[_p(r) for r in rs], int(k), *[_p(r) for r in ro], stream_ptr()))


def gemm_bf16(ctx, A, a_mn, B, b_mn, C, accumulate=False):
    rs = [_r(t) for t in (A, B, C)]
    check(lib.ugn_gemm_bf16(ctx.h, rs[0].ptr, int(a_mn), rs[1].ptr, int(b_mn), rs[2].ptr, int(bool(accumulate)),
                            stream_ptr()))


def grad_scale_update(ctx, ref, target=1024.0):
    r = _r(ref)
    check(lib.ugn_grad_scale_update(ctx.h, r.ptr, float(target), stream_ptr()))


def grad_scale_set(ctx, scale: float):
    check(lib.ugn_grad_scale_set(ctx.h, float(scale), stream_ptr()))
