"""GaitSet branch type of UGaitNet: step engine.

``UWYHSemiNet3Mods.build(..., gaitset=True)`` (/root/reference/nets/mj_uwyhNets_ba.py:1110-1214) around
``UWYHSemiNet.build_gaitset_branch`` (:420-484): per-frame convolutions (TimeDistributed), set pooling over
the T frames, the global branch, horizontal pyramid pooling into 62 parts, the per-part ``MatMul`` (:23-48),
then gate x use-flag, fusion, ``l2_normalize(axis=1)`` on ``[62, B, 256]`` (axis 1 is the batch axis in this
layout -- kept literally), FC1 "code", transpose + Flatten + FC2 "classprob", batch-all triplet over the
62 parts.

The engine reuses the conv / dense / loss / optimiser kernels of the stacked-CNN path through the same C
ABI; ``padding='same'`` is realised with zero-bordered activation buffers (ugn_pad_hw / ugn_crop_hw), the
64-wide layers live in a 2-way column-split layout, and the first 5x5 convolution over 1 | 2 input channels
(K = 25 | 50, far below a tensor-core tile) is a fused fp32 kernel of its own (ugn_gs_conv1_fwd / _wgrad).
PyTorch owns memory, streams and NCCL only.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from ._ffi import TRef, check, lib, ptr_array, stream_ptr
from .config import (ACT_LEAKY, ACT_LINEAR, BRANCH_NAMES, FUSE3_NO_NORM, GS_ALPHA, GS_CONVS, GS_PARTS, MERGE_MAX, GaitSetConfig,
                     round_up)
from .net import GRAD_SCALE_TARGET, UGaitEngine, _Seg

# conv name -> (input buffer, output buffer, pooled)
_GS_WIRING = {"a2": ("a1p", "a2", 1), "b1": ("g0p", "b1", 0), "b2": ("b1p", "b2", 1),
              "a3": ("a2p", "a3", 0), "a4": ("a3p", "a4", 1), "b3": ("s1p", "b3", 0), "b4": ("b3p", "b4", 0),
              "a5": ("a4p", "a5", 0), "a6": ("a5p", "a6", 0)}


class GaitSetEngine(UGaitEngine):
    """One training / descriptor-extraction step of the gaitset=True graph on one GPU."""

    def __init__(self, cfg: GaitSetConfig, force_split: Optional[int] = None, **kw):
        self.force_split = force_split       # tests: exercise the split layout on small frames / in fp32 mode
        if cfg.single and (cfg.nmods != 1 or cfg.nc > 0):
            raise ValueError("GaitSetConfig.single: the 1-modality graph has one branch and no FC1 "
                             "(nets/mj_uwyhNets_ba.py:890-911)")
        # (UGaitEngine sets these in ITS _build_arena, which this class overrides)
        self.aux = False
        self.post2 = getattr(cfg, "postriplet", 1) == 2 and cfg.nc > 0 and not cfg.single
        super().__init__(cfg, **kw)

    # ------------------------------------------------------------------ parameters
    def _conv_shapes(self, m: int):
        """name -> (cout_master, k, cin_master, cin_packed, cout)."""
        c = self.cfg.in_channels[m]
        c2 = 32
        out = {}
        for name, cin, co, k in GS_CONVS:
            if name == "a1":
                out[name] = (co, 1, 25 * c, 25 * c, co)        # fused fp32 kernel: reads the f32 master
            else:
                cm = cin
                cp = c2 if name in ("b1", "a3") else cin
                out[name] = (c2 if name == "a2" else co, k, cm, cp, co)
        return out

    def _build_arena(self):
        cfg = self.cfg
        segs: List[_Seg] = []
        off = 0
        self.true_co: Dict[str, int] = {}

        def add(name, shape, l2=0.0):
            nonlocal off
            n = 1
            for s in shape:
                n *= s
            segs.append(_Seg(name, tuple(shape), off, n, l2))
            off += round_up(n, 64)

        for m in range(cfg.nmods):
            bn = BRANCH_NAMES[m]
            for name, (co, k, cin, cp, true_co) in self._conv_shapes(m).items():
                add(f"{bn}/{name}/w", (co, k, k, cin))
                self.true_co[f"{bn}/{name}/w"] = true_co
            add(f"{bn}/matmul/w", (GS_PARTS, 128, cfg.hidden))
        feat = cfg.hidden
        if cfg.nc > 0:
            add("code/w", (cfg.nc, cfg.hidden))
            add("code/b", (cfg.nc,))
            feat = cfg.nc
        if cfg.nclasses > 0:
            add("classprob/w", (cfg.nclasses, GS_PARTS * feat))
            add("classprob/b", (cfg.nclasses,))
        self.segs = {s.name: s for s in segs}
        self.seg_list = segs
        self.n_arena = off
        self.buckets = {}
        for m in range(cfg.nmods):
            mine = [s for s in segs if s.name.startswith(BRANCH_NAMES[m] + "/")]
            self.buckets[m] = (mine[0].off, round_up(mine[-1].off + mine[-1].n, 64))
        heads = [s for s in segs if s.name.split("/")[0] in ("code", "classprob")]
        self.buckets["heads"] = (heads[0].off, off) if heads else None
        d = self.dev
        self.w = self._new_arena(off, exchanged=True)
        self.g = self._new_arena(off, exchanged=True)
        self.m = torch.zeros(off, device=d)
        self.v = torch.zeros(off, device=d)
        self.seg_off = torch.tensor([s.off for s in segs] + [off], dtype=torch.int64, device=d)
        self.seg_l2 = torch.tensor([s.l2 for s in segs], dtype=torch.float32, device=d)
        self.reg_out = torch.zeros(1, device=d)
        self.philox = False            # (the only dropout of this graph is "dropcode" after FC1: a mask tensor)
        self.gstage = None             # (no early-final gradient ranges worth pushing: 3 M parameters)
        self.cw_arena = None           # (compute copies stay local: re-packed after the f32 all-gather)
        if self._symm:                 # fused data-parallel exchange: summed over the ranks by peer atomics (net.py)
            r = self._new_arena(64, exchanged=True)
            if len(self._symm) > 2:
                self.reg_out = r[:1]
        self.lr_dev = torch.zeros(1, device=d)
        self._lr_host = torch.zeros(1).pin_memory()
        self.R = {k: TRef(t) for k, t in dict(w=self.w, g=self.g, m=self.m, v=self.v, seg_off=self.seg_off,
                                               seg_l2=self.seg_l2, reg_out=self.reg_out, lr_dev=self.lr_dev).items()}
        self.pw, self.pg_, self.Rw, self.Rg = {}, {}, {}, {}
        for s in segs:
            self.pw[s.name] = self.w[s.off:s.off + s.n].view(s.shape)
            self.pg_[s.name] = self.g[s.off:s.off + s.n].view(s.shape)
            self.Rw[s.name] = TRef(self.pw[s.name])
            self.Rg[s.name] = TRef(self.pg_[s.name])
        # compute copies of the conv kernels: [P][Cout][k][k][Cp] 16-bit planes, or the f32 master itself
        self.cw, self.Rcw = {}, {}
        for m in range(cfg.nmods):
            bn = BRANCH_NAMES[m]
            for name, (co, k, cin, cp, _) in self._conv_shapes(m).items():
                key = f"{bn}/{name}/w"
                if name == "a1":
                    self.cw[key] = self.pw[key]
                elif self.P:
                    self.cw[key] = torch.zeros((self.P, co, k, k, cp), dtype=self.dt16, device=d)
                elif cp != cin:
                    self.cw[key] = torch.zeros((co, k, k, cp), device=d)
                else:
                    self.cw[key] = self.pw[key]
        for k_, t in self.cw.items():
            self.Rcw[k_] = TRef(t)
        self.pack_table = None
        self._fused_pack = set()
        if self.P:
            tab = torch.zeros(len(segs), 2, dtype=torch.int64)
            for i, sg in enumerate(segs):
                t = self.cw.get(sg.name)
                if t is not None and t.dtype == self.dt16 and tuple(t.shape[1:]) == sg.shape and sg.n % 4 == 0:
                    tab[i, 0], tab[i, 1] = t.data_ptr(), sg.n
                    self._fused_pack.add(sg.name)
            self.pack_table = tab.to(d)
            self.R["pack_table"] = TRef(self.pack_table)

    def init_weights(self, seed: int):
        """Keras defaults: glorot_uniform conv kernels without bias (:428-466), GlorotUniform on the rank-3
        MatMul kernel (:33-34; receptive field 62), glorot Dense + zero bias."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for s in self.seg_list:
            w = self.pw[s.name]
            w.zero_()
            if s.name.endswith("/b"):
                continue
            if len(s.shape) == 4:
                co = self.true_co[s.name]
                _, kh, kw, cin = s.shape
                if s.name.endswith("/a1/w"):
                    fan_in, fan_out = cin, co * 25
                else:
                    fan_in, fan_out = cin * kh * kw, co * kh * kw
                shape = (co, kh, kw, cin)
            elif len(s.shape) == 3:
                fan_in, fan_out = GS_PARTS * s.shape[1], GS_PARTS * s.shape[2]
                shape = s.shape
            else:
                fan_out, fan_in = s.shape
                shape = s.shape
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            vals = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * limit
            w[:shape[0]].copy_(vals)
        self.repack_weights()

    def load_params(self, params: Dict[str, torch.Tensor]):
        """params in the oracle / PyTorch layout: conv [Cout,Cin,kh,kw] ("a1": [32,c,5,5]), MatMul
        [62,128,hidden], dense [out,in]."""
        for name, val in params.items():
            s = self.segs[name]
            v = val.detach().to(torch.float32)
            if len(s.shape) == 4:
                v = v.permute(0, 2, 3, 1).contiguous()                 # [Cout,kh,kw,Cin]
                if name.endswith("/a1/w"):
                    v = v.reshape(v.shape[0], 1, 1, -1)               # tap-major (ky,kx,ci), the order ugn_gs_conv1_* read
                self.pw[name].zero_()
                self.pw[name][:v.shape[0]].copy_(v.to(self.dev))
            else:
                self.pw[name].copy_(v.contiguous().to(self.dev))
        self.repack_weights()

    def _export(self, views) -> Dict[str, torch.Tensor]:
        out = {}
        for s in self.seg_list:
            v = views[s.name].detach().clone()
            if len(s.shape) == 4:
                v = v[:self.true_co[s.name]]
                if s.name.endswith("/a1/w"):
                    c = v.shape[-1] // 25
                    v = v.reshape(v.shape[0], 5, 5, c)
                v = v.permute(0, 3, 1, 2).contiguous()
            out[s.name] = v
        return out

    # ------------------------------------------------------------------ plans
    def split_for(self, Hs: int) -> int:
        """Column halves of the 64-wide layers: the tensor-core conv kernels keep an input row in at most 64
        pixel slots, a zero-bordered 66-wide row does not fit, two overlapping 34-wide halves do."""
        if self.force_split is not None:
            return self.force_split
        return 2 if (self.P and Hs + 2 > 64 and Hs % 4 == 0) else 1

    def plan(self, B: int, train: bool) -> "_GsPlan":
        self._check_batch(B)
        key = (B, train)
        p = self._plans.get(key)
        if p is None:
            p = self._plans[key] = _GsPlan(self, B, train)
        return p

    def _set_inputs(self, p, inputs, flags, labels=None, drop_masks=None, code_drop_mask=None):
        self._check_shapes(p, inputs, flags)
        for m in range(self.cfg.nmods):
            p.br[m].x_in.copy_(inputs[m], non_blocking=True)
            if flags is not None:
                p.flags[m].copy_(flags[m].reshape(-1, 1), non_blocking=True)
        if p.train and self.cfg.dropout > 0.001 and self.cfg.nc > 0:
            if code_drop_mask is not None:
                p.cmask.copy_(code_drop_mask.reshape(p.cmask.shape))
            else:
                keep = 1.0 - self.cfg.dropout
                p.cmask.bernoulli_(keep).div_(keep)
        if labels is not None:
            p.labels.copy_(labels.reshape(-1).to(torch.int32), non_blocking=True)

    def oracle_shape(self, name):
        """Shape of a parameter in the oracle / PyTorch layout (conv [Cout,Cin,kh,kw])."""
        s = self.segs[name].shape
        if len(s) == 4:
            if name.endswith("/a1/w"):
                return (self.true_co[name], s[3] // 25, 5, 5)
            return (self.true_co[name], s[3], s[1], s[2])
        return tuple(s)

    # ------------------------------------------------------------------ forward
    def _conv(self, b, bn, name):
        src, dst, pool = _GS_WIRING[name]
        check(lib.ugn_conv2d_fwd(self.ctx.h, b.R[src].ptr, self.Rcw[f"{bn}/{name}/w"].ptr, None, b.R[dst].ptr,
                                 b.R[f"idx_{dst}"].ptr if pool else None, ACT_LEAKY, GS_ALPHA, pool, stream_ptr()))

    def _pad(self, b, src, dst):
        check(lib.ugn_pad_hw(self.ctx.h, b.R[src].ptr, b.R[dst].ptr, stream_ptr()))

    def _forward_branch(self, p, m: int, train: bool, expanded: bool):
        h, T = self.ctx.h, self.cfg.frames
        bn, b = BRANCH_NAMES[m], p.br[m]
        R = b.R
        st = stream_ptr()
        check(lib.ugn_gs_conv1_fwd(h, R["x_in"].ptr, self.Rw[f"{bn}/a1/w"].ptr, R["a1p"].ptr, GS_ALPHA, st))
        self._conv(b, bn, "a2"); self._pad(b, "a2", "a2p")
        check(lib.ugn_setmax_fwd(h, R["a2_set"].ptr, T, None, R["m0"].ptr, R["g0_set"].ptr, st))
        self._pad(b, "g0", "g0p")
        self._conv(b, bn, "b1"); self._pad(b, "b1", "b1p")
        self._conv(b, bn, "b2")
        self._conv(b, bn, "a3"); self._pad(b, "a3", "a3p")
        self._conv(b, bn, "a4"); self._pad(b, "a4", "a4p")
        check(lib.ugn_setmax_fwd(h, R["a4"].ptr, T, R["b2"].ptr, R["m1"].ptr, R["s1"].ptr, st))
        self._pad(b, "s1", "s1p")
        self._conv(b, bn, "b3"); self._pad(b, "b3", "b3p")
        self._conv(b, bn, "b4")
        self._conv(b, bn, "a5"); self._pad(b, "a5", "a5p")
        self._conv(b, bn, "a6")
        check(lib.ugn_setmax_fwd(h, R["a6"].ptr, T, R["b4"].ptr, R["m2"].ptr, R["s2"].ptr, st))
        check(lib.ugn_hpp_fwd(h, R["m2"].ptr, 0, R["feat"].ptr, st))
        check(lib.ugn_hpp_fwd(h, R["s2"].ptr, 1, R["feat"].ptr, st))
        check(lib.ugn_bmm_f32(h, R["feat"].ptr, 0, self.Rw[f"{bn}/matmul/w"].ptr, 0, R["out"].ptr, st))

    def _forward(self, p, train: bool, expanded: bool = False):
        cfg, h = self.cfg, self.ctx.h
        check(lib.ugn_set_fwd_passes(h, *self.fwd_passes))
        streams = self._fork()
        for m in range(cfg.nmods):
            with torch.cuda.stream(streams[m] if streams else torch.cuda.current_stream()):
                self._forward_branch(p, m, train, expanded)
        self._join(streams)
        st = stream_ptr()
        if not cfg.single:      # 1-modality graph (UWYHSemiNet.build, :890-905): the branch output IS the signature
            check(lib.ugn_fuse3_fwd(h, cfg.nmods, p.br_ptrs, p.flag_ptrs, p.R["sig"].ptr, p.R["winner"].ptr,
                                    p.R["col_norm"].ptr, cfg.merge | (FUSE3_NO_NORM if self.post2 else 0), st))
        sig = p.R["sig"]
        feat = p.R["sig2d"]
        if self.post2:
            # postriplet == 2 (:819-832): un-normalised fusion -> Dense(nc, activation=None, activity_regularizer)
            # "signature" -> LeakyReLU -> l2_normalize(axis=1) "code" (= the embedding) -> Dropout "dropcode" -> classifier
            check(lib.ugn_linear_fwd(h, p.R["sig2d"].ptr, self.Rw["code/w"].ptr, self.Rw["code/b"].ptr, None,
                                     p.R["code_lin"].ptr, None, ACT_LINEAR, 0.0, st))
            check(lib.ugn_linear_fwd(h, p.R["sig2d"].ptr, self.Rw["code/w"].ptr, self.Rw["code/b"].ptr, None,
                                     p.R["code"].ptr, None, ACT_LEAKY, cfg.alpha, st))
            check(lib.ugn_fuse3_fwd(h, 1, p.code_ptrs, p.one_ptrs, p.R["codeN"].ptr, p.R["cwin"].ptr, p.R["ccol"].ptr,
                                    MERGE_MAX, st))
            if p.dropN is not p.codeN:
                torch.mul(p.codeN, p.cmask.view_as(p.codeN), out=p.dropN)
            if cfg.nclasses > 0:
                check(lib.ugn_permute102(h, p.R["dropN"].ptr, p.R["flat"].ptr, st))
                check(lib.ugn_linear_fwd(h, p.R["flat"].ptr, self.Rw["classprob/w"].ptr, self.Rw["classprob/b"].ptr, None,
                                         p.R["logits"].ptr, None, ACT_LINEAR, 0.0, st))
            return p.R["codeN"], p.R["dropN"]
        if cfg.nc > 0:
            # Dense(activation=None, activity_regularizer) -> LeakyReLU(alpha) -> Dropout (:1199-1203): the linear
            # output is kept for the regulariser, the activated (and, in training, dropped) one feeds FC2
            cmask = p.R["cmask"].ptr if (train and cfg.dropout > 0.001) else None
            if train:
                check(lib.ugn_linear_fwd(h, p.R["sig2d"].ptr, self.Rw["code/w"].ptr, self.Rw["code/b"].ptr, None,
                                         p.R["code_lin"].ptr, None, ACT_LINEAR, 0.0, st))
            check(lib.ugn_linear_fwd(h, p.R["sig2d"].ptr, self.Rw["code/w"].ptr, self.Rw["code/b"].ptr, cmask,
                                     p.R["code"].ptr, None, ACT_LEAKY, cfg.alpha, st))
            feat = p.R["code"]
        if cfg.nclasses > 0:
            check(lib.ugn_permute102(h, (p.R["code3d"] if cfg.nc > 0 else p.R["sig"]).ptr, p.R["flat"].ptr, st))
            check(lib.ugn_linear_fwd(h, p.R["flat"].ptr, self.Rw["classprob/w"].ptr, self.Rw["classprob/b"].ptr, None,
                                     p.R["logits"].ptr, None, ACT_LINEAR, 0.0, st))
        return sig, feat

    @torch.no_grad()
    def predict(self, inputs: Sequence[torch.Tensor], flags: Optional[Sequence[torch.Tensor]] = None,
                layer: str = "signature") -> torch.Tensor:
        """model_code.predict (mains/mj_testUWYHGaitNet_open_tum.py:139-148): layer in {signature [62,B,d],
        flatten [B,62*d] (typecode 3), code, classprob}."""
        B = int(inputs[0].shape[0])
        p = self.plan(B, False)
        self._set_inputs(p, inputs, flags)
        self._forward(p, False)
        return self._layer_output(p, layer)

    def _layer_output(self, p, layer: str) -> torch.Tensor:
        B = p.B
        if self.post2:          # "signature" is the (linear) Dense layer, "code" its normalised LeakyReLU: the model's output 0
            if layer == "signature":
                return p.code_lin.view(GS_PARTS, B, -1).clone()
            if layer in ("code", "embedding"):
                return p.codeN.clone()
            if layer == "flatten":
                return p.codeN.permute(1, 0, 2).reshape(B, -1).clone()
        if layer in ("signature", "embedding"):
            return p.sig.clone()
        if layer == "flatten":
            return (p.code3d if self.cfg.nc > 0 else p.sig).permute(1, 0, 2).reshape(B, -1).clone()
        if layer == "code":
            return p.code3d.clone()
        if layer in ("classprob", "logits"):
            return p.logits.clone()
        raise KeyError(layer)

    @torch.no_grad()
    def predict_prefetched(self, layer: str = "signature") -> torch.Tensor:
        p, hb = self._consume_prefetched()
        self._forward(p, False)
        return self._layer_output(p, layer)

    # ------------------------------------------------------------------ backward
    def _losses_and_backward(self, p, sig: TRef, feat: TRef):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        B = p.B
        self._works = []
        if self.post2:
            self._post2_losses_and_head_backward(p)
        else:
            self._losses_and_head_backward(p, sig)
        self._reduce_bucket("heads")
        if not cfg.single:      # (1-modality graph: dsig is the branch's output gradient buffer itself)
            check(lib.ugn_fuse3_bwd(h, cfg.nmods, p.R["dsig"].ptr, p.R["sig"].ptr, p.R["winner"].ptr, p.R["col_norm"].ptr,
                                    p.flag_ptrs, p.dbr_ptrs, cfg.merge | (FUSE3_NO_NORM if self.post2 else 0), st))
        self._backward_branches(p)

    def _post2_losses_and_head_backward(self, p):
        """postriplet == 2: triplet + CE on the normalised code, back through Dropout, l2_normalize(axis=1), LeakyReLU, the
        activity regulariser of the linear Dense output (divided by shape(output)[0] = 62) and the Dense layer, into dsig =
        the gradient of the un-normalised fusion."""
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        check(lib.ugn_triplet_all(h, p.R["codeN"].ptr, p.R["labels"].ptr, cfg.margin, cfg.wver, p.R["trip_out"].ptr,
                                  p.R["dcodeN"].ptr, p.R["trip_ws"].ptr, st))
        if cfg.nclasses > 0:
            check(lib.ugn_softmax_ce_ls(h, p.R["logits"].ptr, p.R["labels"].ptr, p.R["ce_out"].ptr, p.R["dlogits"].ptr,
                                        cfg.wid, cfg.label_smoothing, st))
            check(lib.ugn_linear_bwd(h, p.R["flat"].ptr, self.Rw["classprob/w"].ptr, p.R["dlogits"].ptr,
                                     p.R["dflat"].ptr, self.Rg["classprob/w"].ptr, self.Rg["classprob/b"].ptr, st))
            check(lib.ugn_permute102(h, p.R["dflat3d"].ptr, p.R["dfeat"].ptr, st))        # [B,62,nc] -> [62,B,nc]
            if p.dropN is not p.codeN:
                p.dfeat.mul_(p.cmask.view_as(p.dfeat))
            p.dcodeN.add_(p.dfeat)
        check(lib.ugn_fuse3_bwd(h, 1, p.R["dcodeN"].ptr, p.R["codeN"].ptr, p.R["cwin"].ptr, p.R["ccol"].ptr, p.one_ptrs,
                                p.dcode_ptrs, MERGE_MAX, st))
        check(lib.ugn_act_mask_bwd(h, p.R["dcode"].ptr, p.R["code"].ptr, None, p.R["dcode_z"].ptr, None, ACT_LEAKY,
                                   cfg.alpha, st))
        p.dcode_z.add_(p.code_lin, alpha=2e-3 / GS_PARTS)
        check(lib.ugn_linear_bwd(h, p.R["sig2d"].ptr, self.Rw["code/w"].ptr, p.R["dcode_z"].ptr, p.R["dsig2d"].ptr,
                                 self.Rg["code/w"].ptr, self.Rg["code/b"].ptr, st))

    def _losses_and_head_backward(self, p, sig: TRef):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        check(lib.ugn_triplet_all(h, sig.ptr, p.R["labels"].ptr, cfg.margin, cfg.wver, p.R["trip_out"].ptr,
                                  p.R["dsig"].ptr, p.R["trip_ws"].ptr, st))
        if cfg.nclasses > 0:
            check(lib.ugn_softmax_ce_ls(h, p.R["logits"].ptr, p.R["labels"].ptr, p.R["ce_out"].ptr, p.R["dlogits"].ptr,
                                        cfg.wid, cfg.label_smoothing, st))
            check(lib.ugn_linear_bwd(h, p.R["flat"].ptr, self.Rw["classprob/w"].ptr, p.R["dlogits"].ptr,
                                     p.R["dflat"].ptr, self.Rg["classprob/w"].ptr, self.Rg["classprob/b"].ptr, st))
            check(lib.ugn_permute102(h, p.R["dflat3d"].ptr, p.R["dfeat"].ptr, st))        # [B,62,f] -> [62,B,f]
            if cfg.nc > 0:
                # through LeakyReLU, then the activity regulariser l2(1e-3) of the LINEAR "code" output,
                # divided by shape(output)[0] = 62 in this layout (:1199-1200)
                check(lib.ugn_act_mask_bwd(h, p.R["dfeat2d"].ptr, p.R["code"].ptr,
                                           p.R["cmask"].ptr if cfg.dropout > 0.001 else None, p.R["dcode_z"].ptr, None,
                                           ACT_LEAKY, cfg.alpha, st))
                p.dcode_z.add_(p.code_lin, alpha=2e-3 / GS_PARTS)
                check(lib.ugn_linear_bwd(h, p.R["sig2d"].ptr, self.Rw["code/w"].ptr, p.R["dcode_z"].ptr,
                                         p.R["dsig2"].ptr, self.Rg["code/w"].ptr, self.Rg["code/b"].ptr, st))
                p.dsig.add_(p.dsig2.view_as(p.dsig))
            else:
                p.dsig.add_(p.dfeat)

    def _backward_branches(self, p):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        # MatMul + HPP backward stay in f32; the conv stacks below consume 16-bit gradient operands
        streams = self._fork() if self._branches_concurrent() else None
        for m in range(cfg.nmods):
            with torch.cuda.stream(streams[m] if streams else torch.cuda.current_stream()):
                self._backward_head(p, m)
        self._join(streams)
        if self.scaled:
            # fp16 gradient operands: this step's power-of-two scale comes from the gradients that ENTER the
            # conv stacks (d_m2 / d_b4 of every branch, one contiguous buffer).  dL/dsignature, the stacked-CNN
            # engine's reference, is 64x smaller here: the batch-axis l2_normalize and the 5-level pyramid
            # amplify the gradient on the way back, and the conv-stack gradients only shrink from there
            check(lib.ugn_grad_scale_update(h, p.R["dhead"].ptr, GRAD_SCALE_TARGET, st))
            self.ctx.grad_scaled = True
        elif getattr(self.ctx, "grad_scaled", False):
            check(lib.ugn_grad_scale_set(h, 1.0, st))
            self.ctx.grad_scaled = False
        streams = self._fork() if self._branches_concurrent() else None
        for m in range(cfg.nmods):
            with torch.cuda.stream(streams[m] if streams else torch.cuda.current_stream()):
                self._backward_branch(p, m)
        self._join(streams)

    def _bwd_conv(self, b, bn, name, crop: bool = True):
        """dy (f32, layer output gradient d_<dst>) -> dz; weight gradient; input gradient cropped into d_<src>."""
        h, st, R = self.ctx.h, stream_ptr(), b.R
        src, dst, pool = _GS_WIRING[name]
        check(lib.ugn_conv2d_bwd_act(h, R[f"d_{dst}"].ptr, R[dst].ptr, R[f"idx_{dst}"].ptr if pool else None,
                                     R[f"dz_{dst}"].ptr, None, ACT_LEAKY, GS_ALPHA, pool, st))
        check(lib.ugn_conv2d_wgrad(h, R[src].ptr, R[f"dz_{dst}"].ptr, self.Rg[f"{bn}/{name}/w"].ptr, None, st))
        check(lib.ugn_conv2d_dgrad(h, R[f"dz_{dst}"].ptr, self.Rcw[f"{bn}/{name}/w"].ptr, R[f"dx_{src}"].ptr, st))
        if crop:
            check(lib.ugn_crop_hw(h, R[f"dx_{src}"].ptr, R[f"d_{src[:-1]}"].ptr, 0, st))

    def _backward_head(self, p, m: int):
        h, st = self.ctx.h, stream_ptr()
        bn, b = BRANCH_NAMES[m], p.br[m]
        R = b.R
        # MatMul
        check(lib.ugn_bmm_f32(h, R["dout"].ptr, 0, self.Rw[f"{bn}/matmul/w"].ptr, 1, R["dfeat"].ptr, st))
        check(lib.ugn_bmm_f32(h, R["feat"].ptr, 1, R["dout"].ptr, 0, self.Rg[f"{bn}/matmul/w"].ptr, st))
        # HPP: d_b4 = d(s2); d(m2) = d(s2) + set-level strips
        check(lib.ugn_hpp_bwd(h, R["dfeat"].ptr, R["s2"].ptr, 1, R["d_b4"].ptr, 0, st))
        b.d_m2.copy_(b.d_b4)
        check(lib.ugn_hpp_bwd(h, R["dfeat"].ptr, R["m2"].ptr, 0, R["d_m2"].ptr, 1, st))

    def _backward_branch(self, p, m: int):
        h, st, T = self.ctx.h, stream_ptr(), self.cfg.frames
        bn, b = BRANCH_NAMES[m], p.br[m]
        R = b.R
        # global branch
        self._bwd_conv(b, bn, "b4")          # -> d_b3
        self._bwd_conv(b, bn, "b3")          # -> d_s1  (== d_b2 and the gradient of the set pooling of a4)
        b.d_b2.copy_(b.d_s1)
        self._bwd_conv(b, bn, "b2")          # -> d_b1
        self._bwd_conv(b, bn, "b1")          # -> d_g0
        # set branch
        check(lib.ugn_setmax_bwd(h, R["d_m2"].ptr, R["a6"].ptr, R["m2"].ptr, T, R["d_a6"].ptr, 0, st))
        self._bwd_conv(b, bn, "a6")          # -> d_a5
        self._bwd_conv(b, bn, "a5")          # -> d_a4
        check(lib.ugn_setmax_bwd(h, R["d_s1"].ptr, R["a4"].ptr, R["m1"].ptr, T, R["d_a4"].ptr, 1, st))
        self._bwd_conv(b, bn, "a4")          # -> d_a3
        self._bwd_conv(b, bn, "a3")          # -> d_a2
        check(lib.ugn_setmax_bwd(h, R["d_g0_set"].ptr, R["a2_set"].ptr, R["m0"].ptr, T, R["d_a2_set"].ptr, 1, st))
        self._bwd_conv(b, bn, "a2", crop=False)          # -> dx_a1p (gradient wrt a1 on the padded, split frame)
        # first convolution: LeakyReLU derivative + crop + kernel gradient in one fp32 kernel
        check(lib.ugn_gs_conv1_wgrad(h, R["x_in"].ptr, R["dx_a1p"].ptr, R["a1p"].ptr, self.Rg[f"{bn}/a1/w"].ptr,
                                     GS_ALPHA, st))
        self._reduce_bucket(m)

    def _report(self, p, with_reg: bool = False) -> Dict[str, torch.Tensor]:
        out = {"triplet": p.trip_out[0], "count": p.trip_out[1], "signature": p.codeN if self.post2 else p.sig}
        if self.cfg.nclasses > 0:
            out["ce"], out["acc"], out["logits"] = p.ce_out[0], p.ce_out[1], p.logits
        if with_reg:
            out["reg"] = p.loss_pack[4]
            out["losses"] = p.loss_pack
        return out

    def train_step_expanded(self, *a, **k):
        raise NotImplementedError("device-side expansion is implemented for the stacked-CNN engine only")

    predict_expanded = train_step_expanded


class _Branch:
    pass


class _GsPlan:
    """All activation / gradient buffers of one batch size of the gaitset graph."""

    def __init__(self, eng: GaitSetEngine, B: int, train: bool):
        cfg, d, P, PB, dt16 = eng.cfg, eng.dev, eng.P, eng.PB, eng.dt16
        self.B, self.train, self.eng = B, train, eng
        self.use_mirror = False
        f32 = dict(device=d, dtype=torch.float32)
        T = cfg.frames
        F = B * T
        Hs = cfg.hw + 4
        H1, H2 = Hs // 2, Hs // 4
        assert Hs % 4 == 0 and (H2 * H2) % 16 == 0, "gaitset: (hw+4) must be a multiple of 4 and the last map of 16 positions"
        c2 = 32
        S = eng.split_for(Hs)

        def act(shape, planes=P):
            if planes:
                return torch.zeros((planes,) + tuple(shape), device=d, dtype=dt16)
            return torch.zeros(tuple(shape), **f32)

        self.br: List[_Branch] = []
        from .net import IOBlock
        self.io = IOBlock(d, B, [(T, cfg.hw, cfg.hw, c) for c in cfg.in_channels])     # one H2D copy per step
        iov = self.io.views(self.io.dev_buf)
        self.flags = iov["flags"]
        for f in self.flags:
            f.fill_(1.0)
        dhead = torch.zeros(cfg.nmods, 2, B, H2, H2, 128, **f32) if train else None
        # (name, N, H, C): activations in storage mode; "<name>p" = zero-bordered copy
        for m in range(cfg.nmods):
            b = _Branch()
            c = cfg.in_channels[m]
            Tn = {}
            b.x_in = Tn["x_in"] = iov["x"][m]
            # name -> (N, H, W, C); "a1p" / "a2" / "g0" live in the S-way column-split layout
            geo = {"a2": (F * S, H1, H1 // S, c2), "a3": (F, H1, H1, 64), "a4": (F, H2, H2, 64), "a5": (F, H2, H2, 128),
                   "a6": (F, H2, H2, 128), "g0": (B * S, H1, H1 // S, c2), "b1": (B, H1, H1, 64), "b2": (B, H2, H2, 64),
                   "s1": (B, H2, H2, 64), "b3": (B, H2, H2, 128), "b4": (B, H2, H2, 128)}
            padded = {"a1p": (F * S, Hs + 2, Hs // S + 2, 32)}
            for name, (n, hh, ww, cc) in geo.items():
                if name not in ("a6", "b4", "b2"):
                    ns, wfull = (n // S, ww * S) if name in ("a2", "g0") else (n, ww)
                    padded[name + "p"] = (ns, hh + 2, wfull + 2, cc)
            for name, (n, hh, ww, cc) in geo.items():
                Tn[name] = act((n, hh, ww, cc))
                if name in ("a2", "a4", "b2"):
                    Tn["idx_" + name] = torch.zeros(n, hh, ww, cc, device=d, dtype=torch.uint8)
            for name, shp in padded.items():
                Tn[name] = act(shp)
            # set pooling sees one frame as one contiguous block: [F, S*H1, H1/S, C] views of the split tensors
            pl = (P,) if P else ()
            Tn["a2_set"] = Tn["a2"].view(pl + (F, S * H1, H1 // S, c2))
            Tn["g0_set"] = Tn["g0"].view(pl + (B, S * H1, H1 // S, c2))
            Tn["m0"] = torch.zeros(B, S * H1, H1 // S, c2, **f32)
            Tn["m1"] = torch.zeros(B, H2, H2, 64, **f32)
            Tn["m2"] = torch.zeros(B, H2, H2, 128, **f32)
            Tn["s2"] = torch.zeros(B, H2, H2, 128, **f32)
            Tn["feat"] = torch.zeros(GS_PARTS, B, 128, **f32)
            b.out = Tn["out"] = torch.zeros(GS_PARTS, B, cfg.hidden, **f32)
            if train:
                b.dout = Tn["dout"] = torch.zeros(GS_PARTS, B, cfg.hidden, **f32)
                Tn["dfeat"] = torch.zeros(GS_PARTS, B, 128, **f32)
                for name, (n, hh, ww, cc) in geo.items():
                    Tn["d_" + name] = dhead[m, 1] if name == "b4" else torch.zeros(n, hh, ww, cc, **f32)   # d(layer output)
                    if name in _DZ_GEO:
                        Tn["dz_" + name] = act((n, hh * _DZ_GEO[name], ww * _DZ_GEO[name], cc), PB)
                for name, shp in padded.items():
                    Tn["dx_" + name] = torch.zeros(shp, **f32)
                Tn["d_a2_set"] = Tn["d_a2"].view(F, S * H1, H1 // S, c2)
                Tn["d_g0_set"] = Tn["d_g0"].view(B, S * H1, H1 // S, c2)
                b.d_m2 = Tn["d_m2"] = dhead[m, 0]
                b.d_b4, b.d_b2, b.d_s1 = Tn["d_b4"], Tn["d_b2"], Tn["d_s1"]
            b.T = Tn
            b.R = {k: TRef(v) for k, v in Tn.items()}
            self.br.append(b)
        Tn = {}
        nd = cfg.hidden
        # 1-modality graph (cfg.single): no gate / fusion / l2_normalize -- the signature and its gradient are the
        # branch's own output / output-gradient buffers
        self.sig = Tn["sig"] = self.br[0].out if cfg.single else torch.zeros(GS_PARTS, B, nd, **f32)
        Tn["sig2d"] = self.sig.view(GS_PARTS * B, nd)
        Tn["winner"] = torch.zeros(GS_PARTS, B, nd, device=d, dtype=torch.uint8)
        Tn["col_norm"] = torch.zeros(GS_PARTS, nd, 2, **f32)
        feat = nd
        if cfg.nc > 0:
            self.code_lin = Tn["code_lin"] = torch.zeros(GS_PARTS * B, cfg.nc, **f32)
            Tn["code"] = torch.zeros(GS_PARTS * B, cfg.nc, **f32)
            self.code3d = Tn["code3d"] = Tn["code"].view(GS_PARTS, B, cfg.nc)
            self.cmask = Tn["cmask"] = torch.ones(GS_PARTS * B, cfg.nc, **f32)
            feat = cfg.nc
        if eng.post2:           # postriplet == 2: the normalised code [62,B,nc] is the embedding (see GaitSetEngine._forward)
            self.codeN = Tn["codeN"] = torch.zeros(GS_PARTS, B, cfg.nc, **f32)
            Tn["cwin"] = torch.zeros(GS_PARTS, B, cfg.nc, device=d, dtype=torch.uint8)
            Tn["ccol"] = torch.zeros(GS_PARTS, cfg.nc, 2, **f32)
            drop = train and cfg.dropout > 0.001
            self.dropN = Tn["dropN"] = torch.zeros(GS_PARTS, B, cfg.nc, **f32) if drop else self.codeN
            self.ones = Tn["ones"] = torch.ones(B, 1, **f32)
        if cfg.nclasses > 0:
            Tn["flat"] = torch.zeros(B, GS_PARTS * feat, **f32)
            self.logits = Tn["logits"] = torch.zeros(B, cfg.nclasses, **f32)
        self.labels = Tn["labels"] = iov["labels"]
        if train:
            self.loss_pack = Tn["loss_pack"] = torch.zeros(8, **f32)     # {triplet, count, ce, acc, reg}
            self.trip_out = Tn["trip_out"] = self.loss_pack[0:2]
            self.dsig = Tn["dsig"] = self.br[0].dout if cfg.single else torch.zeros(GS_PARTS, B, nd, **f32)
            Tn["trip_ws"] = torch.zeros(ops.triplet_workspace_bytes(GS_PARTS, B) // 4 + 16, **f32)
            if cfg.nclasses > 0:
                self.ce_out = Tn["ce_out"] = self.loss_pack[2:4]
                Tn["dlogits"] = torch.zeros(B, cfg.nclasses, **f32)
                Tn["dflat"] = torch.zeros(B, GS_PARTS * feat, **f32)
                Tn["dflat3d"] = Tn["dflat"].view(B, GS_PARTS, feat)
                self.dfeat = Tn["dfeat"] = torch.zeros(GS_PARTS, B, feat, **f32)
                Tn["dfeat2d"] = self.dfeat.view(GS_PARTS * B, feat)
            if cfg.nc > 0:
                self.dcode_z = Tn["dcode_z"] = torch.zeros(GS_PARTS * B, cfg.nc, **f32)
                self.dsig2 = Tn["dsig2"] = torch.zeros(GS_PARTS * B, nd, **f32)
            if eng.post2:
                self.dcodeN = Tn["dcodeN"] = torch.zeros(GS_PARTS, B, cfg.nc, **f32)
                Tn["dcode3d"] = torch.zeros(GS_PARTS, B, cfg.nc, **f32)
                Tn["dcode"] = Tn["dcode3d"].view(GS_PARTS * B, cfg.nc)
                Tn["dsig2d"] = self.dsig.view(GS_PARTS * B, nd)
                if cfg.nclasses == 0:
                    self.dfeat = None
        if train:
            Tn["dhead"] = dhead.view(-1)
        self.T = Tn
        self.R = {k: TRef(v) for k, v in Tn.items()}
        self.R_flags = [TRef(f) for f in self.flags]
        self.br_ptrs = ptr_array([b.R["out"] for b in self.br])
        self.flag_ptrs = ptr_array(self.R_flags)
        if train:
            self.dbr_ptrs = ptr_array([b.R["dout"] for b in self.br])
        if eng.post2:
            self.code_ptrs, self.one_ptrs = ptr_array([self.R["code3d"]]), ptr_array([self.R["ones"]])
            if train:
                self.dcode_ptrs = ptr_array([self.R["dcode3d"]])


# conv outputs that have a pre-activation gradient buffer: name -> spatial factor (2 = the layer is pooled,
# dz lives on the pre-pool grid)
_DZ_GEO = {"a2": 2, "a3": 1, "a4": 2, "a5": 1, "a6": 1, "b1": 1, "b2": 2, "b3": 1, "b4": 1}
