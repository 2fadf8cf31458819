"""UGaitNet step engine: host-side orchestration of the C-ABI kernels.

One :class:`UGaitEngine` owns, for one GPU:
  * a flat f32 parameter arena (+ gradient, Adam m/v arenas of the same layout) whose
    segments are the trainable tensors of the reference graph
    (/root/reference/nets/mj_uwyhNets_ba.py:67-107, :1163-1214),
  * per-batch-size activation plans (all buffers preallocated, exported once via DLPack),
  * the forward / backward / optimiser schedule of one training step, optionally captured
    in a CUDA graph.

PyTorch is used for memory, streams, RNG (dropout masks) and NCCL only; every arithmetic op
of the step is a kernel of libugaitnet_b200.so.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from ._ffi import TRef, check, lib, ptr_array, stream_ptr
from .config import ACT_LEAKY, ACT_LINEAR, AUX_NAMES, BRANCH_NAMES, NetConfig, round_up
from .expand import NOISE

# math mode -> (forward planes, backward/gradient planes, 16-bit dtype)
#   fp32    SIMT FFMA validation path
#   bf16    single-pass bf16 (fast, NOT parity-passing)
#   bf16x3  bf16 hi/lo split, 3 MMA passes forward and backward
#   f16x3   fp16 hi/lo split, 3 passes forward and backward (22-bit operands)
#   f16mix  fp16 hi/lo split 3-pass FORWARD (every discrete decision -- ReLU, max-pool argmax, sign_max
#           winner -- is taken at ~fp32 accuracy), single-pass fp16 BACKWARD on the hi planes with a
#           device-side power-of-two gradient scale (linear maps of the gradient need 11 bits, not 22)
#   f16mix2 as f16mix with TWO forward passes: convolutions hi*hi + hi*lo(weights) (activation lo plane not loaded, the
#           pair issued as one N = 2*Cout MMA where Cout <= 128), dense layers hi*hi + lo(activations)*hi (the weight lo
#           plane is not streamed)
#   f16mix1 single-pass fp16 forward and backward (11-bit operands: the "TF32 class" north_star names)
MATH_MODES = {"fp32": (0, 0, None), "bf16": (1, 1, torch.bfloat16), "bf16x3": (2, 2, torch.bfloat16),
              "f16x3": (2, 2, torch.float16), "f16mix": (2, 1, torch.float16), "f16mix2": (2, 1, torch.float16),
              "f16mix1": (2, 1, torch.float16)}
# (conv, dense) pass codes of ugn_set_fwd_passes per math mode (default 0 = all three)
FWD_PASSES = {"f16mix2": (2, 4), "f16mix1": (1, 1)}
# optical-flow magnitude clip of the generator (__load_dd, data/mj_dataGeneratorMMUWYHsingle.py:318-324), on decoded
# values: raw int16 / compressFactor (100) * 0.1 -> thresholds 2300 / 50 become 2.3 / 0.05, the fill 1e-8 becomes 1e-11
OF_CLIP_LO, OF_CLIP_HI, OF_CLIP_VAL = 0.05, 2.3, 1e-11
GRAD_SCALE_TARGET = 1024.0      # max|dL/dsignature| * s lands in [512, 1024]: 6 binades of headroom below 65504


class IOBlock:
    """Every per-step INPUT of one plan in ONE contiguous device block, mirrored byte for byte by pinned host blocks,
    so that a step's inputs cross PCIe in a single cudaMemcpyAsync (one DMA descriptor instead of 2M+1 small copies).

    Layout (256-byte aligned fields): flags[M] f32 [B,1] | labels i32 [B] | src_row i32 [B] | mirror u8 [B] | volumes.
    The volume area is viewed in two ways that never coexist inside one step: ``x[m]`` = the full batch [B,...], and
    ``xb(B0)[m]`` = the B0 base rows of the device-side expansion packed back to back right behind the header, so
    the bytes a step has to copy are always one prefix ``[0, nbytes_full)`` / ``[0, nbytes_base(B0))`` of the block."""

    ALIGN = 256

    def __init__(self, dev, B: int, vol_shapes, has_flags: bool = True):
        self.B, self.vol_shapes, self.M = B, [tuple(s) for s in vol_shapes], len(vol_shapes)
        off = 0
        self.f_off = []
        for _ in range(self.M if has_flags else 0):
            self.f_off.append(off)
            off = round_up(off + 4 * B, self.ALIGN)
        self.has_flags = has_flags
        self.lab_off = off
        off = round_up(off + 4 * B, self.ALIGN)
        self.src_off = off
        off = round_up(off + 4 * B, self.ALIGN)
        self.mir_off = off
        off = round_up(off + B, self.ALIGN)
        self.shift_off = off                      # (tx, ty) i8 per row, OF-clip flag u8 per row: device-side augmentation
        off = round_up(off + 2 * B, self.ALIGN)
        self.clip_off = off
        off = round_up(off + B, self.ALIGN)
        self.header = off
        self.row_bytes = [4 * int(torch.tensor(s).prod()) for s in self.vol_shapes]
        self.x_off, o = [], off
        for rb in self.row_bytes:
            self.x_off.append(o)
            o = round_up(o + rb * B, self.ALIGN)
        self.nbytes_full = o
        self.dev_buf = torch.zeros(o, dtype=torch.uint8, device=dev)
        self._host = {}

    def base_offsets(self, B0: int):
        offs, o = [], self.header
        for rb in self.row_bytes:
            offs.append(o)
            o = round_up(o + rb * B0, self.ALIGN)
        return offs, o

    def nbytes_base(self, B0: int) -> int:
        return self.base_offsets(B0)[1]

    def raw_offsets(self, rows: int, itemsizes):
        """Volume offsets of a RAW block (stored integers instead of f32: HostBatch(raw=...)), `rows` rows per modality."""
        offs, o = [], self.header
        for rb, isz in zip(self.row_bytes, itemsizes):
            offs.append(o)
            o = round_up(o + rb // 4 * isz * rows, self.ALIGN)
        return offs, o

    @staticmethod
    def _view(buf, off, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n * torch.empty(0, dtype=dtype).element_size()].view(dtype).view(tuple(shape))

    def views(self, buf, B0: Optional[int] = None, raw=None):
        """Typed views of a buffer with this layout (device block, staging block or pinned host block).  raw: the
        volumes are stored integers (see HostBatch)."""
        B = self.B
        v = {"flags": [self._view(buf, o, (B, 1), torch.float32) for o in self.f_off],
             "labels": self._view(buf, self.lab_off, (B,), torch.int32),
             "src_row": self._view(buf, self.src_off, (B,), torch.int32),
             "mirror": self._view(buf, self.mir_off, (B,), torch.uint8),
             "shift": self._view(buf, self.shift_off, (B, 2), torch.int8),
             "clip": self._view(buf, self.clip_off, (B,), torch.uint8)}
        if raw is not None:
            rows = B if B0 is None else B0
            offs, _ = self.raw_offsets(rows, [2 if r[0] is torch.int16 else 1 for r in raw])
            v["x"] = [self._view(buf, o, (rows,) + s, r[0]) for o, s, r in zip(offs, self.vol_shapes, raw)]
        elif B0 is None:
            v["x"] = [self._view(buf, o, (B,) + s, torch.float32) for o, s in zip(self.x_off, self.vol_shapes)]
        else:
            offs, _ = self.base_offsets(B0)
            v["x"] = [self._view(buf, o, (B0,) + s, torch.float32) for o, s in zip(offs, self.vol_shapes)]
        return v


class HostBatch:
    """Pinned host mirror of a plan's IOBlock: the data loader fills ``inputs[m]`` / ``flags[m]`` / ``labels`` (and
    ``src_row`` / ``mirror`` for the device-side expansion) IN PLACE (numpy views), ``UGaitEngine.prefetch_batch`` then
    moves the whole batch with one cudaMemcpyAsync."""

    def __init__(self, io: IOBlock, B0: Optional[int], raw=None):
        """raw: per modality (dtype "int16" | "uint8", divisor, mul, sub[, clip_min, clip_max]) -- the volumes are the
        STORED sample integers (ugaitnet_b200.samples.RAW_FLOW / RAW_GRAY / RAW_SILHOUETTE) and are decoded on the device
        (ugn_decode_samples) when the batch is consumed: a quarter / half of the f32 bytes cross PCIe."""
        self.io, self.B0, self.raw = io, B0, None
        if raw is not None:
            if len(raw) != io.M or any(r[0] not in ("int16", "uint8") or not 4 <= len(r) <= 6 for r in raw):
                raise ValueError(f"raw: one (\"int16\" | \"uint8\", divisor, mul, sub[, clip_min, clip_max]) per modality "
                                 f"({io.M}) expected, got {raw!r}")
            self.raw = [(torch.int16 if r[0] == "int16" else torch.uint8,) + tuple(float(v) for v in r[1:]) +
                        (0.0,) * (6 - len(r)) for r in raw]
            rows = io.B if B0 is None else B0
            self.raw_off, self.nbytes = io.raw_offsets(rows, [2 if r[0] is torch.int16 else 1 for r in self.raw])
            self.buf = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
        else:
            self.nbytes = io.nbytes_full if B0 is None else io.nbytes_base(B0)
            self.buf = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
        v = io.views(self.buf, B0, raw=self.raw)
        self.t = v                                       # torch views
        self.inputs = [x.numpy() for x in v["x"]]
        self.flags = [f.numpy() for f in v["flags"]]
        self.labels, self.src_row, self.mirror = v["labels"].numpy(), v["src_row"].numpy(), v["mirror"].numpy()
        self.shift, self.clip = v["shift"].numpy(), v["clip"].numpy()
        self.use_mirror = False
        self.use_augment = False          # True: shift / clip tables are applied (ugn_pack_input_augment)
        for f in self.flags:
            f[...] = 1.0


class _Seg:
    __slots__ = ("name", "shape", "off", "n", "l2")

    def __init__(self, name, shape, off, n, l2):
        self.name, self.shape, self.off, self.n, self.l2 = name, shape, off, n, l2


class UGaitEngine:
    def __init__(self, cfg: NetConfig, device: Optional[int] = None, math_mode: str = "fp32",
                 seed: int = 232323, optimizer: str = "adam", lr: float = 1e-4, momentum: float = 0.9,
                 beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7, process_group=None,
                 use_graph: bool = False, lr_decay: float = 0.0, decoupled_weight_decay: float = 0.0):
        if not torch.cuda.is_available():
            raise RuntimeError("ugaitnet_b200 needs a CUDA device (no CPU fallback)")
        assert math_mode in MATH_MODES
        self.cfg = cfg
        self.dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.ctx = ops.get_ctx(self.dev.index)
        self.math_mode = math_mode
        self.P, self.PB, self.dt16 = MATH_MODES[math_mode]
        self.fwd_passes = FWD_PASSES.get(math_mode, (0, 0))
        self.scaled = self.dt16 is torch.float16
        self.pad = 32 if self.P else 1
        if self.P and type(cfg) is NetConfig:
            bad = [c for c in cfg.filters_numbers if c % 32]
            if bad:     # the activation of layer i is the (unpadded) K operand of layer i + 1
                raise ValueError(f"math_mode={math_mode!r} (tensor cores) needs filter counts that are multiples of 32, got "
                                 f"{list(cfg.filters_numbers)}; use math_mode='fp32' for such a net")
        self.optimizer, self.lr, self.momentum = optimizer.lower(), float(lr), momentum
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        # optimizers.SGD(decay=...) -> lr / (1 + decay * iterations); tfa AdamW(weight_decay=...) -> decoupled decay
        self.lr_decay, self.decoupled_wd = float(lr_decay), float(decoupled_weight_decay)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.use_graph = use_graph
        self.t = 0
        self._dp_async = True      # bucketed, overlapped all-reduce (False: one blocking all-reduce)
        # data-parallel CUDA graphs: the step is captured as segments cut at the all-reduce points
        # (_capture_segments); NCCL itself is never captured (capturing the async work handles hung)
        self.dp_graph = os.environ.get("UGN_DP_GRAPH", "1") != "0"
        # data-parallel exchange.  "fused" (default; falls back to "split" where symmetric memory cannot be mapped):
        # see the end of this comment.  "split": the backward pass runs in two phases with concurrent branches
        # inside each -- dense layers first, then the convolution stacks -- and the all-reduce of the dense
        # gradients (92 % of the bytes) overlaps the second phase: 3 graph segments.  "single": forward + backward
        # are ONE segment, the whole arena is all-reduced in one call before the optimiser segment.  "bucketed":
        # one all-reduce per finished bucket with the branches in sequence (8 segments).  "fused": no all-reduce at
        # all -- ugn_dp_optim_step reduce-scatters the gradients, updates this rank's slice with its slice of the
        # optimiser state and all-gathers the weights in ONE kernel over NVLink peer memory
        self.dp_reduce = os.environ.get("UGN_DP_REDUCE", "fused")
        self.dp_onegraph = os.environ.get("UGN_DP_ONEGRAPH", "0") == "1"
        self._symm = []            # (tensor, symmetric-memory handle) of the exchanged arenas: [weights, gradients]
        self.multistream = os.environ.get("UGN_MULTISTREAM", "1") != "0"   # concurrent modality branches
        # dense-layer Adam issued right after the dense backward, on a side stream underneath the conv backward.
        # Opt-in: measured 3.79 ms vs 3.74 ms/step -- the persistent tcgen05 conv kernels own every SM (215 KB of
        # shared memory per CTA), so an HBM-bound kernel on another stream cannot co-reside and only interleaves
        self.early_optim = os.environ.get("UGN_EARLY_OPTIM", "0") == "1"
        self._early_active, self._early_done, self._opt_streams = False, [], None
        self._bstreams = None
        self.force_segments = False     # tests: use the segmented capture on a single GPU too
        self._works = []
        self._cap = None           # state of a segmented graph capture (data-parallel CUDA graphs)
        self.graph_launches = 0
        self._plans: Dict[tuple, "_Plan"] = {}
        self._graphs = {}
        # dropout masks from a counter-based generator inside the dense post passes (tensor-core modes): no mask tensors,
        # no torch RNG kernels in the step.  rng = {seed, step}; the step is advanced by a kernel at the start of every
        # training step (a captured CUDA graph therefore draws fresh masks on every replay)
        self.philox = bool(self.P) and os.environ.get("UGN_PHILOX", "1") == "1"
        self.rng_state = torch.tensor([int(seed) * 0x9E3779B97F4A7C15 % (1 << 62), 0], dtype=torch.int64, device=self.dev)
        self._R_rng = TRef(self.rng_state)
        self._build_arena()
        self.init_weights(seed)

    # ------------------------------------------------------------------ parameters
    def _new_arena(self, n: int, exchanged: bool = False) -> torch.Tensor:
        """Flat f32 arena.  The weight and gradient arenas of a data-parallel engine with dp_reduce == "fused" live in
        symmetric memory (torch.distributed._symmetric_memory: every rank maps every rank's buffer), which is what
        ugn_dp_optim_step reads gradients from and writes weights to over NVLink."""
        if exchanged and self.world > 1 and self.dp_reduce == "fused":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(n, dtype=torch.float32, device=self.dev)
                t.zero_()
                hdl = symm_mem.rendezvous(t, self.pg)
                ok = torch.ones(1, device=self.dev)
            except Exception as e:                       # no peer mapping on this box (P2P disabled, old driver ...)
                ok, err = torch.zeros(1, device=self.dev), e
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=self.pg)   # all ranks or none
            if float(ok) == 1.0:
                self._symm.append((t, hdl))
                return t
            import warnings
            warnings.warn("ugaitnet_b200: symmetric memory unavailable, data-parallel exchange falls back to NCCL "
                          f"all-reduce (UGN_DP_REDUCE=split): {locals().get('err', 'another rank failed')}")
            self.dp_reduce, self._symm = "split", []
        return torch.zeros(n, device=self.dev)

    def _build_arena(self):
        cfg = self.cfg
        segs: List[_Seg] = []
        off = 0

        def add(name, shape, l2=0.0):
            nonlocal off
            n = 1
            for s in shape:
                n *= s
            segs.append(_Seg(name, tuple(shape), off, n, l2))
            off += round_up(n, 64)

        if any(cfg.is3d(m) for m in range(cfg.nmods)) and self.P:
            raise ValueError("use3D (Conv3D branches) runs on the fp32 validation engine only: math_mode='fp32'")
        for m in range(cfg.nmods):
            bn = BRANCH_NAMES[m]
            if cfg.is3d(m):     # build_3Dbranch (:346-363): no kernel regulariser on the Conv3D layers, L2 1e-3 on "grayCode"
                for li, L in enumerate(cfg.layers3d(m)):
                    add(f"{bn}/conv{li}/w", (L["co"],) + tuple(L["k"]) + (L["cin"],))
                    add(f"{bn}/conv{li}/b", (L["co"],))
                add(f"{bn}/ofCode/w", (cfg.nd, cfg.filters3d[-1]), 1e-3)
                add(f"{bn}/ofCode/b", (cfg.nd,))
                continue
            for li, L in enumerate(cfg.layers(m, 1)):
                add(f"{bn}/conv{li}/w", (L["co"], L["k"], L["k"], L["cin"]), cfg.weight_decay)
                add(f"{bn}/conv{li}/b", (L["co"],))
            add(f"{bn}/dense/w", (2 * cfg.nd, cfg.flat))
            add(f"{bn}/dense/b", (2 * cfg.nd,))
            add(f"{bn}/ofCode/w", (cfg.nd, 2 * cfg.nd), 1e-3)
            add(f"{bn}/ofCode/b", (cfg.nd,))
        feat = cfg.nd
        if cfg.nc > 0:
            add("code/w", (cfg.nc, cfg.nd))
            add("code/b", (cfg.nc,))
            feat = cfg.nc
        if cfg.nclasses > 0:
            add("classprob/w", (cfg.nclasses, feat))
            add("classprob/b", (cfg.nclasses,))
        self.aux = bool(getattr(cfg, "aux_losses", False)) and cfg.nclasses > 0 and not cfg.single
        self.post2 = getattr(cfg, "postriplet", 1) == 2 and cfg.nc > 0 and not cfg.single
        if self.aux:       # classprob_{of,gray,depth}: Dense(nclasses, softmax) on every gated branch output (:1222-1229)
            for m in range(cfg.nmods):
                add(f"{AUX_NAMES[m]}/w", (cfg.nclasses, cfg.nd))
                add(f"{AUX_NAMES[m]}/b", (cfg.nclasses,))
        self.segs = {s.name: s for s in segs}
        self.seg_list = segs
        self.n_arena = off
        # all-reduce buckets: per branch one for the dense layers (92 % of the gradient bytes; complete
        # right after the two dense backward GEMMs, so their all-reduce overlaps the conv backward of the
        # same branch) and one for the conv layers (complete when the branch's backward ends); + the heads
        self.buckets = {}
        for m in range(cfg.nmods):
            mine = [s for s in segs if s.name.startswith(BRANCH_NAMES[m] + "/")]
            dense0 = self.segs_off(segs, f"{BRANCH_NAMES[m]}/{'ofCode' if cfg.is3d(m) else 'dense'}/w")
            self.buckets[m] = (mine[0].off, dense0)
            self.buckets[(m, "fc")] = (dense0, round_up(mine[-1].off + mine[-1].n, 64))
        heads = [s for s in segs if "/" in s.name and s.name.split("/")[0] in ("code", "classprob")]
        self.buckets["heads"] = (heads[0].off, off) if heads else None
        d = self.dev
        self.w = self._new_arena(off, exchanged=True)
        self.g = self._new_arena(off, exchanged=True)
        self.m = torch.zeros(off, device=d)
        self.v = torch.zeros(off, device=d)
        self.seg_off = torch.tensor([s.off for s in segs] + [off], dtype=torch.int64, device=d)
        self.seg_l2 = torch.tensor([s.l2 for s in segs], dtype=torch.float32, device=d)
        self.reg_out = torch.zeros(1, device=d)
        self.gstage = None
        if self._symm:                 # fused data-parallel exchange: the scalar is summed over the ranks by peer atomics
            r = self._new_arena(64, exchanged=True)
            if len(self._symm) > 2:
                self.reg_out = r[:1]
            # staging buffer of PUSHED dense-layer gradients: slot p = rank p's part of this rank's arena slice
            self.slice_len = ((off // 4 + self.world - 1) // self.world) * 4
            if os.environ.get("UGN_DP_PUSH", "1") == "1" and len(self._symm) > 2:
                gs = self._new_arena(self.world * self.slice_len, exchanged=True)
                if len(self._symm) > 3:
                    self.gstage = gs
        self.lr_dev = torch.zeros(1, device=d)
        self._lr_host = torch.zeros(1).pin_memory()
        self.R = {k: TRef(t) for k, t in dict(w=self.w, g=self.g, m=self.m, v=self.v, seg_off=self.seg_off,
                                               seg_l2=self.seg_l2, reg_out=self.reg_out, lr_dev=self.lr_dev).items()}
        self.pw: Dict[str, torch.Tensor] = {}   # master views
        self.pg_: Dict[str, torch.Tensor] = {}  # gradient views
        self.Rw: Dict[str, TRef] = {}
        self.Rg: Dict[str, TRef] = {}
        for s in segs:
            self.pw[s.name] = self.w[s.off:s.off + s.n].view(s.shape)
            self.pg_[s.name] = self.g[s.off:s.off + s.n].view(s.shape)
            self.Rw[s.name] = TRef(self.pw[s.name])
            self.Rg[s.name] = TRef(self.pg_[s.name])
        # compute copies: conv / big dense weights, padded (+ bf16 planes in tensor-core mode)
        self.cw: Dict[str, torch.Tensor] = {}
        self.Rcw: Dict[str, TRef] = {}
        for m in range(cfg.nmods):
            bn = BRANCH_NAMES[m]
            if cfg.is3d(m):      # fp32 engine: the masters are the operands
                for s3 in segs:
                    if s3.name.startswith(bn + "/") and s3.name.endswith("/w"):
                        self.cw[s3.name] = self.pw[s3.name]
                continue
            for li, L in enumerate(cfg.layers(m, self.pad)):
                name = f"{bn}/conv{li}/w"
                shape = (L["co"], L["k"], L["k"], L["cp"])
                if self.P:
                    self.cw[name] = shape         # allocated below (one arena for every 16-bit compute copy)
                elif L["cp"] != L["cin"]:
                    self.cw[name] = torch.zeros(shape, device=d)
                else:
                    self.cw[name] = self.pw[name]
            for nm in ("dense", "ofCode"):
                name = f"{bn}/{nm}/w"
                if self.P:
                    self.cw[name] = self.segs[name].shape
                else:
                    self.cw[name] = self.pw[name]
        self.cw_arena = None
        if self.P:
            numel = lambda shp: int(torch.tensor(shp).prod())
            dense_names = [k for k, t in self.cw.items() if isinstance(t, tuple)]
            sizes = {k: round_up(self.P * numel(self.cw[k]), 64) for k in dense_names}
            total = sum(sizes.values())
            esz = 2
            arena = None
            if self._symm and os.environ.get("UGN_DP_CW16", "1") == "1":
                # fused data-parallel exchange: the owner of an arena slice writes the 16-bit planes of its updated dense
                # weights into every rank's compute copies (ugn_dp_optim_step: cw_peers) -- they must be peer-mapped
                nsym = len(self._symm)
                a32 = self._new_arena((total * esz + 3) // 4, exchanged=True)
                if len(self._symm) > nsym:
                    arena = a32.view(self.dt16)
                    self.cw_arena = arena
                    self._cw_symm = self._symm[-1][1]
            if arena is None:
                arena = torch.zeros(total, dtype=self.dt16, device=d)
            o = 0
            for k in dense_names:
                shp = self.cw[k]
                self.cw[k] = arena[o:o + self.P * numel(shp)].view((self.P,) + tuple(shp))
                o += sizes[k]
        for k, t in self.cw.items():
            self.Rcw[k] = TRef(t)
        # optimiser-fused refresh of the unpadded (dense) compute copies: {address, numel} per segment
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
        # deferred weight all-gather: the exchange kernel refreshes this rank's slice of the 16-bit copies only; the copy
        # engines send it to the peers on a side stream while the next step's convolution forward runs, and the first
        # dense GEMM of that step waits on this (externally recorded, graph-capturable) event
        self._cw_event = None
        if (self.cw_arena is not None and self.pack_table is not None and not self.dp_onegraph
                and os.environ.get("UGN_DP_DEFER", "1") == "1"):
            self._cw_event = torch.cuda.Event(external=True)
            self._cw_stream = torch.cuda.Stream(device=d)

    @staticmethod
    def segs_off(segs, name):
        return next(s.off for s in segs if s.name == name)

    def _cw_wait(self):
        """Order the current stream behind the deferred peer copies of the 16-bit dense weights (no-op otherwise)."""
        if getattr(self, "_cw_event", None) is not None:
            torch.cuda.current_stream().wait_event(self._cw_event)

    def sync_master_weights(self):
        """Fused data-parallel exchange with exchanged 16-bit copies: only the OWNER of an arena slice holds the current
        f32 master of the dense weights in it.  Before anything reads the f32 arena as a whole (export / save /
        re-pack / checks) every slice is broadcast from its owner."""
        if not getattr(self, "_w_stale", False):
            return
        sl = self.slice_len
        for r in range(self.world):
            a, b = r * sl, min((r + 1) * sl, self.n_arena)
            if a < b:
                torch.distributed.broadcast(self.w[a:b], src=torch.distributed.get_global_rank(self.pg, r), group=self.pg)
        self._w_stale = False

    def repack_weights(self, after_optim: bool = False):
        """master f32 -> padded / 16-bit compute copies.  After an optimiser step only the padded (conv)
        copies are left to do: the dense ones were re-split inside the optimiser kernel."""
        if not after_optim:
            self._cw_wait()
            self.sync_master_weights()
        for name, t in self.cw.items():
            if after_optim and name in self._fused_pack:
                continue
            if t.data_ptr() != self.pw[name].data_ptr():
                check(lib.ugn_pack_weight(self.ctx.h, self.Rw[name].ptr, self.Rcw[name].ptr, stream_ptr()))

    def init_weights(self, seed: int):
        """Keras defaults: glorot_uniform kernels, zero biases, he_uniform for ofCode
        (nets/mj_uwyhNets_ba.py:82-105)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for s in self.seg_list:
            if s.name.endswith("/b"):
                self.pw[s.name].zero_()
                continue
            if len(s.shape) == 4:
                co, kh, kw, cin = s.shape
                fan_in, fan_out = cin * kh * kw, co * kh * kw
            elif len(s.shape) == 5:
                co, kt, kh, kw, cin = s.shape
                fan_in, fan_out = cin * kt * kh * kw, co * kt * kh * kw
            else:
                fan_out, fan_in = s.shape
            limit = math.sqrt(6.0 / fan_in) if s.name.endswith("ofCode/w") else math.sqrt(6.0 / (fan_in + fan_out))
            vals = (torch.rand(s.shape, generator=g, dtype=torch.float32) * 2 - 1) * limit
            self.pw[s.name].copy_(vals)
        self.repack_weights()

    def load_params(self, params: Dict[str, torch.Tensor]):
        """params in the oracle / PyTorch layout: conv [Cout,Cin,kh,kw], dense [out,in]."""
        for name, val in params.items():
            s = self.segs[name]
            v = val.detach().to(torch.float32)
            if len(s.shape) == 4:
                v = v.permute(0, 2, 3, 1)
            elif len(s.shape) == 5:          # Conv3D: oracle [Cout,Cin,kt,kh,kw] -> [Cout,kt,kh,kw,Cin]
                v = v.permute(0, 2, 3, 4, 1)
            self.pw[name].copy_(v.contiguous().to(self.dev))
        self.repack_weights()

    def oracle_shape(self, name):
        """Shape of a parameter in the oracle / PyTorch layout (conv [Cout,Cin,kh,kw], dense [out,in])."""
        s = self.segs[name].shape
        if len(s) == 5:
            return (s[0], s[4], s[1], s[2], s[3])
        return (s[0], s[3], s[1], s[2]) if len(s) == 4 else tuple(s)

    def _export(self, views) -> Dict[str, torch.Tensor]:
        out = {}
        for s in self.seg_list:
            v = views[s.name].detach().clone()
            if len(s.shape) == 4:
                v = v.permute(0, 3, 1, 2).contiguous()
            elif len(s.shape) == 5:
                v = v.permute(0, 4, 1, 2, 3).contiguous()
            out[s.name] = v
        return out

    def export_params(self):
        self.sync_master_weights()
        return self._export(self.pw)

    def export_grads(self):
        return self._export(self.pg_)

    def set_trainable(self, prefix: str, trainable: bool):
        """layer.trainable of the Keras protocol (freeze_convs / freeze_all / freeze_branches, nets/mj_uwyhNets_ba.py:193,
        :1366-1391): every parameter tensor whose name starts with `prefix` is (un)frozen.  Frozen tensors keep their
        regulariser term in the loss value but receive no update (the optimiser kernel skips them)."""
        hit = 0
        l2 = self.seg_l2.cpu()
        for i, s in enumerate(self.seg_list):
            if s.name == prefix or s.name.startswith(prefix.rstrip("/") + "/"):
                hit += 1
                frozen = float(l2[i]) < 0
                if trainable and frozen:
                    l2[i] = -float(l2[i]) - 1.0
                elif not trainable and not frozen:
                    l2[i] = -(float(l2[i]) + 1.0)
        if not hit:
            raise KeyError(f"no parameter tensor under {prefix!r}")
        self.seg_l2.copy_(l2)
        self._opt_ranges = None          # per-range coefficient tables are rebuilt on demand
        return hit

    def frozen(self) -> List[str]:
        return [s.name for s, c in zip(self.seg_list, self.seg_l2.cpu().tolist()) if c < 0]

    def export_decisions(self, B: int, train: bool = True):
        """The discrete decisions the last forward pass of batch size B took, in the oracle's layout: per modality m
        ``{pool{li}: [B,C,Hp,Wp] arg-max position dy*2+dx, act{li}: [B,C,Hp,Wp] bool (selected pre-activation > 0)}``
        and ``winner`` [B,nd] (max / sign_max fusion).  Read back from the buffers the backward pass itself routes
        gradients with (pool arg-max bytes, stored activations, fusion winner bytes): parity tests inject them into the
        CPU oracle so that gradients are compared on identical routing (tests/test_decisions_gpu.py)."""
        p = self._plans[(B, train)]
        dec = {}
        for m, b in enumerate(p.br):
            d = {}
            for li, L in enumerate(b.layers):
                a = b.T[f"a{li + 1}"]
                a = a.float().sum(0) if self.P else a                       # 16-bit planes: value = hi + lo
                d[f"act{li}"] = (a > 0).permute(0, 3, 1, 2).contiguous().cpu()
                if L["pool"]:
                    d[f"pool{li}"] = b.T[f"idx{li}"].permute(0, 3, 1, 2).contiguous().cpu()
            dec[m] = d
        if not self.cfg.single:
            dec["winner"] = p.T["winner"].cpu()
        return dec

    # ------------------------------------------------------------------ plans
    @staticmethod
    def _check_batch(B: int):
        if B <= 0:      # Keras: model.predict / fit on zero samples
            raise ValueError("Expect x to be a non-empty array or dataset.")

    def _check_shapes(self, p, inputs, flags):
        """Keras' complaint for a mis-shaped input, before any copy is enqueued."""
        for m in range(self.cfg.nmods):
            want, got = tuple(p.br[m].x_in.shape), tuple(inputs[m].shape)
            if got != want and not (getattr(self.cfg, "is3d", None) and self.cfg.is3d(m) and got == want + (1,)):
                raise ValueError(f"Input {m} is incompatible with the model: expected shape=(None, "
                                 f"{', '.join(str(v) for v in want[1:])}), found shape={got}")
            if flags is not None and math.prod(tuple(flags[m].shape)) != want[0]:
                raise ValueError(f"use-flag {m}: expected shape=(None, 1) with {want[0]} rows, found shape={tuple(flags[m].shape)}")

    def plan(self, B: int, train: bool) -> "_Plan":
        self._check_batch(B)
        key = (B, train)
        p = self._plans.get(key)
        if p is None:
            p = self._plans[key] = _Plan(self, B, train)
        return p

    # ------------------------------------------------------------------ forward
    # ---- concurrent modality branches: the branches are independent between the input pack and the fusion
    # (forward) and between the fusion backward and the optimiser (backward), so each runs on its own stream;
    # the HBM-bound kernels of one branch (packs, pool/act backward, weight-streaming dense GEMMs) then
    # overlap the tensor-bound conv kernels of another, also inside a captured CUDA graph (fork/join events)
    def _fork(self):
        if not self.multistream or self.cfg.nmods == 1:
            return None
        if self._bstreams is None:
            self._bstreams = [torch.cuda.Stream(device=self.dev) for _ in range(self.cfg.nmods - 1)]
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        for s in self._bstreams:
            s.wait_event(ev)
        return [main] + self._bstreams

    def _join(self, streams):
        if streams is None:
            return
        main = streams[0]
        for s in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)

    def _forward(self, p: "_Plan", train: bool, expanded: bool = False):
        cfg, h = self.cfg, self.ctx.h
        check(lib.ugn_set_fwd_passes(h, *self.fwd_passes))      # ctx state (host side): engines share the ctx
        streams = self._fork()
        for m in range(cfg.nmods):
            with torch.cuda.stream(streams[m] if streams else torch.cuda.current_stream()):
                self._forward_branch(p, m, train, expanded)
        self._join(streams)
        st = stream_ptr()
        if cfg.single:
            sig = p.br[0].R["out"]
        else:
            if cfg.normbfmerge:
                # "nrmbfl2*" Lambdas (:1167-1168): l2_normalize of every branch output before its gate = the
                # fusion kernel on ONE modality with a unit flag (gate x 1 -> max of one -> l2_normalize)
                for m in range(cfg.nmods):
                    b = p.br[m]
                    check(lib.ugn_fuse_fwd(h, 1, b.nrm_in, p.one_ptrs, b.R["outn"].ptr, None, b.R["nwin"].ptr,
                                           b.R["ninv"].ptr, 0, 1, st))
            # north_star: gate + fusion + l2_normalize + FC1 ("code") [+ its dropout] as ONE kernel when the graph has an FC1
            fused_fc1 = cfg.nc > 0 and cfg.nd % 4 == 0 and cfg.nd <= 12288 and os.environ.get("UGN_FUSE_FC1", "1") == "1"
            drop_code = train and cfg.dropout > 0.001 and cfg.nc > 0
            if fused_fc1:
                fuse_drop = drop_code and not self.post2         # postriplet 2 drops the NORMALISED code further down
                check(lib.ugn_fuse_fc1_fwd(h, cfg.nmods, p.brn_ptrs if cfg.normbfmerge else p.br_ptrs, p.flag_ptrs,
                                           p.R["sig"].ptr, p.R["sig16"].ptr if p.tc_gram else None, p.R["winner"].ptr,
                                           p.R["inv_norm"].ptr, cfg.merge, 0 if self.post2 else 1, self.Rw["code/w"].ptr,
                                           self.Rw["code/b"].ptr, p.R["code"].ptr,
                                           p.R["cmask"].ptr if fuse_drop else None,
                                           p.R["dropcode"].ptr if fuse_drop else None, cfg.act, cfg.alpha, st))
            else:
                check(lib.ugn_fuse_fwd(h, cfg.nmods, p.brn_ptrs if cfg.normbfmerge else p.br_ptrs, p.flag_ptrs,
                                       p.R["sig"].ptr, p.R["sig16"].ptr if p.tc_gram else None, p.R["winner"].ptr,
                                       p.R["inv_norm"].ptr, cfg.merge, 0 if self.post2 else 1, st))
            sig = p.R["sig"]
            if self.aux:
                # auxiliary classifiers on the GATED branch outputs: gate = the fusion kernel on one modality without
                # the normalisation (x * flag), then Dense(nclasses)
                for m in range(cfg.nmods):
                    b = p.br[m]
                    check(lib.ugn_fuse_fwd(h, 1, b.gate_in, b.flag1, b.R["gated"].ptr, None, b.R["gwin"].ptr,
                                           b.R["ginv"].ptr, 0, 0, st))
                    check(lib.ugn_linear_fwd(h, b.R["gated"].ptr, self.Rw[f"{AUX_NAMES[m]}/w"].ptr,
                                             self.Rw[f"{AUX_NAMES[m]}/b"].ptr, None, b.R["aux_logits"].ptr, None,
                                             ACT_LINEAR, 0.0, st))
        feat = sig
        if cfg.nc > 0:
            cmask = p.R["cmask"].ptr if (train and cfg.dropout > 0.001) else None
            fused = (not cfg.single) and fused_fc1
            if not fused:
                check(lib.ugn_linear_fwd(h, sig.ptr, self.Rw["code/w"].ptr, self.Rw["code/b"].ptr, None, p.R["code"].ptr,
                                         None, cfg.act, cfg.alpha, st))
            emb = p.code
            if self.post2:
                # postriplet == 2 (:819-832): the Dense above is the layer "signature"; its l2_normalize ("code") is the
                # embedding of the triplet loss and the classifier input = the fusion kernel on ONE modality, unit flag
                check(lib.ugn_fuse_fwd(h, 1, p.code_ptrs, p.one_ptrs, p.R["codeN"].ptr, None, p.R["cwin"].ptr,
                                       p.R["cinv"].ptr, 0, 1, st))
                emb, sig = p.codeN, p.R["codeN"]
            if cmask is not None:
                if not (fused and not self.post2):           # (the fused kernel wrote dropcode = code * mask already)
                    torch.mul(emb, p.cmask, out=p.dropcode)
                feat = p.R["dropcode"]
            else:
                feat = p.R["codeN"] if self.post2 else p.R["code"]
        if cfg.nclasses > 0:
            check(lib.ugn_linear_fwd(h, feat.ptr, self.Rw["classprob/w"].ptr, self.Rw["classprob/b"].ptr, None,
                                     p.R["logits"].ptr, None, ACT_LINEAR, 0.0, st))
        return sig, feat

    def _forward_branch(self, p: "_Plan", m: int, train: bool, expanded: bool):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        bn = BRANCH_NAMES[m]
        b = p.br[m]
        if cfg.is3d(m):
            return self._forward_branch3d(p, m, expanded)
        if expanded and getattr(p, "use_augment", False):
            # + integer shifts of the random transform on every modality, magnitude clip on the optical flow (modality 0)
            check(lib.ugn_pack_input_augment(h, b.R["x_base"].ptr, p.R["src_row"].ptr,
                                             None if cfg.single else p.R_flags[m].ptr,
                                             p.R["mirror"].ptr if p.use_mirror else None, p.R["shift"].ptr,
                                             p.R["clip"].ptr if (m == 0 and cfg.in_channels[0] == 50) else None,
                                             OF_CLIP_LO, OF_CLIP_HI, OF_CLIP_VAL, NOISE, b.R["a0"].ptr, st))
        elif expanded:
            # device-side missing-modality expansion: row i reads base row src_row[i]; a cleared
            # use-flag turns the row's volume into the reference's 1e-9 constant
            check(lib.ugn_pack_input_expand(h, b.R["x_base"].ptr, p.R["src_row"].ptr,
                                            None if cfg.single else p.R_flags[m].ptr,
                                            p.R["mirror"].ptr if p.use_mirror else None, NOISE,
                                            b.R["a0"].ptr, st))
        else:
            check(lib.ugn_pack_input(h, b.R["x_in"].ptr, b.R["a0"].ptr, st))
        for li, L in enumerate(b.layers):
            check(lib.ugn_conv2d_fwd(h, b.R[f"a{li}"].ptr, self.Rcw[f"{bn}/conv{li}/w"].ptr,
                                     self.Rw[f"{bn}/conv{li}/b"].ptr, b.R[f"a{li + 1}"].ptr,
                                     b.R[f"idx{li}"].ptr if L["pool"] else None, cfg.act, cfg.alpha,
                                     int(L["pool"]), st))
        nl = len(b.layers)
        check(lib.ugn_flatten_chw(h, b.R[f"a{nl}"].ptr, b.R["flat"].ptr, st))
        drop = train and cfg.dropout > 0.001
        self._cw_wait()        # data parallel: the peers' slices of the dense weights land underneath the convolutions
        if drop and p.use_philox:
            check(lib.ugn_linear_fwd_philox(h, b.R["flat"].ptr, self.Rcw[f"{bn}/dense/w"].ptr, self.Rw[f"{bn}/dense/b"].ptr,
                                            self._R_rng.ptr, m, 1.0 - cfg.dropout, b.R["h1"].ptr, b.R["h1_16"].ptr,
                                            ACT_LINEAR, 0.0, st))
        else:
            mask = b.R["mask"].ptr if drop else None
            check(lib.ugn_linear_fwd(h, b.R["flat"].ptr, self.Rcw[f"{bn}/dense/w"].ptr, self.Rw[f"{bn}/dense/b"].ptr,
                                     mask, b.R["h1"].ptr, b.R["h1_16"].ptr if self.P else None, ACT_LINEAR, 0.0, st))
        check(lib.ugn_linear_fwd(h, (b.R["h1_16"] if self.P else b.R["h1"]).ptr, self.Rcw[f"{bn}/ofCode/w"].ptr,
                                 self.Rw[f"{bn}/ofCode/b"].ptr, None, b.R["out"].ptr, None, ACT_LINEAR, 0.0, st))

    def _set_inputs(self, p, inputs, flags, labels=None, drop_masks=None, code_drop_mask=None):
        cfg = self.cfg
        self._check_shapes(p, inputs, None if cfg.single else flags)
        for m in range(cfg.nmods):
            p.br[m].x_in.copy_(inputs[m].reshape(p.br[m].x_in.shape) if cfg.is3d(m) else inputs[m], non_blocking=True)
            if not cfg.single:
                p.flags[m].copy_(flags[m].reshape(-1, 1), non_blocking=True)
            if p.train and cfg.dropout > 0.001 and hasattr(p.br[m], "mask"):
                if drop_masks is not None:
                    p.br[m].mask.copy_(drop_masks[m])
                elif not self.philox:
                    keep = 1.0 - cfg.dropout
                    p.br[m].mask.bernoulli_(keep).div_(keep)
        p.use_philox = self.philox and drop_masks is None and p.train and cfg.dropout > 0.001
        if p.train and cfg.dropout > 0.001 and cfg.nc > 0:
            if code_drop_mask is not None:
                p.cmask.copy_(code_drop_mask)
            elif not self.philox:
                keep = 1.0 - cfg.dropout
                p.cmask.bernoulli_(keep).div_(keep)
        p.philox_code = self.philox and code_drop_mask is None and p.train and cfg.dropout > 0.001 and cfg.nc > 0
        if labels is not None:
            lab = labels.reshape(-1).to(torch.int32)
            if getattr(cfg, "pair_loss", False):      # B pair labels for 2B rows
                if lab.numel() * 2 != p.B:
                    raise ValueError(f"pair labels: expected {p.B // 2} entries for {p.B} rows, found {lab.numel()}")
                p.labels[:lab.numel()].copy_(lab, non_blocking=True)
            else:
                p.labels.copy_(lab, non_blocking=True)

    @torch.no_grad()
    def predict(self, inputs: Sequence[torch.Tensor], flags: Optional[Sequence[torch.Tensor]] = None,
                layer: str = "signature") -> torch.Tensor:
        """model_code.predict of the reference test scripts
        (mains/mj_testUWYHGaitNet_open_tum.py:139-148,192-198).  layer in
        {signature, code, classprob(logits), flatten}."""
        B = int(inputs[0].shape[0])
        p = self.plan(B, False)
        self._set_inputs(p, inputs, flags)
        self._forward(p, False)
        return self._layer_output(p, layer)

    def _layer_output(self, p, layer: str) -> torch.Tensor:
        """A named layer of the graph after a forward pass; "embedding" = the model's output 0 (what the triplet loss sees)."""
        if self.post2:      # postriplet == 2: "signature" is the Dense layer (:821-826), "code" its l2_normalize (:828)
            if layer == "signature":
                y = p.code
                return (y.clone() if self.cfg.act != ACT_LEAKY else torch.where(y > 0, y, y / self.cfg.alpha))
            if layer in ("code", "embedding"):
                return p.codeN.clone()
        if layer in ("signature", "embedding"):
            return (p.br[0].out if self.cfg.single else p.sig).clone()
        if layer == "code":
            return p.code.clone()
        if layer in ("classprob", "logits"):
            return p.logits.clone()
        raise KeyError(layer)

    # ------------------------------------------------------------------ backward
    def _reduce_bucket(self, key):
        """Data-parallel gradient exchange, bucketed so that the all-reduce of a finished branch overlaps
        the backward pass of the next one (NCCL runs on its own stream)."""
        if self.dp_reduce in ("single", "fused"):
            return
        if (self.world > 1 or self._cap is not None) and self._dp_async and self.buckets.get(key) is not None:
            if self._cap is not None:          # capturing: close this graph segment, the all-reduce runs between
                self._cut(key)                 # the replays of two segments
                return
            lo, hi = self.buckets[key]
            self._works.append(torch.distributed.all_reduce(self.g[lo:hi], group=self.pg, async_op=True))

    # ---- data-parallel CUDA graphs: the step is captured as SEGMENTS cut at the all-reduce points, so the
    # NCCL calls stay ordinary (un-captured) stream operations between two graph replays
    def _cut(self, key):
        cap = self._cap
        if self.ctx.launches == cap["mark"] and cap["segs"]:
            cap["segs"][-1][1].append(key)      # nothing launched since the last cut: same boundary
            return
        cap["cur"].capture_end()
        cap["segs"].append((cap["cur"], [key]))
        g = torch.cuda.CUDAGraph()
        g.capture_begin(pool=cap["pool"])
        cap["cur"] = g
        cap["mark"] = self.ctx.launches

    def _capture_segments(self, p, expanded):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        torch.cuda.synchronize()
        l0 = self.ctx.launches
        # no garbage collection while a capture is open: the finaliser of an unrelated, dead CUDAGraph (a previous
        # engine) calls cudaGraph reset, which is illegal during capture and invalidates it
        import gc
        gc.collect()
        gc_on = gc.isenabled()
        gc.disable()
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            self._cap = {"cur": g, "segs": [], "pool": torch.cuda.graph_pool_handle(), "mark": self.ctx.launches}
            g.capture_begin(pool=self._cap["pool"])
            try:
                self._step_body(p, True, expanded)
                self._cap["cur"].capture_end()
                self._cap["segs"].append((self._cap["cur"], []))
            finally:
                segs, self._cap = self._cap["segs"], None
                if gc_on:
                    gc.enable()
        cur.wait_stream(side)
        self.graph_launches = self.ctx.launches - l0
        return segs

    def _replay_segments(self, segs):
        self._works = []
        timing = getattr(self, "_dp_timing", None)
        if timing is None and os.environ.get("UGN_DP_TIMING"):
            timing = self._dp_timing = []
        if timing is not None:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * len(segs) + 2)]
            evs[0].record()
        for si, (g, keys) in enumerate(segs):
            g.replay()
            if timing is not None:
                evs[2 * si + 1].record()
            for key in keys:
                if key == "fused":
                    self._dp_fused_exchange()
                    if timing is not None:
                        evs[2 * si + 2].record()
                    continue
                if key == "wait":
                    if self.dp_reduce == "single":
                        if self.world > 1:
                            torch.distributed.all_reduce(self.g, group=self.pg)
                        continue
                    for w in self._works:
                        w.wait()
                elif self.world > 1:
                    lo, hi = self.buckets[key]
                    self._works.append(torch.distributed.all_reduce(self.g[lo:hi], group=self.pg, async_op=True))
        if timing is not None:
            evs[-1].record()
            timing.append(evs)

    def dp_timing_summary(self):
        """Development aid (UGN_DP_TIMING=1): mean ms of [segment 1 | exchange | segment 2] over the recorded steps."""
        t = getattr(self, "_dp_timing", None)
        if not t:
            return None
        torch.cuda.synchronize()
        rows = []
        for evs in t[5:]:
            try:
                rows.append([evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])])
            except Exception:
                pass
        if not rows:
            return None
        m = torch.tensor(rows).mean(0).tolist()
        return {"fwd_bwd_ms": m[0], "exchange_ms": m[1], "repack_ms": m[2], "steps": len(rows)}

    def _losses_and_backward(self, p: "_Plan", sig: TRef, feat: TRef):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        B = p.B
        self._works = []
        # triplet: demb = wver * dL/dsig
        if getattr(cfg, "pair_loss", False):         # UWYHNet.build: VerifLossLayer on the two halves of the batch
            check(lib.ugn_pair_verif_loss(h, sig.ptr, p.R["labels"].ptr, cfg.margin, cfg.wver, p.R["trip_out"].ptr,
                                          p.R["dsig"].ptr, p.R["trip_ws"].ptr, st))
        elif getattr(cfg, "triplet_hard", False):      # compile_hard: tfa TripletHardLoss (nets/mj_uwyhNets_ba.py:1302-1306)
            check(lib.ugn_triplet_hard(h, sig.ptr, p.R["sig16"].ptr if p.tc_gram else None, p.R["labels"].ptr,
                                       cfg.margin, cfg.wver, p.R["trip_out"].ptr,
                                       (p.R["dcodeN"] if self.post2 else p.R["dsig"]).ptr, p.R["trip_ws"].ptr, st))
        elif p.tc_gram:      # B x B Gram matrix on the tensor cores from the hi/lo planes the fusion kernel wrote (north_star)
            check(lib.ugn_triplet_all_tc(h, sig.ptr, p.R["sig16"].ptr, p.R["labels"].ptr, cfg.margin, cfg.wver,
                                         p.R["trip_out"].ptr, p.R["dsig"].ptr, p.R["trip_ws"].ptr, st))
        else:
            check(lib.ugn_triplet_all(h, sig.ptr, p.R["labels"].ptr, cfg.margin, cfg.wver, p.R["trip_out"].ptr,
                                      (p.R["dcodeN"] if self.post2 else p.R["dsig"]).ptr, p.R["trip_ws"].ptr, st))
        if self.post2 and cfg.nclasses == 0:
            self._post2_backward(p, None)
        if cfg.nclasses > 0:
            check(lib.ugn_softmax_ce_ls(h, p.R["logits"].ptr, p.R["labels"].ptr, p.R["ce_out"].ptr, p.R["dlogits"].ptr,
                                        cfg.wid, cfg.label_smoothing, st))
            check(lib.ugn_linear_bwd(h, feat.ptr, self.Rw["classprob/w"].ptr, p.R["dlogits"].ptr, p.R["dfeat"].ptr,
                                     self.Rg["classprob/w"].ptr, self.Rg["classprob/b"].ptr, st))
            dfeat = p.R["dfeat"]
            if self.post2:
                self._post2_backward(p, dfeat)
            elif cfg.nc > 0:
                # activity regulariser l2(1e-3) on "code": + 1e-3*sum(code^2)/B  (:1196)
                use_mask = cfg.dropout > 0.001
                check(lib.ugn_act_mask_bwd(h, dfeat.ptr, None, p.R["cmask"].ptr if use_mask else None,
                                           p.R["dcode"].ptr, None, ACT_LINEAR, 0.0, st))
                relu = cfg.act != ACT_LEAKY
                if relu:        # Dense(activation='relu', activity_regularizer): the activated output is regularised
                    p.dcode.add_(p.code, alpha=2e-3 / B)
                check(lib.ugn_act_mask_bwd(h, p.R["dcode"].ptr, p.R["code"].ptr, None, p.R["dcode_z"].ptr, None,
                                           cfg.act, cfg.alpha, st))
                if not relu:    # Dense(activation=None, activity_regularizer) + LeakyReLU (:1198-1201): the LINEAR output,
                    #             recovered from the activated one (z = y for y > 0, y / alpha otherwise)
                    torch.where(p.code > 0, p.code, p.code / cfg.alpha, out=p.dsig2_code)
                    p.dcode_z.add_(p.dsig2_code, alpha=2e-3 / B)
                self._activity_reg_value(p, relu)
                check(lib.ugn_linear_bwd(h, sig.ptr, self.Rw["code/w"].ptr, p.R["dcode_z"].ptr, p.R["dsig2"].ptr,
                                         self.Rg["code/w"].ptr, self.Rg["code/b"].ptr, st))
                p.dsig.add_(p.dsig2)
            else:
                p.dsig.add_(p.dfeat)
        if self.aux:
            for m in range(cfg.nmods):
                b = p.br[m]
                check(lib.ugn_softmax_ce_ls(h, b.R["aux_logits"].ptr, p.R["labels"].ptr, b.R["aux_ce"].ptr,
                                            b.R["daux_logits"].ptr, cfg.waux, cfg.label_smoothing, st))
                check(lib.ugn_linear_bwd(h, b.R["gated"].ptr, self.Rw[f"{AUX_NAMES[m]}/w"].ptr, b.R["daux_logits"].ptr,
                                         b.R["dgated"].ptr, self.Rg[f"{AUX_NAMES[m]}/w"].ptr,
                                         self.Rg[f"{AUX_NAMES[m]}/b"].ptr, st))
                check(lib.ugn_fuse_bwd(h, 1, b.R["dgated"].ptr, b.R["gated"].ptr, b.R["gwin"].ptr, b.R["ginv"].ptr,
                                       b.flag1, b.aux_dout, 0, 0, st))
        if self.dp_reduce != "split":
            self._reduce_bucket("heads")
        if self.scaled:
            # fp16 gradient operands: pick this step's power-of-two scale from the signature gradient
            check(lib.ugn_grad_scale_update(h, p.R["dsig"].ptr, GRAD_SCALE_TARGET, st))
            self.ctx.grad_scaled = True
        elif getattr(self.ctx, "grad_scaled", False):
            # another engine on this ctx left a scale behind: this mode's gradients are unscaled
            check(lib.ugn_grad_scale_set(h, 1.0, st))
            self.ctx.grad_scaled = False
        if cfg.single:
            p.br[0].dout.copy_(p.dsig)
        else:
            check(lib.ugn_fuse_bwd(h, cfg.nmods, p.R["dsig"].ptr, p.R["sig"].ptr, p.R["winner"].ptr,
                                   p.R["inv_norm"].ptr, p.flag_ptrs, p.dbrn_ptrs if cfg.normbfmerge else p.dbr_ptrs,
                                   cfg.merge, 0 if self.post2 else 1, st))
            if self.aux:       # + the auxiliary heads' gradient wrt the (normalised) branch output
                for m in range(cfg.nmods):
                    b = p.br[m]
                    (b.T["doutn"] if cfg.normbfmerge else b.dout).add_(b.T["daux"])
            if cfg.normbfmerge:
                for m in range(cfg.nmods):
                    b = p.br[m]
                    check(lib.ugn_fuse_bwd(h, 1, b.R["doutn"].ptr, b.R["outn"].ptr, b.R["nwin"].ptr, b.R["ninv"].ptr,
                                           p.one_ptrs, b.nrm_dout, 0, 1, st))
        # per-branch backward on concurrent streams (single GPU; with data parallelism the branches stay in
        # sequence so that every all-reduce bucket is issued from the one stream NCCL orders against)
        self._run_branch_backward(p)

    def _activity_reg_value(self, p, relu: bool):
        """Value of Dense(..., activity_regularizer=l2(1e-3)) on "code": 1e-3 * sum(out^2) / batch -- part of the total
        loss Keras reports (its gradient is folded into the backward pass above).  loss_pack[5]."""
        src = p.code if relu else p.dsig2_code            # activated output (ReLU) | linear output (LeakyReLU path)
        torch.mul(src, src, out=p.act_sq)
        torch.sum(p.act_sq, dim=(0, 1), keepdim=False, out=p.loss_pack[5])
        p.loss_pack[5:6].mul_(1e-3 / p.B)

    def _post2_backward(self, p, dfeat):
        """postriplet == 2: gradient of [triplet(codeN) + CE(classprob(dropout(codeN)))] back to the un-normalised fusion:
        dropout mask -> l2_normalize backward -> activity regulariser + activation -> Dense "signature" -> dsig."""
        cfg, h, st, B = self.cfg, self.ctx.h, stream_ptr(), p.B
        if dfeat is not None:
            use_mask = cfg.dropout > 0.001
            check(lib.ugn_act_mask_bwd(h, dfeat.ptr, None, p.R["cmask"].ptr if use_mask else None, p.R["dcode"].ptr,
                                       None, ACT_LINEAR, 0.0, st))
            p.dcodeN.add_(p.dcode)
        check(lib.ugn_fuse_bwd(h, 1, p.R["dcodeN"].ptr, p.R["codeN"].ptr, p.R["cwin"].ptr, p.R["cinv"].ptr, p.one_ptrs,
                               p.dcode_ptrs, 0, 1, st))                      # -> p.dcode = dL/d(activated Dense output)
        relu = cfg.act != ACT_LEAKY
        if relu:
            p.dcode.add_(p.code, alpha=2e-3 / B)
        check(lib.ugn_act_mask_bwd(h, p.R["dcode"].ptr, p.R["code"].ptr, None, p.R["dcode_z"].ptr, None, cfg.act,
                                   cfg.alpha, st))
        if not relu:
            torch.where(p.code > 0, p.code, p.code / cfg.alpha, out=p.dsig2_code)
            p.dcode_z.add_(p.dsig2_code, alpha=2e-3 / B)
        self._activity_reg_value(p, relu)
        check(lib.ugn_linear_bwd(h, p.R["sig"].ptr, self.Rw["code/w"].ptr, p.R["dcode_z"].ptr, p.R["dsig"].ptr,
                                 self.Rg["code/w"].ptr, self.Rg["code/b"].ptr, st))

    def _run_branch_backward(self, p):
        """Per-branch backward.  One GPU: every branch on its own stream, start to end (the HBM-bound dense GEMMs
        of one branch overlap the tensor-bound convolutions of another).  Data parallel, dp_reduce == "split": two
        phases -- the dense layers of all branches (92 % of the gradient bytes), whose all-reduce is then issued
        and overlaps phase two, the convolution stacks; the small conv buckets follow."""
        cfg = self.cfg
        dp = self.world > 1 or self._cap is not None
        if dp and self.dp_reduce == "split" and hasattr(self, "_backward_branch_fc"):
            for phase, keys in ((self._backward_branch_fc, ["heads"] + [(m, "fc") for m in range(cfg.nmods)]),
                                (self._backward_branch_conv, list(range(cfg.nmods)))):
                streams = self._fork()
                for m in range(cfg.nmods):
                    with torch.cuda.stream(streams[m] if streams else torch.cuda.current_stream()):
                        phase(p, m)
                self._join(streams)
                for key in keys:
                    self._reduce_bucket(key)
            return
        streams = self._fork() if self._branches_concurrent() else None
        push = dp and self.dp_reduce == "fused" and self.gstage is not None and self.world > 1
        for m in range(cfg.nmods):
            with torch.cuda.stream(streams[m] if streams else torch.cuda.current_stream()):
                self._backward_branch_fc(p, m)
                if self._early_active:
                    self._optim_early(m, torch.cuda.current_stream())
                if push:
                    self._push_dense_grads(m, torch.cuda.current_stream())
                self._backward_branch_conv(p, m)
        self._join(streams)
        if push:                           # the pushes must have left before this rank enters the exchange barrier
            ev = torch.cuda.Event()
            ev.record(self._push_stream)
            torch.cuda.current_stream().wait_event(ev)

    def _push_dense_grads(self, m: int, branch_stream):
        """Copy-engine push of branch m's dense-layer gradients (complete right after its two dense backward GEMMs) to
        the ranks that own those arena slices, on a side stream underneath the convolution backward: 92 % of the
        reduce-scatter bytes leave the critical path (the exchange kernel then sums them from its own HBM)."""
        if getattr(self, "_push_stream", None) is None:
            self._push_stream = torch.cuda.Stream(device=self.dev)
            hs = self._symm[3][1]
            n = self.world * self.slice_len
            self._stage_peers = [hs.get_buffer(r, (n,), torch.float32) for r in range(self.world)]
        rank, sl = torch.distributed.get_rank(self.pg), self.slice_len
        lo, hi = self.buckets[(m, "fc")]
        ev = torch.cuda.Event()
        ev.record(branch_stream)
        self._push_stream.wait_event(ev)
        with torch.cuda.stream(self._push_stream):
            for r in range(self.world):
                if r == rank:
                    continue
                a, b = max(lo, r * sl), min(hi, (r + 1) * sl, self.n_arena)
                if a < b:
                    dst = self._stage_peers[r][rank * sl + (a - r * sl): rank * sl + (b - r * sl)]
                    dst.copy_(self.g[a:b], non_blocking=True)

    def _branches_concurrent(self) -> bool:
        """Branches on concurrent streams from start to end: always on one GPU; with data parallelism only when
        the gradients are exchanged in one call after the backward pass (dp_reduce == "single")."""
        return (self.world == 1 and self._cap is None) or self.dp_reduce in ("single", "fused")

    def _backward_branch(self, p: "_Plan", m: int):
        self._backward_branch_fc(p, m)
        self._backward_branch_conv(p, m)

    # ---- use3D: Conv3D branch (build_3Dbranch{,LReLU}, nets/mj_uwyhNets_ba.py:336-417) on the fp32 validation engine
    def _forward_branch3d(self, p: "_Plan", m: int, expanded: bool):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        if expanded:
            raise NotImplementedError("device-side expansion is not wired for use3D branches: feed the expanded batch")
        bn, b = BRANCH_NAMES[m], p.br[m]
        for li, L in enumerate(b.layers3d):
            s3 = L["s"]
            check(lib.ugn_conv3d_fwd(h, b.R[f"a{li}"].ptr, self.Rw[f"{bn}/conv{li}/w"].ptr, self.Rw[f"{bn}/conv{li}/b"].ptr,
                                     b.R[f"a{li + 1}"].ptr, s3[0], s3[1], s3[2], cfg.act, cfg.alpha, st))
        # Conv3D(nd, 1x1x1) "grayCode" on the 1x1x1 volume + Flatten == a Dense layer
        check(lib.ugn_linear_fwd(h, b.R["flat"].ptr, self.Rw[f"{bn}/ofCode/w"].ptr, self.Rw[f"{bn}/ofCode/b"].ptr, None,
                                 b.R["out"].ptr, None, ACT_LINEAR, 0.0, st))

    def _backward_branch3d_fc(self, p: "_Plan", m: int):
        h, st = self.ctx.h, stream_ptr()
        bn, R = BRANCH_NAMES[m], p.br[m].R
        check(lib.ugn_linear_bwd(h, R["flat"].ptr, self.Rw[f"{bn}/ofCode/w"].ptr, R["dout"].ptr, R["dflat"].ptr,
                                 self.Rg[f"{bn}/ofCode/w"].ptr, self.Rg[f"{bn}/ofCode/b"].ptr, st))

    def _backward_branch3d_conv(self, p: "_Plan", m: int):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        bn, b = BRANCH_NAMES[m], p.br[m]
        R = b.R
        for li in range(len(b.layers3d) - 1, -1, -1):
            s3 = b.layers3d[li]["s"]
            # dz = da * act'(a) on the flattened [rows, Cout] views, then kernel / bias / input gradients
            check(lib.ugn_act_mask_bwd(h, R[f"da{li + 1}f"].ptr, R[f"a{li + 1}f"].ptr, None, R[f"dz{li}f"].ptr, None,
                                       cfg.act, cfg.alpha, st))
            check(lib.ugn_conv3d_wgrad(h, R[f"a{li}"].ptr, R[f"dz{li}"].ptr, self.Rg[f"{bn}/conv{li}/w"].ptr,
                                       self.Rg[f"{bn}/conv{li}/b"].ptr, s3[0], s3[1], s3[2], st))
            if li > 0:
                check(lib.ugn_conv3d_dgrad(h, R[f"dz{li}"].ptr, self.Rw[f"{bn}/conv{li}/w"].ptr, R[f"da{li}"].ptr,
                                           s3[0], s3[1], s3[2], st))
        if self.dp_reduce == "bucketed":
            self._reduce_bucket(m)

    def _backward_branch_fc(self, p: "_Plan", m: int):
        """ofCode + dense (+dropout) backward of one branch: 92 % of the branch's gradient bytes."""
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        if cfg.is3d(m):
            return self._backward_branch3d_fc(p, m)
        bn = BRANCH_NAMES[m]
        b = p.br[m]
        R = b.R
        use_mask = cfg.dropout > 0.001
        if self.P:
            check(lib.ugn_act_mask_bwd(h, R["dout"].ptr, None, None, None, R["dout16"].ptr, ACT_LINEAR, 0.0, st))
        # tensor-core mode: the bias gradient comes from the f32 dout, not from its rounded 16-bit copy (the rows of
        # dL/dsignature cancel under the triplet loss: summing the rounded operand leaves mostly rounding noise)
        if self.P and getattr(p, "use_philox", False):
            # the same with the dropout mask REGENERATED inside the post pass (no mask tensor)
            check(lib.ugn_linear_bwd_philox(h, R["h1_16"].ptr, self.Rcw[f"{bn}/ofCode/w"].ptr, R["dout16"].ptr, R["dh1"].ptr,
                                            self._R_rng.ptr, m, 1.0 - cfg.dropout, R["dz1_16"].ptr,
                                            self.Rg[f"{bn}/dense/b"].ptr, self.Rg[f"{bn}/ofCode/w"].ptr, None, st))
            check(lib.ugn_colsum(h, R["dout"].ptr, self.Rg[f"{bn}/ofCode/b"].ptr, st))
            check(lib.ugn_linear_bwd(h, R["flat"].ptr, self.Rcw[f"{bn}/dense/w"].ptr, R["dz1_16"].ptr, R["dflat"].ptr,
                                     self.Rg[f"{bn}/dense/w"].ptr, None, st))
            if self.dp_reduce == "bucketed":
                self._reduce_bucket((m, "fc"))
            return
        if self.P and p.B <= 128:
            # ofCode backward with the fused input-gradient post pass: ONE kernel applies the dropout mask to the split-K
            # sums, writes the dense layer's 16-bit gradient operand and its bias gradient (was: mask/convert + column sums)
            check(lib.ugn_linear_bwd_ex(h, R["h1_16"].ptr, self.Rcw[f"{bn}/ofCode/w"].ptr, R["dout16"].ptr, R["dh1"].ptr,
                                        R["mask"].ptr if use_mask else None, R["dz1_16"].ptr,
                                        self.Rg[f"{bn}/dense/b"].ptr, self.Rg[f"{bn}/ofCode/w"].ptr, None, st))
            check(lib.ugn_colsum(h, R["dout"].ptr, self.Rg[f"{bn}/ofCode/b"].ptr, st))
            check(lib.ugn_linear_bwd(h, R["flat"].ptr, self.Rcw[f"{bn}/dense/w"].ptr, R["dz1_16"].ptr, R["dflat"].ptr,
                                     self.Rg[f"{bn}/dense/w"].ptr, None, st))
            if self.dp_reduce == "bucketed":
                self._reduce_bucket((m, "fc"))
            return
        check(lib.ugn_linear_bwd(h, (R["h1_16"] if self.P else R["h1"]).ptr, self.Rcw[f"{bn}/ofCode/w"].ptr,
                                 (R["dout16"] if self.P else R["dout"]).ptr, R["dh1"].ptr,
                                 self.Rg[f"{bn}/ofCode/w"].ptr, None if self.P else self.Rg[f"{bn}/ofCode/b"].ptr, st))
        if self.P:
            check(lib.ugn_colsum(h, R["dout"].ptr, self.Rg[f"{bn}/ofCode/b"].ptr, st))
        check(lib.ugn_act_mask_bwd(h, R["dh1"].ptr, None, R["mask"].ptr if use_mask else None,
                                   None if self.P else R["dz1"].ptr, R["dz1_16"].ptr if self.P else None,
                                   ACT_LINEAR, 0.0, st))
        check(lib.ugn_linear_bwd(h, R["flat"].ptr, self.Rcw[f"{bn}/dense/w"].ptr,
                                 (R["dz1_16"] if self.P else R["dz1"]).ptr, R["dflat"].ptr,
                                 self.Rg[f"{bn}/dense/w"].ptr, self.Rg[f"{bn}/dense/b"].ptr, st))
        if self.dp_reduce == "bucketed":
            self._reduce_bucket((m, "fc"))

    def _backward_branch_conv(self, p: "_Plan", m: int):
        cfg, h, st = self.cfg, self.ctx.h, stream_ptr()
        if cfg.is3d(m):
            return self._backward_branch3d_conv(p, m)
        bn = BRANCH_NAMES[m]
        b = p.br[m]
        R = b.R
        nl = len(b.layers)
        check(lib.ugn_unflatten_chw(h, R["dflat"].ptr, R[f"da{nl}"].ptr, st))
        for li in range(nl - 1, -1, -1):
            L = b.layers[li]
            check(lib.ugn_conv2d_bwd_act(h, R[f"da{li + 1}"].ptr, R[f"a{li + 1}"].ptr,
                                         R[f"idx{li}"].ptr if L["pool"] else None, R[f"dz{li}c"].ptr,
                                         self.Rg[f"{bn}/conv{li}/b"].ptr, cfg.act, cfg.alpha, int(L["pool"]), st))
            check(lib.ugn_conv2d_wgrad(h, R[f"a{li}"].ptr, R[f"dz{li}c"].ptr, self.Rg[f"{bn}/conv{li}/w"].ptr,
                                       None, st))
            if li > 0:
                check(lib.ugn_conv2d_dgrad(h, R[f"dz{li}c"].ptr, self.Rcw[f"{bn}/conv{li}/w"].ptr, R[f"da{li}"].ptr, st))
        if self.dp_reduce == "bucketed":
            self._reduce_bucket(m)

    # ---- optimiser over arena RANGES: the dense layers hold 92 % of the parameter bytes and their gradients are
    # final right after the two dense backward GEMMs, so their (HBM-bound) Adam update is issued there, on a side
    # stream, and runs underneath the (tensor-bound) convolution backward instead of after it
    def _opt_range(self, key):
        rng = getattr(self, "_opt_ranges", None)
        if rng is None:
            rng = self._opt_ranges = {}
            keys = [("fc", m) for m in range(self.cfg.nmods)] + [("conv", m) for m in range(self.cfg.nmods)] + ["heads"]
            self.reg_parts = torch.zeros(len(keys), device=self.dev)
            for i, k in enumerate(keys):
                b = self.buckets.get("heads" if k == "heads" else ((k[1], "fc") if k[0] == "fc" else k[1]))
                if b is None:
                    continue
                lo, hi = b
                idx = [j for j, sg in enumerate(self.seg_list) if lo <= sg.off < hi]
                t = dict(w=self.w[lo:hi], g=self.g[lo:hi], m=self.m[lo:hi], v=self.v[lo:hi],
                         seg_off=torch.tensor([self.seg_list[j].off - lo for j in idx] + [hi - lo], dtype=torch.int64,
                                              device=self.dev),
                         seg_l2=self.seg_l2[idx[0]:idx[-1] + 1].clone(), reg_out=self.reg_parts[i:i + 1])
                if self.pack_table is not None:
                    t["pack_table"] = self.pack_table[idx[0]:idx[-1] + 1].clone()
                if self.optimizer == "amsgrad":
                    if not hasattr(self, "vhat"):
                        self.vhat = torch.zeros_like(self.v)
                        self.R["vhat"] = TRef(self.vhat)
                    t["vhat"] = self.vhat[lo:hi]
                rng[k] = (t, {n: TRef(x) for n, x in t.items()})
        return rng.get(key)

    def _optim_call(self, R, gscale: float):
        h, st = self.ctx.h, stream_ptr()
        pk = R["pack_table"].ptr if "pack_table" in R else None
        f16 = int(self.dt16 is torch.float16)
        lr_dev = self.R["lr_dev"].ptr
        if self.optimizer in ("adam", "amsgrad", "adamw"):
            if self.optimizer == "amsgrad" and "vhat" not in R:
                self.vhat = torch.zeros_like(self.v)
                R["vhat"] = TRef(self.vhat)
            check(lib.ugn_adam_step_ex(h, R["w"].ptr, R["g"].ptr, R["m"].ptr, R["v"].ptr,
                                       R["vhat"].ptr if self.optimizer == "amsgrad" else None,
                                       self.decoupled_wd if self.optimizer == "adamw" else 0.0, R["seg_off"].ptr,
                                       R["seg_l2"].ptr, 0.0, self.beta1, self.beta2, self.eps, gscale,
                                       R["reg_out"].ptr, lr_dev, pk, max(self.P, 1), f16, st))
        elif self.optimizer == "sgd":
            check(lib.ugn_sgd_step(h, R["w"].ptr, R["g"].ptr, R["v"].ptr, R["seg_off"].ptr, R["seg_l2"].ptr, 0.0,
                                   self.momentum, gscale, R["reg_out"].ptr, lr_dev, pk, max(self.P, 1), f16, st))
        else:
            raise ValueError(f"unknown optimizer {self.optimizer}")

    def _optim_early(self, m: int, branch_stream):
        """Adam / SGD on the dense-layer range of branch m, on its own stream, ordered after the branch's dense
        backward (the caller's current position on branch_stream)."""
        if self._opt_streams is None:
            self._opt_streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.cfg.nmods)]
        ev = torch.cuda.Event()
        ev.record(branch_stream)
        side = self._opt_streams[m]
        side.wait_event(ev)
        with torch.cuda.stream(side):
            self._optim_call(self._opt_range(("fc", m))[1], 1.0)
        self._early_done.append(side)

    def _optim(self, gscale: float):
        if self._early_done:
            # the dense ranges were updated underneath the conv backward: join them, then the small rest
            main = torch.cuda.current_stream()
            for side in self._early_done:
                ev = torch.cuda.Event()
                ev.record(side)
                main.wait_event(ev)
            self._early_done = []
            for k in [("conv", m) for m in range(self.cfg.nmods)] + ["heads"]:
                r = self._opt_range(k)
                if r is not None:
                    self._optim_call(r[1], gscale)
            torch.sum(self.reg_parts, dim=0, keepdim=True, out=self.reg_out)
        else:
            R = dict(self.R)
            if self.pack_table is None:
                R.pop("pack_table", None)
            self._optim_call(R, gscale)
            if self.optimizer == "amsgrad":
                self.R["vhat"] = R["vhat"]
        self.repack_weights(after_optim=True)

    def _step_body(self, p: "_Plan", do_optim: bool, expanded: bool = False):
        if getattr(p, "use_philox", False) or getattr(p, "philox_code", False):
            h, st = self.ctx.h, stream_ptr()
            check(lib.ugn_dropout_advance(h, self._R_rng.ptr, st))
            if getattr(p, "philox_code", False):      # Dropout("dropcode") after FC1: its mask as a tensor (layer id 8)
                check(lib.ugn_dropout_mask(h, self._R_rng.ptr, 8, 1.0 - self.cfg.dropout, p.R["cmask"].ptr, st))
        sig, feat = self._forward(p, True, expanded)
        # early optimiser (one GPU): the dense ranges are updated inside the backward pass
        self._early_active = bool(do_optim and self.early_optim and self.world == 1 and self._cap is None
                                  and not self.cfg.single and type(self) is UGaitEngine)
        self._losses_and_backward(p, sig, feat)
        self._early_active = False
        if do_optim and self.dp_reduce == "fused" and (self.world > 1 or self._cap is not None):
            # gradient exchange fused into the optimiser: [forward + backward] | barrier, ONE kernel, barrier | [repack]
            # the two group barriers and the exchange kernel are plain stream-ordered kernels: with dp_onegraph they are
            # captured too and the whole data-parallel step replays as ONE CUDA graph (no host work between segments)
            if self._cap is not None and not self.dp_onegraph:
                self._cut("fused")
            else:
                self._dp_fused_exchange()
            # exchanged 16-bit copies: only the padded conv copies are left to re-pack locally
            self.repack_weights(after_optim=getattr(self, "_cw_exchange", False) and self.world > 1)
            p.loss_pack[4:5].copy_(self.reg_out, non_blocking=True)
            return
        if do_optim:
            if self._cap is not None:
                self._cut("wait")
            elif self.world > 1:
                if self._dp_async and self.dp_reduce != "single":
                    for w in self._works:
                        w.wait()
                else:
                    torch.distributed.all_reduce(self.g, group=self.pg)
            self._optim(1.0 / self.world)
            p.loss_pack[4:5].copy_(self.reg_out, non_blocking=True)

    def _dp_fused_exchange(self):
        """barrier | ugn_dp_optim_step (reduce-scatter + optimiser on this rank's slice + all-gather of the weights over
        peer memory) | barrier.  On one GPU (tests of the segment machinery) it degenerates to the plain optimiser."""
        if self.world == 1:
            R = dict(self.R)
            R.pop("pack_table", None)
            self._optim_call(R, 1.0)
            return
        import ctypes
        h, st, R = self.ctx.h, stream_ptr(), self.R
        (_, hw), (_, hg) = self._symm[:2]
        if not hasattr(self, "_peer_tabs"):
            n = self.world
            self._peer_tabs = ((ctypes.c_int64 * n)(*[int(p) for p in hg.buffer_ptrs]),
                               (ctypes.c_int64 * n)(*[int(p) for p in hw.buffer_ptrs]))
            # regulariser value: every rank's scalar (symmetric memory) receives the sum of the slices by peer atomics
            self._reg_tab = None
            if len(self._symm) > 2:
                hr = self._symm[2][1]
                self._reg_tab = (ctypes.c_int64 * n)(*[int(p) for p in hr.buffer_ptrs])
            self._staged_ranges, self._n_staged = None, 0
            if self.gstage is not None:
                self.R["gstage"] = TRef(self.gstage)
                rng = [self.buckets[(m, "fc")] for m in range(self.cfg.nmods)]
                flat = [int(x) for ab in rng for x in ab]
                self._staged_ranges, self._n_staged = (ctypes.c_int64 * len(flat))(*flat), len(rng)
            # NVSwitch multicast mappings (multimem.ld_reduce / multimem.st) when the fabric offers them
            # measured: N = 2 unicast 4.02 ms vs multicast 4.32 ms per step, N = 8 unicast 4.53 vs multicast 4.29 -- the
            # unicast path moves 2(N-1)/N arenas per GPU and direction, the multicast path (1 + 1/N): on from N = 4
            mc = (0, 0)
            want = os.environ.get("UGN_DP_MULTIMEM", "auto")
            if want == "1" or (want == "auto" and n >= 4):
                try:
                    mc = (int(hg.multicast_ptr or 0), int(hw.multicast_ptr or 0))
                except Exception:
                    mc = (0, 0)
            self._mc = mc if all(mc) else (0, 0)
            self._cw_tab, self._cw_mc = None, 0
            if getattr(self, "cw_arena", None) is not None:
                self._cw_tab = (ctypes.c_int64 * n)(*[int(p) for p in self._cw_symm.buffer_ptrs])
                if self._cw_event is not None:
                    self._cw_mc = -1                 # UGN_CW_DEFERRED
                    self._cw_deferred_setup()
                elif all(self._mc):
                    try:
                        self._cw_mc = int(self._cw_symm.multicast_ptr or 0)
                    except Exception:
                        self._cw_mc = 0
                    if not self._cw_mc:
                        self._cw_tab = None          # multicast arenas but no multicast copy arena: keep the f32 path
            self._cw_exchange = self._cw_tab is not None and getattr(self, "pack_table", None) is not None
            if not self._cw_exchange:
                self._cw_tab = None
        gp, wp = self._peer_tabs
        adam = self.optimizer in ("adam", "amsgrad", "adamw")
        if self.optimizer == "amsgrad" and "vhat" not in R:
            self.vhat = torch.zeros_like(self.v)
            R["vhat"] = TRef(self.vhat)
        if not adam and self.optimizer != "sgd":
            raise ValueError(f"unknown optimizer {self.optimizer}")
        if self._reg_tab is not None:
            self.reg_out.zero_()                # before the barrier: peers add their slice values into it
        hg.barrier(channel=0)                   # every rank's gradients are complete
        check(lib.ugn_dp_optim_step(h, 0 if adam else 1, self.world, torch.distributed.get_rank(self.pg), gp, wp,
                                    self._mc[0], self._mc[1], R["w"].ptr, R["g"].ptr, R["m"].ptr if adam else None, R["v"].ptr,
                                    R["vhat"].ptr if self.optimizer == "amsgrad" else None,
                                    self.decoupled_wd if self.optimizer == "adamw" else 0.0, R["seg_off"].ptr,
                                    R["seg_l2"].ptr, self.beta1 if adam else self.momentum, self.beta2, self.eps,
                                    R["reg_out"].ptr, self._reg_tab, R["lr_dev"].ptr,
                                    self.R["gstage"].ptr if self.gstage is not None else None,
                                    self._staged_ranges, self._n_staged,
                                    R["pack_table"].ptr if self._cw_tab is not None else None, max(self.P, 1),
                                    int(self.dt16 is torch.float16), self._cw_tab, self._cw_mc, st))
        if self._cw_tab is not None:
            self._w_stale = True
        if self._cw_mc == -1:
            self._cw_send()
        hw.barrier(channel=0)                   # every rank's slice of the new weights (and of the regulariser sum) has landed
        if self._reg_tab is None:
            torch.distributed.all_reduce(self.reg_out, group=self.pg)  # fallback: sum of the slice values (4 bytes)

    def _cw_deferred_setup(self):
        """(destination view in a peer's arena, local source view) for every plane range of the exchanged 16-bit copies
        that lies in this rank's arena slice."""
        from .dist import owned_segment_ranges
        rank = torch.distributed.get_rank(self.pg)
        arena = self.cw_arena
        peers = [self._cw_symm.get_buffer(r, (arena.numel(),), self.dt16) for r in range(self.world)]
        self._cw_pairs = []
        segs = [(sg.name, sg.off, sg.n) for sg in self.seg_list if sg.name in self._fused_pack]
        for name, lo, hi in owned_segment_ranges(segs, rank, self.world, self.n_arena):
            planes = self.cw[name].view(self.P, -1)
            for pl in range(self.P):
                src = planes[pl, lo:hi]
                e0 = (src.data_ptr() - arena.data_ptr()) // 2
                for r in range(self.world):
                    if r != rank:
                        self._cw_pairs.append((peers[r][e0:e0 + (hi - lo)], src))

    def _cw_send(self):
        """After the exchange kernel: this rank's slice of the updated 16-bit planes goes to every peer by the copy
        engines on the side stream; a group barrier there, then the event the next reader of the copies waits on."""
        cur, side = torch.cuda.current_stream(), self._cw_stream
        ev = torch.cuda.Event()
        ev.record(cur)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            for dst, src in self._cw_pairs:
                dst.copy_(src, non_blocking=True)
            self._cw_symm.barrier(channel=1)     # every rank's copies INTO this rank have landed
            self._cw_event.record(side)

    def recompile(self, optimizer=None, lr=None, margin=None, wver=None, wid=None, triplet_hard=None, **opt_kw):
        """model.compile(...) on an existing model: new loss constants and a FRESH optimiser (Keras creates new slot
        variables and restarts `iterations`); the weights stay.  Captured step graphs bake the constants in: dropped."""
        import dataclasses
        ch = {k: v for k, v in dict(margin=margin, wver=wver, wid=wid, triplet_hard=triplet_hard).items() if v is not None}
        if ch:
            self.cfg = dataclasses.replace(self.cfg, **ch)
        if optimizer is not None:
            self.optimizer = optimizer
        if lr is not None:
            self.lr = float(lr)
        for k in ("momentum", "beta1", "beta2", "eps", "lr_decay"):
            if k in opt_kw and opt_kw[k] is not None:
                setattr(self, k, float(opt_kw[k]))
        if opt_kw.get("decoupled_weight_decay") is not None:
            self.decoupled_wd = float(opt_kw["decoupled_weight_decay"])
        torch.cuda.synchronize(self.dev)
        self.m.zero_()
        self.v.zero_()
        if getattr(self, "vhat", None) is not None:
            self.vhat.zero_()
        self.t = 0
        self._graphs.clear()

    def _next_lr(self):
        self.t += 1
        if self.optimizer in ("adam", "amsgrad", "adamw"):
            lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        else:
            lr_t = self.lr / (1.0 + self.lr_decay * (self.t - 1))      # Keras: iterations counts finished steps
        self._lr_host[0] = lr_t
        self.lr_dev.copy_(self._lr_host, non_blocking=True)

    @torch.no_grad()
    def loss_and_grad(self, inputs, flags, labels, drop_masks=None, code_drop_mask=None) -> Dict[str, torch.Tensor]:
        """Forward + losses + full backward into the gradient arena, no optimiser step.
        Gradients exclude the L2-regulariser terms (those are applied inside the optimiser
        kernel, ugn_adam_step)."""
        B = int(inputs[0].shape[0])
        p = self.plan(B, True)
        self._set_inputs(p, inputs, flags, labels, drop_masks, code_drop_mask)
        self._step_body(p, False)
        return self._report(p)

    def _set_base_inputs(self, p, base_inputs, src_row, use, labels, mirror=None, shift=None, clip=None):
        """Stage the BASE rows (+ the expansion pattern) of train_step_expanded / predict_expanded."""
        cfg = self.cfg
        B0 = int(base_inputs[0].shape[0])
        p.ensure_base(B0)
        for m in range(cfg.nmods):
            p.br[m].x_base.copy_(base_inputs[m], non_blocking=True)
            if not cfg.single:
                p.flags[m].copy_(torch.as_tensor(use[:, m]).reshape(-1, 1), non_blocking=True)
        self._draw_dropout(p)
        p.src_row.copy_(torch.as_tensor(src_row, dtype=torch.int32), non_blocking=True)
        p.use_mirror = mirror is not None
        if mirror is not None:
            p.mirror.copy_(torch.as_tensor(mirror, dtype=torch.uint8), non_blocking=True)
        p.use_augment = shift is not None or clip is not None
        if p.use_augment:
            B = p.B
            p.shift.copy_(torch.as_tensor(shift if shift is not None else torch.zeros(B, 2), dtype=torch.int8), non_blocking=True)
            p.clip.copy_(torch.as_tensor(clip if clip is not None else torch.zeros(B), dtype=torch.uint8), non_blocking=True)
        if labels is not None:
            lab = torch.as_tensor(labels).reshape(-1).to(torch.int32)
            if lab.numel() == B0:        # one label per base row: replicate along the expansion
                lab = lab.to(self.dev)[p.src_row.long()]
            p.labels.copy_(lab, non_blocking=True)

    @torch.no_grad()
    def train_step_expanded(self, base_inputs, base_labels, src_row, use, mirror=None, shift=None,
                            clip=None) -> Dict[str, torch.Tensor]:
        """train_step on the reference generator's E-fold batch WITHOUT materialising it: base_inputs are
        the B0 sequences that have every modality ([B0,C,60,60] per modality), (src_row, use) is the
        expansion pattern (ugaitnet_b200.expand.expansion_pattern) and the volumes are expanded while
        they are packed on the device -- only the base rows cross PCIe."""
        B = int(len(src_row))
        p = self.plan(B, True)
        self._set_base_inputs(p, base_inputs, src_row, use, base_labels, mirror, shift, clip)
        return self._run_train(p, B, expanded=True)

    @torch.no_grad()
    def predict_expanded(self, base_inputs, src_row, use, mirror=None, layer: str = "signature") -> torch.Tensor:
        """Descriptor extraction with device-side expansion / mirror augmentation
        (mains/mj_testUWYHGaitNet_open_tum.py:174-190 stacks the mirrored copies on the host)."""
        B = int(len(src_row))
        p = self.plan(B, False)
        self._set_base_inputs(p, base_inputs, src_row, use, None, mirror)
        self._forward(p, False, expanded=True)
        return self._layer_output(p, layer)

    @torch.no_grad()
    def train_step(self, inputs, flags, labels, drop_masks=None, code_drop_mask=None) -> Dict[str, torch.Tensor]:
        """One Keras train_function step: fwd -> losses -> bwd -> (all-reduce) -> optimiser."""
        B = int(inputs[0].shape[0])
        p = self.plan(B, True)
        self._set_inputs(p, inputs, flags, labels, drop_masks, code_drop_mask)
        return self._run_train(p, B, expanded=False)

    def input_buffers(self, B: int, train: bool = True):
        """(volumes [B,C,60,60] per modality, use-flags [B,1] per modality, labels i32 [B]): the plan's OWN input block as
        device tensors.  A producer that already runs on the GPU (a device-side loader, a previous kernel) writes the
        batch here and calls train_step_resident() / predict_resident(): no per-step input copy at all."""
        p = self.plan(B, train)
        return [b.x_in for b in p.br], list(p.flags), p.labels

    @torch.no_grad()
    def train_step_resident(self, B: int) -> Dict[str, torch.Tensor]:
        """train_step on whatever input_buffers(B) currently hold."""
        p = self.plan(B, True)
        self._draw_dropout(p)
        return self._run_train(p, B, expanded=False)

    @torch.no_grad()
    def predict_resident(self, B: int, layer: str = "signature") -> torch.Tensor:
        p = self.plan(B, False)
        self._forward(p, False)
        return self._layer_output(p, layer)

    def _run_train(self, p, B, expanded):
        self._next_lr()
        if self.use_graph and (self.world == 1 or self.dp_graph):
            gkey = (B, expanded, p.use_mirror, getattr(p, "_B0", None) if expanded else None,
                    bool(getattr(p, "use_augment", False)) if expanded else False,
                    bool(getattr(p, "use_philox", False)), bool(getattr(p, "philox_code", False)))
            gr = self._graphs.get(gkey)
            if gr is None:
                # warm-up on a side stream (first-use allocations / attribute sets), then capture
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                saved = (self.w.clone(), self.m.clone(), self.v.clone())
                with torch.cuda.stream(s):
                    self._step_body(p, True, expanded)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                self.w.copy_(saved[0]); self.m.copy_(saved[1]); self.v.copy_(saved[2])
                self.repack_weights()
                if self.world > 1 or self.force_segments:
                    gr = self._capture_segments(p, expanded)
                else:
                    gr = torch.cuda.CUDAGraph()
                    l0 = self.ctx.launches
                    import gc
                    gc_on = gc.isenabled()
                    gc.disable()                     # (see _capture_segments)
                    try:
                        with torch.cuda.graph(gr):
                            self._step_body(p, True, expanded)
                    finally:
                        if gc_on:
                            gc.enable()
                    self.graph_launches = self.ctx.launches - l0   # kernels of ours inside one replay
                self._graphs[gkey] = gr
                # the capture itself did not execute: fall through to replay
            if isinstance(gr, list):
                self._replay_segments(gr)
            else:
                gr.replay()
        else:
            self._step_body(p, True, expanded)
        return self._report(p, with_reg=True)

    @torch.no_grad()
    def eval_losses(self, inputs, flags, labels) -> Dict[str, torch.Tensor]:
        """Forward in inference mode + the two losses (Keras test_function / validation pass)."""
        cfg, h = self.cfg, self.ctx.h
        B = int(inputs[0].shape[0])
        p = self.plan(B, True)
        self._set_inputs(p, inputs, flags, labels)
        sig, _ = self._forward(p, False)
        st = stream_ptr()
        if getattr(cfg, "pair_loss", False):
            check(lib.ugn_pair_verif_loss(h, sig.ptr, p.R["labels"].ptr, cfg.margin, 1.0, p.R["trip_out"].ptr, None,
                                          p.R["trip_ws"].ptr, st))
        elif getattr(cfg, "triplet_hard", False):
            check(lib.ugn_triplet_hard(h, sig.ptr, None, p.R["labels"].ptr, cfg.margin, 1.0, p.R["trip_out"].ptr, None,
                                       p.R["trip_ws"].ptr, st))
        else:
            check(lib.ugn_triplet_all(h, sig.ptr, p.R["labels"].ptr, cfg.margin, 1.0, p.R["trip_out"].ptr, None,
                                      p.R["trip_ws"].ptr, st))
        if cfg.nclasses > 0:
            check(lib.ugn_softmax_ce_ls(h, p.R["logits"].ptr, p.R["labels"].ptr, p.R["ce_out"].ptr, None, 1.0,
                                        cfg.label_smoothing, st))
            if getattr(self, "aux", False):
                for b in p.br:
                    check(lib.ugn_softmax_ce_ls(h, b.R["aux_logits"].ptr, p.R["labels"].ptr, b.R["aux_ce"].ptr, None, 1.0,
                                                cfg.label_smoothing, st))
        return self._report(p)

    # ------------------------------------------------------------------ host -> device pipelining
    def prefetch(self, inputs, flags, labels):
        """Enqueue the H2D copy of the NEXT step's (pinned) host batch on a side stream into the alternate
        staging set, so it overlaps the step that is running; consume it with train_step_staged()."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._staging = [None, None]
            self._stage_evt = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_k = 0
        k = self._stage_k ^ 1
        B = int(inputs[0].shape[0])
        st = self._staging[k]
        if st is None or st[0][0].shape[0] != B:
            st = ([torch.empty(tuple(x.shape), device=self.dev) for x in inputs],
                  None if flags is None else [torch.empty(B, 1, device=self.dev) for _ in flags],
                  torch.empty(B, dtype=torch.int32, device=self.dev))
            self._staging[k] = st
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free[k])      # previous consumer of this set is done
            for d, h in zip(st[0], inputs):
                d.copy_(h, non_blocking=True)
            if flags is not None:
                for d, h in zip(st[1], flags):
                    d.copy_(h.reshape(-1, 1), non_blocking=True)
            st[2].copy_(labels.reshape(-1).to(torch.int32), non_blocking=True)
            self._stage_evt[k].record(self._copy_stream)
        self._stage_k = k

    def train_step_staged(self):
        """train_step on the batch most recently passed to prefetch()."""
        k = self._stage_k
        torch.cuda.current_stream().wait_event(self._stage_evt[k])
        st = self._staging[k]
        out = self.train_step(st[0], st[1], st[2])
        self._stage_free[k].record(torch.cuda.current_stream())
        return out

    # ---- single-copy input path: the loader writes into a pinned HostBatch whose bytes mirror the plan's IOBlock
    def host_batch(self, B: int, base_rows: Optional[int] = None, train: bool = True, raw=None) -> HostBatch:
        """A NEW pinned host batch for batch size B (base_rows = B0 selects the device-side-expansion layout: only the
        B0 complete sequences + the expansion tables cross PCIe).  Allocate two and alternate them to overlap the
        loader / the H2D copy of step i+1 with step i.  raw: the volumes hold the STORED sample integers (one spec per
        modality, ugaitnet_b200.samples.RAW_*) and are decoded on the device -- see HostBatch."""
        return HostBatch(self.plan(B, train).io, base_rows, raw)

    def prefetch_batch(self, hb: HostBatch, train: bool = True):
        """ONE cudaMemcpyAsync of the whole batch (flags, labels, expansion tables, volumes) on the copy stream into
        the alternate device staging block; consume with train_step_prefetched() / predict_prefetched()."""
        p = self.plan(hb.io.B, train)
        if not hasattr(self, "_io_stage"):
            self._io_copy_stream = torch.cuda.Stream(device=self.dev)
            self._io_stage, self._io_k = {}, 0
        key = (hb.io.B, train)
        st = self._io_stage.get(key)
        if st is None:
            st = self._io_stage[key] = {"buf": [torch.empty(p.io.nbytes_full, dtype=torch.uint8, device=self.dev) for _ in range(2)],
                                        "ready": [torch.cuda.Event(), torch.cuda.Event()],
                                        "free": [torch.cuda.Event(), torch.cuda.Event()], "hb": [None, None]}
        k = self._io_k ^ 1
        with torch.cuda.stream(self._io_copy_stream):
            self._io_copy_stream.wait_event(st["free"][k])           # the step that consumed this block is done
            st["buf"][k][:hb.nbytes].copy_(hb.buf, non_blocking=True)
            st["ready"][k].record(self._io_copy_stream)
        st["hb"][k] = hb
        self._io_k, self._io_key = k, key

    # The same copy in PREFIX pieces, for a loader that fills the block front to back (header, then one volume after the
    # other): prefetch_open(hb); fill header + volume 0; prefetch_upto(io.x_off[1]); fill volume 1; ...; prefetch_close().
    # Each piece is a cudaMemcpyAsync that runs while the host produces the next one (the compat model's f64 -> f32 cast
    # of modality m+1 overlaps the H2D copy of modality m).  The bytes that arrive are those of prefetch_batch(hb).
    def prefetch_open(self, hb: HostBatch, train: bool = True):
        p = self.plan(hb.io.B, train)
        if not hasattr(self, "_io_stage"):
            self._io_copy_stream = torch.cuda.Stream(device=self.dev)
            self._io_stage, self._io_k = {}, 0
        key = (hb.io.B, train)
        st = self._io_stage.get(key)
        if st is None:
            st = self._io_stage[key] = {"buf": [torch.empty(p.io.nbytes_full, dtype=torch.uint8, device=self.dev) for _ in range(2)],
                                        "ready": [torch.cuda.Event(), torch.cuda.Event()],
                                        "free": [torch.cuda.Event(), torch.cuda.Event()], "hb": [None, None]}
        k = self._io_k ^ 1
        self._io_copy_stream.wait_event(st["free"][k])               # the step that consumed this block is done
        self._io_open = [st, k, hb, key, 0]

    def prefetch_upto(self, end: int):
        """Copy the not yet copied bytes below `end` of the HostBatch given to prefetch_open (asynchronous)."""
        st, k, hb, _, done = self._io_open
        end = min(int(end), hb.nbytes)
        if end > done:
            with torch.cuda.stream(self._io_copy_stream):
                st["buf"][k][done:end].copy_(hb.buf[done:end], non_blocking=True)
            self._io_open[4] = end

    def prefetch_close(self):
        st, k, hb, key, _ = self._io_open
        self.prefetch_upto(hb.nbytes)
        st["ready"][k].record(self._io_copy_stream)
        st["hb"][k] = hb
        self._io_k, self._io_key = k, key
        self._io_open = None

    def _consume_prefetched(self):
        key, k = self._io_key, self._io_k
        st = self._io_stage[key]
        hb = st["hb"][k]
        p = self.plan(*key)
        cur = torch.cuda.current_stream()
        cur.wait_event(st["ready"][k])
        if hb.raw is not None:
            # stored integers: the header moves as bytes, the volumes are DECODED into the plan's f32 block
            # (ugn_decode_samples: the generator's __load_dd arithmetic, bit for bit) instead of copied
            p.io.dev_buf[:p.io.header].copy_(st["buf"][k][:p.io.header], non_blocking=True)
            src = p.io.views(st["buf"][k], hb.B0, raw=hb.raw)["x"]
            dst = p.io.views(p.io.dev_buf, hb.B0)["x"]
            h, sp = self.ctx.h, stream_ptr()
            cache = st.setdefault("raw_refs", {})
            ck = (k, hb.B0, tuple(r[0] for r in hb.raw))
            refs = cache.get(ck)
            if refs is None:        # DLPack handles of the (fixed) staging / plan views, created once
                refs = cache[ck] = [(TRef(a), TRef(b)) for a, b in zip(src, dst)]
            for (ra, rb), r in zip(refs, hb.raw):
                check(lib.ugn_decode_samples(h, ra.ptr, r[1], r[2], r[3], r[4], r[5], rb.ptr, sp))
        else:
            p.io.dev_buf[:hb.nbytes].copy_(st["buf"][k][:hb.nbytes], non_blocking=True)   # ONE device-to-device copy
        st["free"][k].record(cur)
        return p, hb

    def _draw_dropout(self, p):
        cfg = self.cfg
        p.use_philox = self.philox and p.train and cfg.dropout > 0.001 and any(hasattr(b, "mask") for b in p.br)
        p.philox_code = self.philox and p.train and cfg.dropout > 0.001 and cfg.nc > 0
        if self.philox:
            return
        if p.train and cfg.dropout > 0.001:
            keep = 1.0 - cfg.dropout
            for b in p.br:
                if hasattr(b, "mask"):
                    b.mask.bernoulli_(keep).div_(keep)
            if cfg.nc > 0:
                p.cmask.bernoulli_(keep).div_(keep)

    @torch.no_grad()
    def train_step_prefetched(self) -> Dict[str, torch.Tensor]:
        """train_step on the HostBatch most recently passed to prefetch_batch()."""
        p, hb = self._consume_prefetched()
        self._draw_dropout(p)
        expanded = hb.B0 is not None
        if expanded:
            p.ensure_base(hb.B0)
            p.use_mirror, p.use_augment = bool(hb.use_mirror), bool(hb.use_augment)
        return self._run_train(p, p.B, expanded=expanded)

    @torch.no_grad()
    def predict_prefetched(self, layer: str = "signature") -> torch.Tensor:
        p, hb = self._consume_prefetched()
        expanded = hb.B0 is not None
        if expanded:
            p.ensure_base(hb.B0)
            p.use_mirror = bool(hb.use_mirror)
        self._forward(p, False, expanded)
        return self._layer_output(p, layer)

    def _report(self, p: "_Plan", with_reg: bool = False) -> Dict[str, torch.Tensor]:
        out = {"triplet": p.trip_out[0], "count": p.trip_out[1],
               "signature": p.br[0].out if self.cfg.single else (p.codeN if self.post2 else p.sig)}
        if self.cfg.nclasses > 0:
            out["ce"], out["acc"], out["logits"] = p.ce_out[0], p.ce_out[1], p.logits
        if getattr(self, "aux", False) and p.train:
            out["aux_ce"] = [b.T["aux_ce"][0] for b in p.br]
            out["aux_acc"] = [b.T["aux_ce"][1] for b in p.br]
        if with_reg:
            out["reg"] = p.loss_pack[4]
            out["losses"] = p.loss_pack          # [triplet, count, ce, acc, reg, activity reg, -, -]: one D2H read
            if self.cfg.nc > 0:
                out["act_reg"] = p.loss_pack[5]
        return out

    def launches_per_step(self) -> int:
        return self.ctx.launches


class _Branch:
    pass


class _Plan:
    """All activation / gradient buffers of one batch size, exported once through DLPack."""

    def __init__(self, eng: UGaitEngine, B: int, train: bool):
        cfg, d, P, PB, dt16 = eng.cfg, eng.dev, eng.P, eng.PB, eng.dt16
        self.B, self.train = B, train
        f32 = dict(device=d, dtype=torch.float32)

        def act(shape, planes=P):
            if planes:
                return torch.zeros((planes,) + tuple(shape), device=d, dtype=dt16)
            return torch.zeros(tuple(shape), **f32)

        self.eng = eng
        self.use_mirror = False
        self.use_philox = self.philox_code = False
        self.br: List[_Branch] = []
        # every per-step input lives in ONE device block (IOBlock): one H2D copy per step, see UGaitEngine.prefetch_batch
        self.io = IOBlock(d, B, [(cfg.in_channels[m], cfg.hw, cfg.hw) for m in range(cfg.nmods)])
        iov = self.io.views(self.io.dev_buf)
        self.flags = iov["flags"]
        for f in self.flags:
            f.fill_(1.0)
        self._base = {}
        for m in range(cfg.nmods):
            b = _Branch()
            if cfg.is3d(m):
                self.br.append(self._branch3d(eng, b, m, iov["x"][m], train))
                continue
            b.layers = cfg.layers(m, eng.pad)
            T = {}
            L0 = b.layers[0]
            b.x_in = T["x_in"] = iov["x"][m]
            T["a0"] = act((B, L0["h"], L0["h"], L0["cp"]))
            for li, L in enumerate(b.layers):
                T[f"a{li + 1}"] = act((B, L["hp"], L["hp"], L["co"]))
                if L["pool"]:
                    T[f"idx{li}"] = torch.zeros(B, L["hp"], L["hp"], L["co"], device=d, dtype=torch.uint8)
                if train:
                    T[f"dz{li}c"] = act((B, L["ho"], L["ho"], L["co"]), PB)
                    T[f"da{li + 1}"] = torch.zeros(B, L["hp"], L["hp"], L["co"], **f32)
            T["flat"] = act((B, cfg.flat))
            b.h1 = T["h1"] = torch.zeros(B, 2 * cfg.nd, **f32)
            if P:
                T["h1_16"] = torch.zeros(P, B, 2 * cfg.nd, device=d, dtype=dt16)
            b.out = T["out"] = torch.zeros(B, cfg.nd, **f32)
            if train:
                b.mask = T["mask"] = torch.ones(B, 2 * cfg.nd, **f32)
                b.dout = T["dout"] = torch.zeros(B, cfg.nd, **f32)
                T["dh1"] = torch.zeros(B, 2 * cfg.nd, **f32)
                T["dflat"] = torch.zeros(B, cfg.flat, **f32)
                if P:
                    T["dout16"] = torch.zeros(PB, B, cfg.nd, device=d, dtype=dt16)
                    T["dz1_16"] = torch.zeros(PB, B, 2 * cfg.nd, device=d, dtype=dt16)
                else:
                    T["dz1"] = torch.zeros(B, 2 * cfg.nd, **f32)
            if cfg.normbfmerge and not cfg.single:
                T["outn"] = torch.zeros(B, cfg.nd, **f32)
                T["nwin"] = torch.zeros(B, cfg.nd, device=d, dtype=torch.uint8)
                T["ninv"] = torch.zeros(B, 2, **f32)
                if train:
                    T["doutn"] = torch.zeros(B, cfg.nd, **f32)
            if eng.aux:
                T["gated"] = torch.zeros(B, cfg.nd, **f32)
                T["gwin"] = torch.zeros(B, cfg.nd, device=d, dtype=torch.uint8)
                T["ginv"] = torch.zeros(B, 2, **f32)
                T["aux_logits"] = torch.zeros(B, cfg.nclasses, **f32)
                if train:
                    T["aux_ce"] = torch.zeros(2, **f32)
                    T["daux_logits"] = torch.zeros(B, cfg.nclasses, **f32)
                    T["dgated"] = torch.zeros(B, cfg.nd, **f32)
                    T["daux"] = torch.zeros(B, cfg.nd, **f32)
            b.T = T
            b.R = {k: TRef(v) for k, v in T.items()}
            if eng.aux:
                b.gate_in = ptr_array([b.R["outn" if cfg.normbfmerge else "out"]])
                b._flag1 = TRef(self.flags[m])
                b.flag1 = ptr_array([b._flag1])
                if train:
                    b.aux_dout = ptr_array([b.R["daux"]])
            if cfg.normbfmerge and not cfg.single:
                b.nrm_in = ptr_array([b.R["out"]])
                if train:
                    b.nrm_dout = ptr_array([b.R["dout"]])
            self.br.append(b)
        T = {}
        self.sig = T["sig"] = torch.zeros(B, cfg.nd, **f32)
        # tensor-core modes: the fusion kernel also writes the signature's hi/lo planes, the triplet's Gram GEMM reads them
        self.tc_gram = bool(P) and not cfg.single and not eng.post2 and cfg.nd % 64 == 0 and \
            os.environ.get("UGN_TC_GRAM", "1") == "1"
        if self.tc_gram:
            T["sig16"] = torch.zeros(2, B, cfg.nd, device=d, dtype=dt16)
        T["winner"] = torch.zeros(B, cfg.nd, device=d, dtype=torch.uint8)
        T["inv_norm"] = torch.zeros(B, 2, **f32)
        feat = cfg.nd
        if cfg.nc > 0:
            self.code = T["code"] = torch.zeros(B, cfg.nc, **f32)
            self.dropcode = T["dropcode"] = torch.zeros(B, cfg.nc, **f32)
            self.cmask = T["cmask"] = torch.ones(B, cfg.nc, **f32)
            feat = cfg.nc
            if eng.post2:
                self.codeN = T["codeN"] = torch.zeros(B, cfg.nc, **f32)
                T["cwin"] = torch.zeros(B, cfg.nc, device=d, dtype=torch.uint8)
                T["cinv"] = torch.zeros(B, 2, **f32)
                if train:
                    self.dcodeN = T["dcodeN"] = torch.zeros(B, cfg.nc, **f32)
        if cfg.nclasses > 0:
            self.logits = T["logits"] = torch.zeros(B, cfg.nclasses, **f32)
        self.labels = T["labels"] = iov["labels"]
        self.src_row = T["src_row"] = iov["src_row"]
        self.mirror = T["mirror"] = iov["mirror"]
        self.shift, self.clip = iov["shift"], iov["clip"]
        T["shift"], T["clip"] = self.shift, self.clip
        self.use_augment = False
        if train:
            # {triplet, count, ce, acc, reg}: one buffer, one D2H read per step
            self.loss_pack = T["loss_pack"] = torch.zeros(8, **f32)
            self.trip_out = T["trip_out"] = self.loss_pack[0:2]
            self.dsig = T["dsig"] = torch.zeros(B, cfg.nd, **f32)
            T["trip_ws"] = torch.zeros(ops.triplet_workspace_bytes(1, B) // 4 + 16, **f32)
            if cfg.nclasses > 0:
                self.ce_out = T["ce_out"] = self.loss_pack[2:4]
                T["dlogits"] = torch.zeros(B, cfg.nclasses, **f32)
                self.dfeat = T["dfeat"] = torch.zeros(B, feat, **f32)
            if cfg.nc > 0:
                self.dcode = T["dcode"] = torch.zeros(B, cfg.nc, **f32)
                self.dcode_z = T["dcode_z"] = torch.zeros(B, cfg.nc, **f32)
                self.dsig2_code = torch.zeros(B, cfg.nc, **f32)
                self.act_sq = torch.zeros(B, cfg.nc, **f32)
                self.dsig2 = T["dsig2"] = torch.zeros(B, cfg.nd, **f32)
        self.T = T
        self.R = {k: TRef(v) for k, v in T.items()}
        self.R_flags = [TRef(f) for f in self.flags]
        self.br_ptrs = ptr_array([b.R["out"] for b in self.br])
        self.flag_ptrs = ptr_array(self.R_flags)
        if train:
            self.dbr_ptrs = ptr_array([b.R["dout"] for b in self.br])
        if eng.post2:
            self.code_ptrs = ptr_array([self.R["code"]])
            if train:
                self.dcode_ptrs = ptr_array([self.R["dcode"]])
        if (cfg.normbfmerge and not cfg.single) or eng.post2:
            self._ones = TRef(torch.ones(B, 1, **f32))
            self.one_ptrs = ptr_array([self._ones])
        if cfg.normbfmerge and not cfg.single:
            self.brn_ptrs = ptr_array([b.R["outn"] for b in self.br])
            if train:
                self.dbrn_ptrs = ptr_array([b.R["doutn"] for b in self.br])

    def _branch3d(self, eng, b, m, x_in, train):
        """Buffers of a Conv3D branch (use3D): activations [B,T,H,W,C] f32, the gradient of every pre-activation, flat views."""
        cfg, B = eng.cfg, self.B
        f32 = dict(device=eng.dev, dtype=torch.float32)
        b.layers, b.layers3d = [], cfg.layers3d(m)
        T = {}
        b.x_in = T["x_in"] = x_in                                        # [B,25,60,60] == channels-last [B,25,60,60,1]
        T["a0"] = x_in.view(B, cfg.in_channels[m], cfg.hw, cfg.hw, 1)
        for li, L in enumerate(b.layers3d):
            shp = (B, L["to"], L["ho"], L["ho"], L["co"])
            T[f"a{li + 1}"] = torch.zeros(shp, **f32)
            T[f"a{li + 1}f"] = T[f"a{li + 1}"].view(-1, L["co"])
            if train:
                T[f"dz{li}"] = torch.zeros(shp, **f32)
                T[f"dz{li}f"] = T[f"dz{li}"].view(-1, L["co"])
                T[f"da{li + 1}"] = torch.zeros(shp, **f32)
                T[f"da{li + 1}f"] = T[f"da{li + 1}"].view(-1, L["co"])
        nl = len(b.layers3d)
        T["flat"] = T[f"a{nl}"].view(B, -1)
        b.out = T["out"] = torch.zeros(B, cfg.nd, **f32)
        if train:
            b.dout = T["dout"] = torch.zeros(B, cfg.nd, **f32)
            T["dflat"] = T[f"da{nl}"].view(B, -1)
        if (cfg.normbfmerge and not cfg.single) or eng.aux:
            raise NotImplementedError("use3D together with normbfmerge / aux_losses")
        b.T = T
        b.R = {k: TRef(v) for k, v in T.items()}
        return b

    def ensure_base(self, B0: int):
        """Views of the device-side expansion's base rows (B0 rows per modality, packed right behind the header of the
        IO block).  Nothing is allocated or freed here: a CUDA graph captured for one B0 keeps reading valid memory
        when another B0 is used in between (the graph key carries B0)."""
        assert 0 < B0 <= self.B, "base rows must be in [1, B]"
        ent = self._base.get(B0)
        if ent is None:
            xs = self.io.views(self.io.dev_buf, B0)["x"]
            ent = self._base[B0] = (xs, [TRef(x) for x in xs])
        for b, x, r in zip(self.br, ent[0], ent[1]):
            b.x_base, b.R["x_base"] = x, r
        self._B0 = B0
