"""Open-world k-NN classifier with the call surface the reference test scripts use:

    clf = KNeighborsClassifier(n_neighbors=knn); clf.fit(G, y); clf.predict(Q)
    (/root/reference/mains/mj_testUWYHGaitNet_open_tum.py:331-341)

Semantics follow scikit-learn's defaults for that call (Euclidean, uniform vote, vote ties ->
smallest label).  Neighbours are ordered by (exact fp64 distance, gallery index).
The gallery can be row-sharded across the ranks of a ``torch.distributed`` process group: every
rank searches its shard and the per-rank top-k lists are merged after one all-gather.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops


def _as_cuda(a, dtype, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device, non_blocking=True)


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Row range [lo, hi) of shard `rank` out of `world` (contiguous, sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class KNeighborsClassifier:
    def __init__(self, n_neighbors: int = 5, device: Optional[int] = None, process_group=None,
                 query_block: int = 4096):
        if not torch.cuda.is_available():
            raise RuntimeError("ugaitnet_b200 needs a CUDA device (no CPU fallback)")
        self.k = int(n_neighbors)
        self.dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.ctx = ops.get_ctx(self.dev.index)
        self.pg = process_group
        self.query_block = query_block
        self.idx_base = 0
        # candidate scan on the tensor cores (tcgen05 distance GEMM + fused top-k) unless disabled
        self.use_tc = self.ctx.has_tcgen05 and not os.environ.get("UGN_KNN_SIMT")
        self.flagged = 0            # queries whose candidate-containment proof failed (recomputed exactly)
        # replay the search of a query count as CUDA graphs from its second call on (the fixed per-call cost -- a dozen
        # launches and their host work -- is what limits strong scaling once a shard's scan takes ~1 ms)
        self.use_graph = os.environ.get("UGN_KNN_GRAPH", "1") == "1"
        self._works = {}

    # ---- fit: keep (this rank's shard of) the gallery resident in HBM ------------------------
    def fit(self, X, y, idx_base: int = 0, sharded: bool = False):
        """X [N,D] float, y [N] int labels.  With a process group and ``sharded=False`` the full
        gallery is given on every rank and this rank keeps rows shard_bounds(N, rank, world);
        with ``sharded=True`` X/y already are this rank's shard and idx_base its global offset."""
        # input validation with scikit-learn's exceptions and wording (check_X_y / check_array of KNeighborsClassifier.fit)
        if len(np.shape(X)) != 2:
            raise ValueError(f"Expected 2D array, got {len(np.shape(X))}D array instead")
        if np.shape(X)[0] != np.shape(y)[0]:
            raise ValueError(f"Found input variables with inconsistent numbers of samples: [{np.shape(X)[0]}, {np.shape(y)[0]}]")
        if np.shape(X)[0] == 0:
            raise ValueError(f"Found array with 0 sample(s) (shape={tuple(np.shape(X))}) while a minimum of 1 is required "
                             "by KNeighborsClassifier.")
        self.n_total = int(np.shape(X)[0])
        if self.pg is not None and not sharded:
            rank, world = torch.distributed.get_rank(self.pg), torch.distributed.get_world_size(self.pg)
            lo, hi = shard_bounds(len(X), rank, world)
            X, y, idx_base = X[lo:hi], y[lo:hi], lo
        self.G = _as_cuda(X, torch.float32, self.dev)
        self.labels = _as_cuda(y, torch.int32, self.dev)
        self.idx_base = int(idx_base)
        if self.pg is not None and sharded:
            tot = torch.tensor([self.G.shape[0]], dtype=torch.int64, device=self.dev)
            torch.distributed.all_reduce(tot, group=self.pg)
            self.n_total = int(tot)
        if not bool(torch.isfinite(self.G).all()):
            raise ValueError("Input X contains NaN." if bool(torch.isnan(self.G).any()) else
                             "Input X contains infinity or a value too large for dtype('float32').")
        if self.G.shape[0] < self.k and self.n_total >= self.k:
            raise ValueError(f"gallery shard of {self.G.shape[0]} rows is smaller than n_neighbors = {self.k}: use fewer ranks")
        n, d = self.G.shape
        npad = (n + 255) // 256 * 256
        self.g2 = torch.empty(npad, device=self.dev)
        self.gmax2 = torch.zeros(1, device=self.dev)
        ops.knn_gallery_norms(self.ctx, self.G, self.g2, self.gmax2)
        if self.use_tc:
            self.kch = (d + 63) // 64
            self.dp = self.kch * 64
            self.G16 = torch.empty(2, self.kch, n, 64, dtype=torch.float16, device=self.dev)
            ops.knn_pack(self.ctx, self.G, self.G16)        # fp16 hi/lo planes, K-chunk major, zero padded
        self.classes_ = None
        self._works = {}
        if hasattr(self, "_RG"):
            del self._RG
        return self

    # ---- per-query-count workspace: every buffer of a search allocated once, DLPack handles cached -------------
    def _work(self, nq: int):
        w = self._works.get(nq)
        if w is not None:
            return w
        from ._ffi import TRef
        k, dev = self.k, self.dev
        n = nq * k
        world = torch.distributed.get_world_size(self.pg) if self.pg is not None else 1
        row = (20 * n + 7) // 8 * 8
        w = {"nq": nq, "row": row, "world": world}
        # this rank's lists back to back in ONE byte buffer: d2 f64 | idx i64 | lab i32  (-> one all-gather)
        w["pack"] = torch.zeros(row, dtype=torch.uint8, device=dev)
        w["d2"] = w["pack"][:8 * n].view(torch.float64).view(nq, k)
        w["idx"] = w["pack"][8 * n:16 * n].view(torch.int64).view(nq, k)
        w["lab"] = w["pack"][16 * n:20 * n].view(torch.int32).view(nq, k)
        w["gathered"] = torch.zeros(world, row, dtype=torch.uint8, device=dev) if world > 1 else w["pack"].view(1, row)
        w["od2"] = torch.empty(nq, k, dtype=torch.float64, device=dev)
        w["oidx"] = torch.empty(nq, k, dtype=torch.int64, device=dev)
        w["olab"] = torch.empty(nq, k, dtype=torch.int32, device=dev)
        w["pred"] = torch.empty(nq, dtype=torch.int32, device=dev)
        w["q"] = torch.empty(nq, self.G.shape[1], device=dev)
        blocks = []
        for s in range(0, nq, self.query_block):
            e = min(nq, s + self.query_block)
            b = {"s": s, "e": e,
                 "ws": torch.empty(max(ops.knn_workspace_bytes(e - s, self.G.shape[0], self.G.shape[1], k) // 4, 4),
                                   dtype=torch.float32, device=dev)}
            if self.use_tc:
                b["q16"] = torch.empty(2, self.kch, e - s, 64, dtype=torch.float16, device=dev)
            b["R"] = {n_: TRef(t) for n_, t in dict(q=w["q"][s:e], d2=w["d2"][s:e], idx=w["idx"][s:e], lab=w["lab"][s:e],
                                                    ws=b["ws"], **({"q16": b["q16"]} if self.use_tc else {})).items()}
            blocks.append(b)
        w["blocks"] = blocks
        w["flags"] = torch.zeros(nq, dtype=torch.int32, device=dev) if self.use_tc else None
        if self.use_tc:
            for b in blocks:
                b["R"]["flags"] = TRef(w["flags"][b["s"]:b["e"]])
        w["R"] = {n_: TRef(w[n_]) for n_ in ("gathered", "od2", "oidx", "olab", "pred")}
        if not hasattr(self, "_RG"):
            self._RG = {n_: TRef(t) for n_, t in dict(G=self.G, g2=self.g2, gmax2=self.gmax2, labels=self.labels,
                                                      **({"G16": self.G16} if self.use_tc else {})).items()}
        w["graph"] = None
        self._works[nq] = w
        return w

    # ---- local shard search -------------------------------------------------------------------
    def _local_topk_into(self, w):
        """Candidate scan + exact re-rank of w["q"] into the packed list buffer (kernels only: capturable)."""
        from ._ffi import check, lib, stream_ptr
        h, st, G, k = self.ctx.h, stream_ptr(), self._RG, self.k
        if w["flags"] is not None:
            w["flags"].zero_()
        for b in w["blocks"]:
            R = b["R"]
            if self.use_tc:
                check(lib.ugn_knn_pack(h, R["q"].ptr, R["q16"].ptr, st))
                check(lib.ugn_knn_topk_tc(h, R["q"].ptr, R["q16"].ptr, G["G"].ptr, G["G16"].ptr, G["g2"].ptr,
                                          G["gmax2"].ptr, G["labels"].ptr, k, self.idx_base, R["d2"].ptr, R["idx"].ptr,
                                          R["lab"].ptr, R["flags"].ptr, R["ws"].ptr, st))
            else:
                check(lib.ugn_knn_topk(h, R["q"].ptr, G["G"].ptr, G["g2"].ptr, G["labels"].ptr, k, self.idx_base,
                                       R["d2"].ptr, R["idx"].ptr, R["lab"].ptr, R["ws"].ptr, st))

    def _local_topk(self, Q: torch.Tensor):
        w = self._work(Q.shape[0])
        w["q"].copy_(Q)
        self._local_topk_into(w)
        self._flags = w["flags"]
        return w["d2"], w["idx"], w["lab"]

    def flagged_queries(self) -> int:
        """How many queries of the last search failed the containment proof and were recomputed exactly."""
        return 0 if getattr(self, "_flags", None) is None else int((self._flags != 0).sum())

    def _merge_into(self, w):
        from ._ffi import check, lib, stream_ptr
        R = w["R"]
        check(lib.ugn_knn_merge_vote_packed(self.ctx.h, R["gathered"].ptr, w["nq"], self.k, R["od2"].ptr, R["oidx"].ptr,
                                            R["olab"].ptr, R["pred"].ptr, stream_ptr()))

    def _search(self, Q):
        """Search of this rank's shard -> ONE all-gather of the packed (d2, idx, label) lists -> merge + vote.
        The kernels of the two halves replay as CUDA graphs after the first call of a query count (use_graph)."""
        nq = int(Q.shape[0])
        w = self._work(nq)
        if isinstance(Q, torch.Tensor) and Q.is_cuda and Q.dtype == torch.float32:
            w["q"].copy_(Q, non_blocking=True)
        else:
            w["q"].copy_(_as_cuda(Q, torch.float32, self.dev), non_blocking=True)
        self._flags = w["flags"]
        world = w["world"]
        if self.use_graph and w["graph"] is None and w.get("warm", 0) >= 1:
            g1 = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g1, stream=side):
                    self._local_topk_into(w)
                    if world == 1:
                        self._merge_into(w)
                g2 = None
                if world > 1:
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2, stream=side):
                        self._merge_into(w)
            torch.cuda.current_stream().wait_stream(side)
            w["graph"] = (g1, g2)
        if w["graph"] is not None:
            w["graph"][0].replay()
            if world > 1:
                torch.distributed.all_gather_into_tensor(w["gathered"].view(-1), w["pack"], group=self.pg)
                w["graph"][1].replay()
        else:
            self._local_topk_into(w)
            if world > 1:
                torch.distributed.all_gather_into_tensor(w["gathered"].view(-1), w["pack"], group=self.pg)
            self._merge_into(w)
            w["warm"] = w.get("warm", 0) + 1
        return w["od2"], w["oidx"], w["olab"], w["pred"]

    def predict_device(self, Q) -> torch.Tensor:
        """Predicted labels i32 [Q] on the device (a view of the search workspace: valid until the next search).
        Device-side entry: shapes are validated, the values are not (no host synchronisation)."""
        self._validate(Q)
        return self._search(Q)[3]

    def _validate(self, Q):
        """scikit-learn's exceptions and wording for the same mistakes (KNeighborsClassifier.predict / kneighbors)."""
        shp = tuple(np.shape(Q))
        if len(shp) != 2:
            raise ValueError(f"Expected 2D array, got {len(shp)}D array instead")
        if shp[0] == 0:
            raise ValueError(f"Found array with 0 sample(s) (shape={shp}) while a minimum of 1 is required by "
                             "KNeighborsClassifier.")
        if shp[1] != self.G.shape[1]:
            raise ValueError(f"X has {shp[1]} features, but KNeighborsClassifier is expecting {self.G.shape[1]} features "
                             "as input.")
        if self.k > self.n_total:
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {self.k}, n_samples_fit = "
                             f"{self.n_total}, n_samples = {shp[0]}")

    def _finite_or_raise(self, nq):
        q = self._works[nq]["q"]
        if not bool(torch.isfinite(q).all()):
            raise ValueError("Input X contains NaN." if bool(torch.isnan(q).any()) else
                             "Input X contains infinity or a value too large for dtype('float32').")

    def predict(self, Q) -> np.ndarray:
        pred = self.predict_device(Q)
        self._finite_or_raise(int(np.shape(Q)[0]))
        return pred.cpu().numpy()

    def kneighbors_exact(self, Q):
        """(squared fp64 distances [Q,k], global indices [Q,k]) ordered by (distance, index)."""
        self._validate(Q)
        d2, idx, _, _ = self._search(Q)
        self._finite_or_raise(int(np.shape(Q)[0]))
        return d2.cpu().numpy().copy(), idx.cpu().numpy().copy()

    def kneighbors(self, Q, return_distance=True):
        d2, idx = self.kneighbors_exact(Q)
        return (np.sqrt(d2), idx) if return_distance else idx


def merge_vote(ctx, D2, IX, LB, k):
    nq = D2.shape[1]
    dev = D2.device
    od2 = torch.empty(nq, k, dtype=torch.float64, device=dev)
    oidx = torch.empty(nq, k, dtype=torch.int64, device=dev)
    olab = torch.empty(nq, k, dtype=torch.int32, device=dev)
    pred = torch.empty(nq, dtype=torch.int32, device=dev)
    ops.knn_merge_vote(ctx, D2, IX, LB, k, od2, oidx, olab, pred)
    return od2, oidx, olab, pred


def knn_sharded_local(G, y, Q, k, shards: int):
    """Single-GPU emulation of the sharded search (all shards searched by this process, then the
    same merge kernel): used to test the N>1 data path without N GPUs."""
    parts = []
    for r in range(shards):
        lo, hi = shard_bounds(len(G), r, shards)
        clf = KNeighborsClassifier(n_neighbors=k).fit(G[lo:hi], y[lo:hi], idx_base=lo, sharded=True)
        parts.append(tuple(t.clone() for t in clf._local_topk(_as_cuda(Q, torch.float32, clf.dev))))
        ctx = clf.ctx
    D2 = torch.stack([p[0] for p in parts]).contiguous()
    IX = torch.stack([p[1] for p in parts]).contiguous()
    LB = torch.stack([p[2] for p in parts]).contiguous()
    _, oidx, _, pred = merge_vote(ctx, D2, IX, LB, k)
    return pred.cpu().numpy(), oidx.cpu().numpy()
