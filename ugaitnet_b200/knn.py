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

    # ---- fit: keep (this rank's shard of) the gallery resident in HBM ------------------------
    def fit(self, X, y, idx_base: int = 0, sharded: bool = False):
        """X [N,D] float, y [N] int labels.  With a process group and ``sharded=False`` the full
        gallery is given on every rank and this rank keeps rows shard_bounds(N, rank, world);
        with ``sharded=True`` X/y already are this rank's shard and idx_base its global offset."""
        if self.pg is not None and not sharded:
            rank, world = torch.distributed.get_rank(self.pg), torch.distributed.get_world_size(self.pg)
            lo, hi = shard_bounds(len(X), rank, world)
            X, y, idx_base = X[lo:hi], y[lo:hi], lo
        self.G = _as_cuda(X, torch.float32, self.dev)
        self.labels = _as_cuda(y, torch.int32, self.dev)
        self.idx_base = int(idx_base)
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
        return self

    # ---- local shard search -------------------------------------------------------------------
    def _local_topk(self, Q: torch.Tensor):
        nq, k = Q.shape[0], self.k
        d2 = torch.empty(nq, k, dtype=torch.float64, device=self.dev)
        idx = torch.empty(nq, k, dtype=torch.int64, device=self.dev)
        lab = torch.empty(nq, k, dtype=torch.int32, device=self.dev)
        flags = torch.zeros(nq, dtype=torch.int32, device=self.dev) if self.use_tc else None
        for s in range(0, nq, self.query_block):
            e = min(nq, s + self.query_block)
            ws = torch.empty(max(ops.knn_workspace_bytes(e - s, self.G.shape[0], self.G.shape[1], k) // 4, 4),
                             dtype=torch.float32, device=self.dev)
            if self.use_tc:
                q16 = torch.empty(2, self.kch, e - s, 64, dtype=torch.float16, device=self.dev)
                ops.knn_pack(self.ctx, Q[s:e], q16)
                ops.knn_topk_tc(self.ctx, Q[s:e], q16, self.G, self.G16, self.g2, self.gmax2, self.labels, k,
                                self.idx_base, d2[s:e], idx[s:e], lab[s:e], flags[s:e], ws)
            else:
                ops.knn_topk(self.ctx, Q[s:e], self.G, self.g2, self.labels, k, self.idx_base, d2[s:e], idx[s:e],
                             lab[s:e], ws)
        self._flags = flags
        return d2, idx, lab

    def flagged_queries(self) -> int:
        """How many queries of the last search failed the containment proof and were recomputed exactly."""
        return 0 if getattr(self, "_flags", None) is None else int(self._flags.sum())

    def _search(self, Q):
        Q = _as_cuda(Q, torch.float32, self.dev)
        d2, idx, lab = self._local_topk(Q)
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            world = torch.distributed.get_world_size(self.pg)
            D2 = torch.empty((world,) + d2.shape, dtype=d2.dtype, device=self.dev)
            IX = torch.empty((world,) + idx.shape, dtype=idx.dtype, device=self.dev)
            LB = torch.empty((world,) + lab.shape, dtype=lab.dtype, device=self.dev)
            torch.distributed.all_gather_into_tensor(D2, d2, group=self.pg)
            torch.distributed.all_gather_into_tensor(IX, idx, group=self.pg)
            torch.distributed.all_gather_into_tensor(LB, lab, group=self.pg)
        else:
            D2, IX, LB = d2.unsqueeze(0), idx.unsqueeze(0), lab.unsqueeze(0)
        return merge_vote(self.ctx, D2.contiguous(), IX.contiguous(), LB.contiguous(), self.k)

    def predict_device(self, Q) -> torch.Tensor:
        return self._search(Q)[3]

    def predict(self, Q) -> np.ndarray:
        return self.predict_device(Q).cpu().numpy()

    def kneighbors_exact(self, Q):
        """(squared fp64 distances [Q,k], global indices [Q,k]) ordered by (distance, index)."""
        d2, idx, _, _ = self._search(Q)
        return d2.cpu().numpy(), idx.cpu().numpy()

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
        parts.append(clf._local_topk(_as_cuda(Q, torch.float32, clf.dev)))
        ctx = clf.ctx
    D2 = torch.stack([p[0] for p in parts]).contiguous()
    IX = torch.stack([p[1] for p in parts]).contiguous()
    LB = torch.stack([p[2] for p in parts]).contiguous()
    _, oidx, _, pred = merge_vote(ctx, D2, IX, LB, k)
    return pred.cpu().numpy(), oidx.cpu().numpy()
