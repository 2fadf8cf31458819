"""Host-side multi-GPU plumbing (torch.distributed; NCCL on the box, gloo in the CPU tests).

* Training is batch data-parallel with the reference's MirroredStrategy semantics
  (/root/reference/mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:342-349): every rank mines triplets on
  its own rows, gradients are summed by one all-reduce over the flat gradient arena and scaled by 1/G
  inside the fused optimiser kernel (ugn_adam_step gscale).
  Default on NVLink boxes: the exchange is fused into the optimiser (ugn_dp_optim_step): rank r owns the r-th
  slice of the arena (owner_slice), sums the ranks' gradients for it through peer memory, updates it with its slice
  of the optimiser state and stores the new weights into every rank's arena.
* The k-NN gallery is row-sharded; the only exchange is an all-gather of the per-rank
  (dist2, global idx, label)[Q,k] lists, merged with the (distance, index) order.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_slice(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Element range of the flat arena (length n, a multiple of 4) that `rank` owns in ugn_dp_optim_step: equal
    slices of ceil(n/4 / world) float4 vectors, the last one(s) possibly shorter or empty (the twin of the slice
    arithmetic in csrc/elementwise.cu: ew_optim)."""
    n4 = n // 4
    per = (n4 + world - 1) // world
    return 4 * min(n4, per * rank), 4 * min(n4, per * (rank + 1))


def owned_segment_ranges(segments, rank: int, world: int, n_arena: int):
    """For the deferred weight all-gather: the parts of the given arena segments ``(name, offset, numel)`` that lie inside
    the slice `rank` owns -> [(name, lo, hi)] with lo / hi relative to the segment start.  Over all ranks every element
    of every segment is covered exactly once."""
    q0, q1 = owner_slice(n_arena, rank, world)
    out = []
    for name, off, n in segments:
        a, b = max(off, q0), min(off + n, q1)
        if a < b:
            out.append((name, a - off, b - off))
    return out


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks of a flat gradient arena (CPU/gloo test helper; on the GPU the 1/G
    factor is folded into the optimiser kernel instead of a separate pass)."""
    dist.all_reduce(flat, group=group)
    flat.div_(dist.get_world_size(group))
    return flat


def allgather_topk(d2: torch.Tensor, idx: torch.Tensor, lab: torch.Tensor, group=None):
    world = dist.get_world_size(group)
    outs = []
    for t in (d2, idx, lab):
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t.contiguous(), group=group)
        outs.append(torch.stack(buf))
    return outs


def merge_topk_host(D2: np.ndarray, IX: np.ndarray, LB: np.ndarray, k: int):
    """numpy twin of ugn_knn_merge_vote for the CPU tests: [G,Q,k] lists -> merged [Q,k] + vote."""
    G, Q, _ = D2.shape
    d = D2.transpose(1, 0, 2).reshape(Q, G * k)
    i = IX.transpose(1, 0, 2).reshape(Q, G * k)
    l = LB.transpose(1, 0, 2).reshape(Q, G * k)
    order = np.lexsort((i, d), axis=1)[:, :k]
    md, mi, ml = (np.take_along_axis(a, order, 1) for a in (d, i, l))
    pred = np.empty(Q, dtype=ml.dtype)
    for q in range(Q):
        vals, cnts = np.unique(ml[q], return_counts=True)
        pred[q] = vals[np.argmax(cnts)]
    return md, mi, ml, pred
