"""Video-level stage of the open-world test.

Mirrors /root/reference/mains/mj_testUWYHGaitNet_open_tum.py:355-461: sub-sequence descriptors are pooled
per video (mean or max), sub-sequence labels / predictions are voted per video with `statistics.mode`,
and a second k-NN runs on the pooled descriptors.  The grouping table (np.unique / np.where in the
reference) is index bookkeeping and stays on the host; pooling, votes and both k-NN searches run on the GPU.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import ops
from ._ffi import TRef, check, lib, stream_ptr
from .knn import KNeighborsClassifier


def video_groups(vids):
    """(unique video ids, order i32 [N], offsets i32 [V+1]) -- rows of video uvids[v] are
    order[offsets[v]:offsets[v+1]] in their original order (== np.where(vids == vix)[0])."""
    vids = np.asarray(vids).reshape(-1)
    uvids, inv, counts = np.unique(vids, return_inverse=True, return_counts=True)
    order = np.argsort(inv, kind="stable").astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return uvids, order, offsets


def _dev(a, dtype, dev):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(dev)


def pool_per_video(ctx, codes, order, offsets, use_avg=True) -> torch.Tensor:
    dev = torch.device("cuda", ctx.device)
    c, o, f = _dev(codes, torch.float32, dev), _dev(order, torch.int32, dev), _dev(offsets, torch.int32, dev)
    out = torch.empty(len(offsets) - 1, c.shape[1], device=dev)
    rs = [TRef(t) for t in (c, o, f, out)]
    check(lib.ugn_segment_pool(ctx.h, rs[0].ptr, rs[1].ptr, rs[2].ptr, int(bool(use_avg)), rs[3].ptr, stream_ptr()))
    return out


def mode_per_video(ctx, labels, order, offsets, legacy_ties=False) -> torch.Tensor:
    dev = torch.device("cuda", ctx.device)
    l, o, f = _dev(labels, torch.int32, dev), _dev(order, torch.int32, dev), _dev(offsets, torch.int32, dev)
    out = torch.empty(len(offsets) - 1, dtype=torch.int32, device=dev)
    rs = [TRef(t) for t in (l, o, f, out)]
    check(lib.ugn_segment_mode(ctx.h, rs[0].ptr, rs[1].ptr, rs[2].ptr, int(bool(legacy_ties)), rs[3].ptr, stream_ptr()))
    return out


def evaluate_open_world(codes_gallery, labs_gallery, vids_gallery, codes_test, labs_test, vids_test, knn=3,
                        use_avg=True, legacy_ties=False) -> Dict[str, object]:
    """evalUWYHNet's classification stage (:331-461) on the GPU.  Returns sub-sequence predictions and
    accuracy, per-video voted predictions / accuracy and the accuracy of the k-NN on pooled descriptors
    (`summary = (acc, acc_vid, score)` of the reference)."""
    ctx = ops.get_ctx()
    clf = KNeighborsClassifier(n_neighbors=knn).fit(codes_gallery, np.asarray(labs_gallery).reshape(-1))
    pred = clf.predict_device(codes_test)
    labs_test = np.asarray(labs_test).reshape(-1).astype(np.int32)
    acc = float((pred.cpu().numpy() == labs_test).mean())
    _, og, fg = video_groups(vids_gallery)
    uv, ot, ft = video_groups(vids_test)
    codes_vid_g = pool_per_video(ctx, codes_gallery, og, fg, use_avg)
    labs_vid_g = mode_per_video(ctx, np.asarray(labs_gallery).reshape(-1), og, fg, legacy_ties)
    codes_vid_t = pool_per_video(ctx, codes_test, ot, ft, use_avg)
    labs_vid_t = mode_per_video(ctx, labs_test, ot, ft, legacy_ties)
    pred_vid = mode_per_video(ctx, pred, ot, ft, legacy_ties)
    acc_vid = float((pred_vid == labs_vid_t).float().mean())
    k2 = min(knn, codes_vid_g.shape[0])
    clf2 = KNeighborsClassifier(n_neighbors=k2).fit(codes_vid_g, labs_vid_g)
    pred_merged = clf2.predict_device(codes_vid_t)
    score = float((pred_merged == labs_vid_t).float().mean())
    return {"pred": pred, "acc": acc, "video_ids": uv, "pred_vid": pred_vid, "labs_vid": labs_vid_t, "acc_vid": acc_vid,
            "pred_vid_merged": pred_merged, "score": score, "codes_vid_test": codes_vid_t,
            "codes_vid_gallery": codes_vid_g, "labs_vid_gallery": labs_vid_g, "summary": (acc, acc_vid, score)}
