"""Drop-in for nets/triplet_loss_all.py: `triplet_loss(margin) -> loss(y_true, y_pred)`.

y_true: labels [m,1] (or [m]); y_pred: embeddings [n_parts, m, d] (GaitSet layout, the only rank the
reference body accepts, :33-36) or [m, d] (stacked-frame CNN signature == n_parts 1).  The value is
computed by the CUDA batch-all kernel (ugn_triplet_all)."""
from __future__ import annotations

import numpy as np
import torch

from ugaitnet_b200 import ops


class _TripletLoss:
    __name__ = "loss"

    def __init__(self, margin=1.0):
        self.margin = margin           # mutable, as in `losses[0].margin = margin` (mj_uwyhNets_ba.py:1293)

    def __call__(self, y_true, y_pred):
        dev = torch.device("cuda", torch.cuda.current_device())
        emb = torch.as_tensor(np.asarray(y_pred) if not torch.is_tensor(y_pred) else y_pred,
                              dtype=torch.float32).to(dev).contiguous()
        lab = torch.as_tensor(np.asarray(y_true) if not torch.is_tensor(y_true) else y_true).to(dev)
        lab = lab.reshape(-1).to(torch.int32).contiguous()
        n, B = (emb.shape[0], emb.shape[1]) if emb.dim() == 3 else (1, emb.shape[0])
        ctx = ops.get_ctx(dev.index)
        out = torch.zeros(2, device=dev)
        ws = torch.zeros(ops.triplet_workspace_bytes(n, B) // 4 + 8, device=dev)
        ops.triplet_all(ctx, emb, lab, float(self.margin), 1.0, out, None, ws)
        return out[0]


def triplet_loss(margin=1.0):
    return _TripletLoss(margin)


def batch_dist(x):
    """nets/triplet_loss_all.py:70-77 on a [n,m,d] tensor.  A stand-alone helper of the reference that its loss calls
    internally; here the loss (above) computes its distances inside ugn_triplet_all{,_tc}, and this function only
    restates the same formula with torch operators for callers that want the matrix itself (not on the step's path)."""
    x = torch.as_tensor(x)
    x2 = (x * x).sum(2)
    d = x2.unsqueeze(2) + x2.unsqueeze(1) - 2.0 * torch.matmul(x, x.transpose(1, 2))
    d = torch.clamp(d, min=0.0)
    err = d <= 0.0
    return torch.sqrt(d + err.to(d.dtype) * 1e-16) * (~err).to(d.dtype)
