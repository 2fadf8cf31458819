"""Drop-in for nets/mj_metrics.py (CPU metric, kept on the host as in the reference)."""
import numpy as np


def _roc(y, score):
    order = np.argsort(-score, kind="mergesort")
    score, y = score[order], y[order]
    idx = np.r_[np.where(np.diff(score))[0], y.size - 1]
    tps = np.cumsum(y)[idx]
    fps = 1 + idx - tps
    if tps.size > 2:     # sklearn.metrics.roc_curve(drop_intermediate=True)
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        tps, fps, idx = tps[keep], fps[keep], idx[keep]
    return np.r_[0, fps] / fps[-1], np.r_[0, tps] / tps[-1], np.r_[np.inf, score[idx]]


def mj_eerVerifDist(gt_labels, distances):
    """Equal error rate of a verification experiment: (EER, threshold); lower distance <-> label 1."""
    fpr, tpr, thr = _roc(np.asarray(gt_labels, dtype=np.float64), -np.asarray(distances, dtype=np.float64))
    i = np.nanargmin(np.absolute((1 - tpr) - fpr))
    return fpr[i], -thr[i]
