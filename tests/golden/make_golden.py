"""Generates the committed golden vectors.  Run in the build container:

    python tests/golden/make_golden.py

* knn_*.npz    : outputs of the REAL reference call
                 sklearn.neighbors.KNeighborsClassifier(n_neighbors=k).fit(G, y).predict(Q)
                 (/root/reference/mains/mj_testUWYHGaitNet_open_tum.py:331-341) and .kneighbors(Q).
* eer.npz      : outputs of the reference's own nets/mj_metrics.py:mj_eerVerifDist imported from
                 /root/reference (including its demo known-answer EER 0.25 / thr 0.07).
* triplet.npz  : the literal op-by-op fp64 restatement of nets/triplet_loss_all.py:33-61 (TensorFlow is
                 not installable here: "parity unpinned" for this file) on balanced batches with
                 near-margin and duplicate rows.
* step_*.npz   : fp64 oracle (oracle/ugait_oracle.py, oracle/gaitset_oracle.py) losses, descriptors and gradient
                 norms of one training step of a tiny stacked-CNN and a tiny GaitSet model ("parity unpinned":
                 TensorFlow is not installable; the fixtures pin the restatement against drift and give the GPU
                 tests a committed target).   python tests/golden/make_golden.py steps   regenerates only these.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def knn_case(seed, N, D, Q, k, ncls, dup):
    from sklearn.neighbors import KNeighborsClassifier
    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(ncls, D))
    lab = rng.integers(0, ncls, N)
    G = cent[lab] + 0.35 * rng.normal(size=(N, D))
    G = (G / np.linalg.norm(G, axis=1, keepdims=True)).astype(np.float32)
    if dup:
        src = rng.integers(0, N, dup)
        dst = rng.integers(0, N, dup)
        G[dst] = G[src]          # exact duplicate rows (labels may differ -> tie cases)
    ql = rng.integers(0, ncls, Q)
    Qm = cent[ql] + 0.35 * rng.normal(size=(Q, D))
    Qm = (Qm / np.linalg.norm(Qm, axis=1, keepdims=True)).astype(np.float32)
    if dup:
        Qm[: min(Q, 8)] = G[dst[: min(Q, 8)]]   # queries that coincide with duplicated gallery rows
    clf = KNeighborsClassifier(n_neighbors=k).fit(G, lab)
    pred = clf.predict(Qm)
    dist, idx = clf.kneighbors(Qm)
    return dict(G=G, y=lab.astype(np.int32), Q=Qm, k=k, pred=pred.astype(np.int32), idx=idx.astype(np.int64),
                dist=dist)


STEP_CASES = {
    # name: (kind, config kwargs, batch kwargs, seed)
    "step_stacked": ("stacked", dict(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nc=8, nclasses=10, merge=2,
                                     act=2, wver=1.0, wid=0.3, label_smoothing=0.1, normbfmerge=True, aux_losses=True,
                                     waux=0.3), dict(base_rows=6, expand=4, kinds=("of", "gray", "depth")), 21),
    # (seed: no max-pool / set-max / LeakyReLU-sign decision of this tiny net within fp32 rounding of its boundary)
    "step_gaitset": ("gaitset", dict(in_channels=(2, 1), frames=4, hw=12, nc=16, nclasses=10, merge=0, wver=1.0, wid=1.0,
                                     label_smoothing=0.1), dict(ids=3, per_id=2), 9),
    # the two GaitSet graphs added last: ONE modality (branch output = signature) and postriplet == 2 (CPU pin of the
    # restatement; the engine is compared with the live oracle in tests/test_gaitset_gpu.py)
    "step_gaitset_single": ("gaitset", dict(in_channels=(1,), frames=3, hw=12, nc=0, nclasses=12, wver=1.0, wid=0.1,
                                            single=True), dict(ids=3, per_id=2), 7),
    "step_gaitset_post2": ("gaitset", dict(in_channels=(2, 1), frames=3, hw=12, nc=16, nclasses=10, merge=2, wver=1.0,
                                           wid=0.5, postriplet=2), dict(ids=3, per_id=2), 9),
}


def step_case(name):
    """One fp64 oracle step; returns the fixture dict (everything needed to rebuild the inputs is the seed)."""
    import torch
    from oracle import gaitset_oracle as G
    from oracle import ugait_oracle as O
    kind, ckw, bkw, seed = STEP_CASES[name]
    if kind == "stacked":
        oc = O.NetConfig(**ckw)
        xs, fl, lab = O.synth_batch(oc, seed=seed, **bkw)
        lab = lab % oc.nclasses
        P = O.init_params(oc, seed=seed, dtype=torch.float64)
        g = torch.Generator().manual_seed(seed)
        for k in P:
            if k.endswith("/b"):
                P[k] = torch.randn(P[k].shape, generator=g, dtype=torch.float64) * 0.05
        res, grads = O.loss_and_grads([torch.tensor(x, dtype=torch.float64) for x in xs],
                                      [torch.tensor(f, dtype=torch.float64) for f in fl], torch.tensor(lab), P, oc)
        sig = res["signature"]
    else:
        oc = G.GaitSetConfig(**ckw)
        xs, fl, lab = G.synth_batch(oc, seed=seed, dtype=torch.float64, **bkw)
        P = G.init_params(oc, seed=seed, dtype=torch.float64)
        res, grads = G.loss_and_grads(xs, fl, lab, P, oc)
        sig = res["signature"][[0, 1, 2, 30, 61]]                 # five of the 62 parts
    out = dict(triplet=float(res["triplet"]), ce=float(res["ce"]), count=float(res["count"].sum()), reg=float(res["reg"]),
               loss=float(res["loss"]), signature=sig.numpy(), names=np.array(sorted(grads)),
               grad_norms=np.array([float(grads[k].norm()) for k in sorted(grads)]))
    if "aux_ce" in res:
        out["aux_ce"] = np.array([float(v) for v in res["aux_ce"]])
    return out


def steps():
    for name in STEP_CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **step_case(name))
    print("step fixtures written to", HERE)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "steps":
        return steps()
    import sklearn
    for name, args in {
        "knn_small": (1, 2000, 64, 96, 3, 20, 0),
        "knn_dups": (2, 3000, 32, 64, 3, 12, 40),
        "knn_k7": (3, 2500, 48, 70, 7, 15, 10),
    }.items():
        c = knn_case(*args)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sklearn_version=sklearn.__version__, **c)

    sys.path.insert(0, "/root/reference")
    from nets.mj_metrics import mj_eerVerifDist
    y0 = np.array([1, 1, 1, 1, 1, 0, 0, 0, 0])
    d0 = np.array([0.01, 0.02, 0.015, 0.08, 0.05, 0.07, 0.2, 0.15, 0.18])
    cases_y, cases_d, outs = [y0], [d0], [mj_eerVerifDist(y0, d0)]
    for s in range(12):
        r = np.random.default_rng(100 + s)
        y = (r.random(150) < 0.4).astype(int)
        d = r.random(150) + 0.35 * (1 - y)
        cases_y.append(y); cases_d.append(d); outs.append(mj_eerVerifDist(y, d))
    np.savez_compressed(os.path.join(HERE, "eer.npz"), n=len(outs),
                        **{f"y{i}": v for i, v in enumerate(cases_y)},
                        **{f"d{i}": v for i, v in enumerate(cases_d)},
                        eer=np.array([o[0] for o in outs]), thr=np.array([o[1] for o in outs]))

    from oracle import ugait_oracle as O
    trip = {}
    rng = np.random.default_rng(7)
    for ci, (ids, per, d, n, margin) in enumerate([(6, 4, 16, 1, 0.2), (12, 8, 32, 1, 0.2), (4, 6, 8, 3, 0.5),
                                                    (5, 2, 24, 1, 0.05)]):
        B = ids * per
        lab = np.repeat(np.arange(ids), per).astype(np.float32)
        e = rng.normal(size=(n, B, d))
        e /= np.linalg.norm(e, axis=2, keepdims=True)
        e[:, 1] = e[:, 0]                       # exact duplicate of a same-label row
        e[:, per] = e[:, 0] + 1e-3 * rng.normal(size=(n, d))   # near-duplicate with a different label
        e = e.astype(np.float32)
        loss, cnt = O.triplet_loss_all_literal_np(lab, e, margin)
        trip[f"lab{ci}"], trip[f"emb{ci}"], trip[f"margin{ci}"] = lab, e, margin
        trip[f"loss{ci}"], trip[f"cnt{ci}"] = loss, cnt
    np.savez_compressed(os.path.join(HERE, "triplet.npz"), n=4, **trip)
    steps()
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
