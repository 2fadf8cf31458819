"""Development aid: per-tensor errors of the GaitSet step against the oracle (run on the GPU box)."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_gaitset_gpu as T
from oracle import gaitset_oracle as G

import itertools
for (name, mode), seed in itertools.product([("small_2mod_code", "fp32"), ("small_3mod_signmax", "f16mix")], [7, 8, 9]):
    oc, eng, P, xs, fl, lab = T.setup(name, math_mode=mode, seed=seed)
    res, grads = G.loss_and_grads(xs, fl, lab, P, oc)
    out = eng.loss_and_grad(T.cu(xs), T.cu(fl), lab.cuda())
    eng.ctx.check()
    print("==", name, mode, seed, "trip", float(out["triplet"]), float(res["triplet"]), "sig", T.rel(out["signature"], res["signature"]))
    got = eng.export_grads()
    for k, g in grads.items():
        print(f"  {k:28s} {T.rel(got[k], g):.3e}  |g| {float(g.abs().max()):.3e}")
    continue
    p = eng.plan(xs[0].shape[0], True)
    for m, b in enumerate(p.br):
        for k, t in b.T.items():
            if k.startswith("dz_") or k.startswith("d_"):
                v = t.float().abs()
                nz = v[v > 0]
                print(f"    br{m} {k:8s} max {float(v.max()):.3e} min-nz {float(nz.min()) if nz.numel() else 0:.3e} "
                      f"frac<6e-5 {float((nz < 6e-5).float().mean()) if nz.numel() else 0:.3f}")
