"""Probe edge cases on the GPU: tiny batches, tiny galleries, k > N, empty query sets."""
import sys, traceback
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import ugait_oracle as O
from ugaitnet_b200.net import UGaitEngine
from ugaitnet_b200.config import NetConfig
from ugaitnet_b200.knn import KNeighborsClassifier

def attempt(name, fn):
    try:
        print(name, "->", fn())
    except Exception as e:
        print(name, "RAISED", type(e).__name__, str(e)[:200])

rng = np.random.default_rng(0)
G = rng.normal(size=(50, 24)).astype(np.float32); y = rng.integers(0, 5, 50)
attempt("knn N=50 Q=1 k=3", lambda: KNeighborsClassifier(n_neighbors=3).fit(G, y).predict(G[:1] + 0.01))
attempt("knn N=2 k=3", lambda: KNeighborsClassifier(n_neighbors=3).fit(G[:2], y[:2]).predict(G[:4]))
attempt("knn N=3 k=3", lambda: KNeighborsClassifier(n_neighbors=3).fit(G[:3], y[:3]).predict(G[:4]))
attempt("knn Q=0", lambda: KNeighborsClassifier(n_neighbors=3).fit(G, y).predict(G[:0]))
attempt("knn D mismatch", lambda: KNeighborsClassifier(n_neighbors=3).fit(G, y).predict(G[:3, :10]))
attempt("knn NaN query", lambda: KNeighborsClassifier(n_neighbors=3).fit(G, y).predict(np.full((2, 24), np.nan, np.float32)))
from sklearn.neighbors import KNeighborsClassifier as SK
attempt("sk N=2 k=3", lambda: SK(n_neighbors=3).fit(G[:2], y[:2]).predict(G[:4]))
attempt("sk Q=0", lambda: SK(n_neighbors=3).fit(G, y).predict(G[:0]))
attempt("sk NaN", lambda: SK(n_neighbors=3).fit(G, y).predict(np.full((2, 24), np.nan, np.float32)))

for mode in ("fp32", "f16mix"):
    cfg = NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=64, nclasses=10, merge=2, wver=1.0, wid=0.1)
    eng = UGaitEngine(cfg, math_mode=mode, lr=1e-3)
    for B in (1, 2, 7):
        xs = [torch.randn(B, c, 60, 60, device="cuda") for c in cfg.in_channels]
        fl = [torch.ones(B, 1, device="cuda") for _ in cfg.in_channels]
        lab = torch.arange(B, device="cuda") % 3
        attempt(f"{mode} train B={B}", lambda: {k: (float(v) if torch.is_tensor(v) and v.numel() == 1 else None) for k, v in eng.train_step(xs, fl, lab).items() if k in ("triplet", "count", "ce")})
        attempt(f"{mode} predict B={B}", lambda: tuple(eng.predict(xs, fl).shape))
    attempt(f"{mode} all modalities missing", lambda: float(eng.predict([torch.randn(2, c, 60, 60, device="cuda") for c in cfg.in_channels], [torch.zeros(2, 1, device="cuda")] * 3).abs().max()))
    attempt(f"{mode} B=0", lambda: eng.predict([torch.randn(0, c, 60, 60, device="cuda") for c in cfg.in_channels], [torch.ones(0, 1, device="cuda")] * 3).shape)
    attempt(f"{mode} wrong hw", lambda: eng.predict([torch.randn(2, c, 50, 60, device="cuda") for c in cfg.in_channels], [torch.ones(2, 1, device="cuda")] * 3).shape)
    attempt(f"{mode} label out of range", lambda: float(eng.train_step([torch.randn(2, c, 60, 60, device="cuda") for c in cfg.in_channels], [torch.ones(2, 1, device="cuda")] * 3, torch.tensor([3, 99], device="cuda"))["ce"]))
    torch.cuda.synchronize()
