"""Development aid: every shard of the 8-way sharded D = 2048 bench gallery searched on ONE GPU (time + flagged)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ugaitnet_b200.knn import KNeighborsClassifier, shard_bounds
N, D, k, world = 1_000_000, 2048, 3, 8
for rank in range(world):
    lo, hi = shard_bounds(N, rank, world)
    g = torch.Generator(device="cuda").manual_seed(5)
    cent = torch.randn(155, D, device="cuda", generator=g)
    lab_all = torch.randint(0, 155, (N,), device="cuda", generator=g, dtype=torch.int32)
    G = torch.empty(hi - lo, D, device="cuda")
    step = 25_000
    for s in range(0, N, step):
        blk = cent[lab_all[s:s + step].long()] + 0.35 * torch.randn(step, D, device="cuda", generator=g)
        blk = blk / blk.norm(dim=1, keepdim=True)
        a, b = max(s, lo), min(s + step, hi)
        if a < b:
            G[a - lo:b - lo] = blk[a - s:b - s]
    dup = torch.arange(0, hi - lo - 1, 1000, device="cuda")
    G[dup + 1] = G[dup]
    lab = lab_all[lo:hi].contiguous()
    Qall = torch.randn(4096, D, device="cuda", generator=g) * 0.05
    Qall += cent[torch.randint(0, 155, (4096,), device="cuda", generator=g)] + 0.3 * torch.randn(4096, D, device="cuda", generator=g)
    Qall = Qall / Qall.norm(dim=1, keepdim=True)
    clf = KNeighborsClassifier(n_neighbors=k).fit(G, lab, idx_base=lo, sharded=True)
    for _ in range(3):
        clf.predict_device(Qall)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    clf.predict_device(Qall)
    torch.cuda.synchronize()
    print(f"shard {rank}: {(time.perf_counter() - t0) * 1e3:.2f} ms flagged {clf.flagged_queries()}", flush=True)
    del clf, G
    torch.cuda.empty_cache()
