"""Development aid: one rank's share of the 8-way sharded D = 2048 search on a single GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ugaitnet_b200.knn import KNeighborsClassifier
N, D, k = int(sys.argv[1]) if len(sys.argv) > 1 else 125000, int(sys.argv[2]) if len(sys.argv) > 2 else 2048, 3
g = torch.Generator(device="cuda").manual_seed(5)
cent = torch.randn(155, D, device="cuda", generator=g)
lab = torch.randint(0, 155, (N,), device="cuda", generator=g, dtype=torch.int32)
G = cent[lab.long()] + 0.35 * torch.randn(N, D, device="cuda", generator=g)
G = G / G.norm(dim=1, keepdim=True)
Q = cent[torch.randint(0, 155, (4096,), device="cuda", generator=g)] + 0.3 * torch.randn(4096, D, device="cuda", generator=g)
Q = Q / Q.norm(dim=1, keepdim=True)
clf = KNeighborsClassifier(n_neighbors=k).fit(G, lab)
for nq in (4096, 64):
    q = Q[:nq].contiguous()
    for _ in range(3):
        clf.predict_device(q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        clf.predict_device(q)
    torch.cuda.synchronize()
    print(f"N={N} D={D} Q={nq}: {(time.perf_counter() - t0) / 3 * 1e3:.3f} ms, flagged {clf.flagged_queries()}")
