"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: gallery sharding + top-k list merge,
and the data-parallel gradient averaging contract (sum all-reduce, 1/G scale inside the optimiser)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ugait_oracle as O


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ugaitnet_b200.dist import allgather_topk, merge_topk_host, shard_bounds, allreduce_mean_
    rng = np.random.default_rng(0)                 # same data on every rank
    N, D, Q, k = 1001, 16, 37, 3
    G = rng.normal(size=(N, D)).astype(np.float32)
    G[500:510] = G[0:10]
    y = rng.integers(0, 9, N).astype(np.int32)
    Qm = rng.normal(size=(Q, D)).astype(np.float32)
    lo, hi = shard_bounds(N, rank, world)
    d2, idx = O.knn_search(G[lo:hi], Qm, k)        # the local search (GPU kernel on the box; oracle here)
    d2t, idxt = torch.from_numpy(d2), torch.from_numpy(idx + lo)
    labt = torch.from_numpy(y[lo:hi][idx])
    D2, IX, LB = allgather_topk(d2t, idxt, labt)
    md2, midx, mlab, pred = merge_topk_host(D2.numpy(), IX.numpy(), LB.numpy(), k)
    rd2, ridx = O.knn_search(G, Qm, k)
    ok_knn = np.array_equal(midx, ridx) and np.array_equal(pred, O.knn_vote(y[ridx]))
    # DP contract: every rank holds a different gradient; after the collective all hold the mean
    g = torch.full((130,), float(rank + 1))
    allreduce_mean_(g)
    ok_dp = torch.allclose(g, torch.full((130,), (world + 1) / 2.0))
    # fused exchange contract (ugn_dp_optim_step): reduce-scatter to the slice owner -> Adam on the slice with the
    # owner's slice of m / v -> all-gather of the weights  ==  all-reduce(mean) + full Adam on every rank
    from ugaitnet_b200.dist import owner_slice
    n = 4 * 37
    gen = torch.Generator().manual_seed(7)
    w0 = torch.randn(n, generator=gen, dtype=torch.float64)
    w_ref, m_ref, v_ref = {"w": w0.clone()}, {"w": torch.zeros(n, dtype=torch.float64)}, {"w": torch.zeros(n, dtype=torch.float64)}
    w, m, v = w0.clone(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    lo, hi = owner_slice(n, rank, world)
    for t in range(1, 4):
        g = torch.randn(n, generator=torch.Generator().manual_seed(100 * t + rank), dtype=torch.float64)
        allg = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(allg, g)                                   # the peers' gradient arenas
        gm = torch.stack(allg).mean(0)
        O.adam_step(w_ref, {"w": gm.clone()}, m_ref, v_ref, t, lr=1e-2)
        sl = {"w": w[lo:hi]}                                       # views: the owner touches only its slice of m, v
        O.adam_step(sl, {"w": gm[lo:hi].clone()}, {"w": m[lo:hi]}, {"w": v[lo:hi]}, t, lr=1e-2)
        pad = max(owner_slice(n, r, world)[1] - owner_slice(n, r, world)[0] for r in range(world))
        mine = torch.zeros(pad, dtype=torch.float64)
        mine[:hi - lo] = w[lo:hi]
        parts = [torch.empty(pad, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, mine)                               # the owners' stores into every rank's arena
        w = torch.cat([parts[r][:owner_slice(n, r, world)[1] - owner_slice(n, r, world)[0]] for r in range(world)])
    ok_fused = torch.allclose(w, w_ref["w"], rtol=0, atol=1e-12) and float(m[:lo].abs().sum() + m[hi:].abs().sum()) == 0.0
    with open(os.path.join(tmp, f"r{rank}"), "w") as f:
        f.write(f"{int(ok_knn)} {int(ok_dp)} {int(ok_fused)}")
    dist.destroy_process_group()


def test_sharded_knn_merge_and_dp_mean_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"r{r}").read() == "1 1 1"


def test_owner_slices_cover_the_arena():
    from ugaitnet_b200.dist import owner_slice
    for n in (0, 4, 8, 4 * 37, 4 * 1001, 89_600_000):
        for w in (2, 3, 4, 8):
            b = [owner_slice(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in b)
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_shard_bounds_cover_everything():
    from ugaitnet_b200.dist import shard_bounds
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_integer_shift_restatement_equals_scipy_affine_transform():
    """expand.shift_sequence == what ImageDataGenerator.apply_transform does for a pure integer displacement
    (keras_preprocessing apply_affine_transform: scipy.ndimage.affine_transform(channel, identity, offset=(tx, ty),
    order=1, mode='nearest'); data/mj_augmentation.py:35-50,62-64), and the decoded-value clip == __load_dd's raw clip."""
    import numpy as np
    from scipy import ndimage
    from ugaitnet_b200.expand import clip_flow, shift_sequence
    rng = np.random.default_rng(0)
    x = rng.normal(size=(6, 20, 20)).astype(np.float32)
    for tx, ty in [(-5, 3), (0, 0), (5, 5), (3, -3), (-3, 0)]:
        ref = np.stack([ndimage.affine_transform(x[i], np.eye(2), offset=(tx, ty), order=1, mode="nearest", cval=0.0)
                        for i in range(x.shape[0])])
        assert np.array_equal(shift_sequence(x, tx, ty), ref), (tx, ty)
    raw = rng.integers(-3000, 3000, size=(50, 8, 8)).astype(np.float32)
    lit = raw.copy()
    lit[np.abs(lit) > 2300] = 1e-8
    lit[np.abs(lit) < 50] = 1e-8
    lit = lit / 100 * 0.1
    assert np.allclose(clip_flow(raw / 100 * 0.1), lit, rtol=1e-6, atol=1e-12)


def test_owned_segment_ranges_partition_every_segment():
    """Deferred weight all-gather: the plane ranges each rank sends (its arena slice cut by the exchanged segments) tile
    every exchanged segment exactly once, for any world size, and agree with the slice arithmetic of ugn_dp_optim_step."""
    from ugaitnet_b200.dist import owned_segment_ranges, owner_slice
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 4, 8):
        off, segs = 0, []
        for i in range(9):
            n = int(rng.integers(1, 40)) * 4
            if i % 3 != 1:                       # some segments are not exchanged (conv weights, biases)
                segs.append((f"s{i}", off, n))
            off += (n + 63) // 64 * 64           # arena offsets are rounded up to 64 elements
        n_arena = off
        cover = {name: np.zeros(n, dtype=np.int32) for name, _, n in segs}
        for rank in range(world):
            q0, q1 = owner_slice(n_arena, rank, world)
            for name, lo, hi in owned_segment_ranges(segs, rank, world, n_arena):
                so = next(o for nm, o, _ in segs if nm == name)
                assert 0 <= lo < hi <= len(cover[name]) and q0 <= so + lo and so + hi <= q1
                cover[name][lo:hi] += 1
        assert all((c == 1).all() for c in cover.values())
