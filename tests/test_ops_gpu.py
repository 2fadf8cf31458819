"""GPU parity tests, operator by operator, through the C ABI, against the CPU oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ugait_oracle as O

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5   # north_star: <= 1e-5 in the fp32 validation mode


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ctx():
    from ugaitnet_b200 import ops
    assert torch.cuda.is_available()
    return ops.get_ctx(0)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("B,C,Cp,H,Co,k,pool,act", [
    (3, 25, 32, 20, 24, 7, True, 1), (2, 50, 64, 15, 40, 5, True, 2), (4, 32, 32, 9, 64, 3, True, 1),
    (5, 64, 64, 4, 96, 2, False, 1), (1, 3, 3, 12, 5, 3, True, 0), (2, 8, 8, 11, 70, 3, False, 2)])
def test_conv_layer_fp32(ctx, B, C, Cp, H, Co, k, pool, act):
    from ugaitnet_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(Co, C, k, k, generator=g) * 0.1
    b = torch.randn(Co, generator=g)
    xd = torch.zeros(B, H, H, Cp, device="cuda")
    ops.pack_input(ctx, x.cuda(), xd)
    assert torch.equal(xd[..., :C].cpu(), nhwc(x)) and (Cp == C or float(xd[..., C:].abs().max()) == 0)
    wm = w.permute(0, 2, 3, 1).contiguous().cuda()
    wp = torch.zeros(Co, k, k, Cp, device="cuda")
    ops.pack_weight(ctx, wm, wp)
    Ho = H - k + 1
    Hp = Ho // 2 if pool else Ho
    y = torch.zeros(B, Hp, Hp, Co, device="cuda")
    idx = torch.zeros(B, Hp, Hp, Co, dtype=torch.uint8, device="cuda") if pool else None
    ops.conv2d_fwd(ctx, xd, wp, b.cuda(), y, idx, act=act, alpha=0.3, pool=pool)
    # oracle (fp64)
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    z = F.conv2d(x64, w64, b64)
    a = O._act(z, act, 0.3)
    ref = F.max_pool2d(a, 2) if pool else a
    assert rel(y.permute(0, 3, 1, 2), ref.detach()) < FP32_TOL
    # backward
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())
    dz = torch.zeros(B, Ho, Ho, Co, device="cuda")
    db2 = torch.full((Co,), 3.0, device="cuda")
    ops.conv2d_bwd_act(ctx, nhwc(dy).cuda(), y, idx, dz, act=act, alpha=0.3, pool=pool, db=db2)
    assert rel(db2, b64.grad) < FP32_TOL       # bias gradient reduced in the same pass
    dw = torch.zeros(Co, k, k, C, device="cuda")
    db = torch.zeros(Co, device="cuda")
    ops.conv2d_wgrad(ctx, xd, dz, dw, db)
    assert rel(dw.permute(0, 3, 1, 2), w64.grad) < FP32_TOL
    assert rel(db, b64.grad) < FP32_TOL
    dx = torch.zeros(B, H, H, Cp, device="cuda")
    ops.conv2d_dgrad(ctx, dz, wp, dx)
    assert rel(dx[..., :C].permute(0, 3, 1, 2), x64.grad) < FP32_TOL
    if Cp > C:
        assert float(dx[..., C:].abs().max()) == 0.0


def test_pool_floor_rows_get_zero_gradient(ctx):
    from ugaitnet_b200 import ops
    B, C, Ho, Hp = 2, 8, 9, 4          # 9 -> 4 drops the last row/col (nets/mj_uwyhNets_ba.py:85)
    y = torch.rand(B, Hp, Hp, C, device="cuda") + 0.1
    idx = torch.randint(0, 4, (B, Hp, Hp, C), dtype=torch.uint8, device="cuda")
    dy = torch.randn(B, Hp, Hp, C, device="cuda")
    dz = torch.full((B, Ho, Ho, C), 7.0, device="cuda")
    ops.conv2d_bwd_act(ctx, dy, y, idx, dz, act=1, pool=True)
    assert float(dz[:, 8].abs().max()) == 0 and float(dz[:, :, 8].abs().max()) == 0
    assert torch.allclose(dz.sum((1, 2)), dy.sum((1, 2)), atol=1e-5)


def test_flatten_is_chw_order(ctx):
    from ugaitnet_b200 import ops
    y = torch.randn(3, 3, 3, 16, device="cuda")
    flat = torch.zeros(3, 144, device="cuda")
    ops.flatten_chw(ctx, y, flat)
    assert torch.equal(flat, y.permute(0, 3, 1, 2).reshape(3, -1))
    back = torch.zeros_like(y)
    ops.unflatten_chw(ctx, flat, back)
    assert torch.equal(back, y)


@pytest.mark.parametrize("B,K,N,act,mask", [(24, 300, 130, 0, False), (7, 64, 150, 1, True), (96, 515, 64, 2, True)])
def test_linear_fp32(ctx, B, K, N, act, mask):
    from ugaitnet_b200 import ops
    g = torch.Generator().manual_seed(K)
    x, w, b = torch.randn(B, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    m = ((torch.rand(B, N, generator=g) > 0.4).float() / 0.6) if mask else None
    y = torch.zeros(B, N, device="cuda")
    ops.linear_fwd(ctx, x.cuda(), w.cuda(), b.cuda(), m.cuda() if mask else None, y, None, act=act, alpha=0.3)
    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x, w, b))
    ref = O._act(F.linear(x64, w64, b64), act, 0.3)
    if mask:
        ref = ref * m.double()
    assert rel(y, ref.detach()) < FP32_TOL
    dy = torch.randn(B, N, generator=g)
    ref.backward(dy.double())
    dz = torch.zeros(B, N, device="cuda")
    ops.act_mask_bwd(ctx, dy.cuda(), y if act else None, m.cuda() if mask else None, dz, None, act=act, alpha=0.3)
    dx, dw, db = torch.zeros(B, K, device="cuda"), torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    ops.linear_bwd(ctx, x.cuda(), w.cuda(), dz, dx, dw, db)
    assert rel(dx, x64.grad) < FP32_TOL and rel(dw, w64.grad) < FP32_TOL and rel(db, b64.grad) < FP32_TOL


@pytest.mark.parametrize("merge", [0, 1, 2])
@pytest.mark.parametrize("nmods", [2, 3])
def test_fusion_fwd_bwd(ctx, merge, nmods):
    from ugaitnet_b200 import ops
    B, d = 13, 200
    g = torch.Generator().manual_seed(merge * 10 + nmods)
    br = [torch.randn(B, d, generator=g) for _ in range(nmods)]
    br[1][:, :20] = br[0][:, :20]                 # exact ties between modalities
    br[1][:, 20:30] = -br[0][:, 20:30]            # |x| ties for sign_max
    fl = [(torch.rand(B, 1, generator=g) > 0.3).float() for _ in range(nmods)]
    for f in fl:
        f[0] = 0.0                                # row 0: every modality missing -> eps path
    fl[0][1] = 1.0
    sig = torch.zeros(B, d, device="cuda")
    win = torch.zeros(B, d, dtype=torch.uint8, device="cuda")
    inv = torch.zeros(B, 2, device="cuda")
    ops.fuse_fwd(ctx, [t.cuda() for t in br], [f.cuda() for f in fl], sig, None, win, inv, merge, True)
    b64 = [t.double().requires_grad_(True) for t in br]
    ref = O.l2_normalize(O.merge_modalities([t * f.double() for t, f in zip(b64, fl)], merge), 1)
    assert torch.allclose(sig.cpu().double(), ref.detach(), atol=2e-6, rtol=1e-5)
    assert float(sig[0].abs().max()) == 0.0
    dsig = torch.randn(B, d, generator=g)
    ref.backward(dsig.double())
    dbr = [torch.zeros(B, d, device="cuda") for _ in range(nmods)]
    ops.fuse_bwd(ctx, dsig.cuda(), sig, win, inv, [f.cuda() for f in fl], dbr, merge, True)
    for m in range(nmods):
        assert rel(dbr[m][1:], b64[m].grad[1:]) < 2e-5, m


def test_softmax_ce(ctx):
    from ugaitnet_b200 import ops
    B, C = 37, 150
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(B, C, generator=g) * 3
    lab = torch.randint(0, C, (B,), generator=g)
    out = torch.zeros(2, device="cuda")
    dl = torch.zeros(B, C, device="cuda")
    ops.softmax_ce(ctx, logits.cuda(), lab.int().cuda(), out, dl, 0.1)
    l64 = logits.double().requires_grad_(True)
    loss, acc = O.softmax_ce(l64, F.one_hot(lab, C).double())
    (0.1 * loss).backward()
    assert float(out[0]) == pytest.approx(float(loss), rel=1e-5)
    assert float(out[1]) == pytest.approx(float(acc), abs=1e-6)
    assert rel(dl, l64.grad) < FP32_TOL


def _run_triplet(ctx, lab, emb, margin, scale=1.0):
    from ugaitnet_b200 import ops
    e = torch.tensor(emb, device="cuda")
    n, B = (e.shape[0], e.shape[1]) if e.dim() == 3 else (1, e.shape[0])
    out = torch.zeros(2, device="cuda")
    de = torch.zeros_like(e)
    ws = torch.zeros(ops.triplet_workspace_bytes(n, B) // 4 + 8, device="cuda")
    ops.triplet_all(ctx, e, torch.tensor(lab).int().cuda(), margin, scale, out, de, ws)
    return out.cpu(), de.cpu()


def test_triplet_golden_and_gradient(ctx, golden_dir):
    z = np.load(os.path.join(golden_dir, "triplet.npz"))
    for i in range(int(z["n"])):
        lab, emb, margin = z[f"lab{i}"], z[f"emb{i}"], float(z[f"margin{i}"])
        out, de = _run_triplet(ctx, lab, emb, margin)
        assert float(out[0]) == pytest.approx(float(z[f"loss{i}"]), rel=1e-3)
        e64 = torch.tensor(emb, dtype=torch.float64, requires_grad=True)
        loss, cnt = O.triplet_loss_all(torch.tensor(lab), e64, margin)
        loss.backward()
        # active-set counts may differ by a handful of boundary triplets between fp32 and fp64
        assert abs(float(out[1]) - float(cnt.sum())) <= max(3, 1e-3 * float(cnt.sum()))
        assert rel(de, e64.grad) < 5e-3


def test_triplet_large_balanced_batch(ctx):
    rng = np.random.default_rng(3)
    ids, per, d = 48, 2, 128         # cfg2-like: B=96
    lab = np.repeat(np.arange(ids), per).astype(np.float32)
    e = rng.normal(size=(ids * per, d))
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    out, de = _run_triplet(ctx, lab, e.astype(np.float32), 0.2, scale=0.5)
    e64 = torch.tensor(e.astype(np.float32), dtype=torch.float64, requires_grad=True)
    loss, cnt = O.triplet_loss_all(torch.tensor(lab), e64, 0.2)
    (0.5 * loss).backward()
    assert float(out[0]) == pytest.approx(float(loss), rel=1e-5)
    assert float(out[1]) == float(cnt.sum())
    assert rel(de, e64.grad) < 1e-4


def test_triplet_no_active_triplets_is_zero(ctx):
    lab = np.array([0, 0, 1, 1], dtype=np.float32)
    e = np.array([[1, 0], [1, 0], [-1, 0], [-1, 0]], dtype=np.float32)
    out, de = _run_triplet(ctx, lab, e, 0.2)
    assert float(out[0]) == 0.0 and float(out[1]) == 0.0 and float(de.abs().max()) == 0.0


@pytest.mark.parametrize("case", ["balanced", "ragged", "one_class", "tc_gram"])
def test_triplet_hard_vs_oracle(ctx, case):
    """ugn_triplet_hard (compile_hard: tfa TripletHardLoss, nets/mj_uwyhNets_ba.py:1302-1306) against the fp64 restatement of
    the tfa algorithm: loss, active anchors and gradient; an anchor without positives, a batch without negatives,
    duplicated rows, and the tensor-core Gram variant."""
    from ugaitnet_b200 import ops
    rng = np.random.default_rng(5)
    if case == "balanced":
        lab, d, margin = np.repeat(np.arange(24), 4), 128, 0.2
    elif case == "ragged":
        lab, d, margin = np.array([0, 0, 0, 1, 2, 2, 3, 3, 3, 3, 4]), 32, 1.0
    elif case == "one_class":
        lab, d, margin = np.zeros(6, dtype=np.int64), 16, 0.5
    else:
        lab, d, margin = np.repeat(np.arange(64), 2), 256, 0.2
    B = len(lab)
    e = rng.normal(size=(B, d)).astype(np.float32)
    e[1] = e[0]                                   # identical rows: zero distance, zero gradient between them
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    x = torch.tensor(e).cuda()
    x16 = None
    if case == "tc_gram":
        hi = x.half()
        x16 = torch.stack([hi, (x - hi.float()).half()]).contiguous()
    out, de = torch.zeros(2, device="cuda"), torch.zeros(B, d, device="cuda")
    ws = torch.zeros(ops.triplet_workspace_bytes(1, B) // 4 + 16, device="cuda")
    ops.triplet_hard(ctx, x, torch.tensor(lab).int().cuda(), margin, 0.5, out, de, ws, emb16=x16)
    ctx.check()
    e64 = torch.tensor(e, dtype=torch.float64, requires_grad=True)
    loss, act = O.triplet_hard_loss(torch.tensor(lab), e64, margin)
    (0.5 * loss).backward()
    tol = 1e-4 if case == "tc_gram" else 1e-5
    assert float(out[0]) == pytest.approx(float(loss), rel=tol)
    assert float(out[1]) == float(act)
    assert rel(de, e64.grad) < (2e-3 if case == "tc_gram" else 1e-4)


def test_adam_and_sgd_step(ctx):
    from ugaitnet_b200 import ops
    n = 64 * 5
    g = torch.Generator().manual_seed(1)
    w0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    off = torch.tensor([0, 128, 192, n], dtype=torch.int64, device="cuda")
    l2 = torch.tensor([5e-5, 0.0, 1e-3], device="cuda")
    l2full = torch.cat([torch.full((128,), 5e-5), torch.zeros(64), torch.full((n - 192,), 1e-3)]).double()
    w, m, v = w0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    reg = torch.zeros(1, device="cuda")
    P = {"w": w0.double().clone()}
    M, V = {"w": torch.zeros(n, dtype=torch.float64)}, {"w": torch.zeros(n, dtype=torch.float64)}
    for t in range(1, 4):
        import math
        lr_t = 1e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        ops.adam_step(ctx, w, gr.cuda(), m, v, off, l2, lr_t, gscale=0.5, reg_out=reg)
        G = {"w": gr.double() * 0.5 + 2 * l2full * P["w"]}
        expect_reg = float((l2full * P["w"] ** 2).sum())
        O.adam_step(P, G, M, V, t, lr=1e-3)
        assert float(reg) == pytest.approx(expect_reg, rel=1e-5)
        assert rel(w, P["w"]) < 1e-6
    w, v = w0.clone().cuda(), torch.zeros(n, device="cuda")
    ops.sgd_step(ctx, w, gr.cuda(), v, off, l2, 0.01, momentum=0.9)
    ref = w0.double() - 0.01 * (gr.double() + 2 * l2full * w0.double())
    assert rel(w, ref) < 1e-6


@pytest.mark.parametrize("name", ["knn_small", "knn_dups", "knn_k7"])
def test_knn_bit_exact_vs_oracle_and_sklearn_golden(ctx, golden_dir, name):
    from ugaitnet_b200.knn import KNeighborsClassifier
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    k = int(z["k"])
    clf = KNeighborsClassifier(n_neighbors=k).fit(z["G"], z["y"])
    pred = clf.predict(z["Q"])
    d2, idx = clf.kneighbors_exact(z["Q"])
    od2, oidx = O.knn_search(z["G"], z["Q"], k)
    opred = O.knn_vote(z["y"][oidx])
    assert np.array_equal(idx, oidx)                    # bit-exact indices (same (dist, idx) rule)
    assert np.array_equal(pred, opred)                  # bit-exact labels
    assert np.allclose(d2, od2, rtol=1e-12, atol=1e-15)
    d2b, _ = O.knn_search(z["G"], z["Q"], k + 1)
    nb = d2b[:, k - 1] != d2b[:, k]
    assert np.array_equal(pred[nb], z["pred"][nb])      # == sklearn wherever sklearn is well defined


@pytest.mark.parametrize("N,D,Q,k", [(5000, 100, 37, 3), (20000, 256, 300, 7), (3000, 64, 130, 20), (70000, 32, 64, 3),
                                     (6000, 328, 150, 3)])   # D > 256: streaming-query variant
def test_knn_tensor_core_scan_bit_exact(ctx, N, D, Q, k):
    """tcgen05 distance GEMM + fused top-k (ugn_knn_topk_tc) against the C oracle: clustered unit-norm
    descriptors with exact duplicates (distance ties), ragged N / D / Q (not multiples of the tiles)."""
    from ugaitnet_b200.knn import KNeighborsClassifier
    if not ctx.has_tcgen05:
        pytest.skip("needs sm_100")
    rng = np.random.default_rng(N + D)
    cent = rng.normal(size=(31, D)).astype(np.float32)
    y = rng.integers(0, 31, N).astype(np.int32)
    G = cent[y] + 0.3 * rng.normal(size=(N, D)).astype(np.float32)
    G /= np.linalg.norm(G, axis=1, keepdims=True)
    dup = rng.integers(0, N, N // 50)
    G[dup] = G[(dup * 7 + 3) % N]                       # exact duplicate rows -> distance ties
    Qm = G[rng.integers(0, N, Q)] + 0.05 * rng.normal(size=(Q, D)).astype(np.float32)
    Qm[: Q // 4] = G[rng.integers(0, N, Q // 4)]        # queries that ARE gallery rows (distance 0)
    clf = KNeighborsClassifier(n_neighbors=k).fit(G, y)
    assert clf.use_tc
    pred = clf.predict(Qm)
    d2, idx = clf.kneighbors_exact(Qm)
    od2, oidx = O.knn_search(G, Qm, k)
    assert np.array_equal(idx, oidx)
    assert np.array_equal(pred, O.knn_vote(y[oidx]))
    assert np.allclose(d2, od2, rtol=1e-12, atol=1e-15)
    ctx.check()
    print(f"[knn-tc N={N} D={D} Q={Q} k={k}] flagged (recomputed exactly): {clf.flagged_queries()} of {Q}")
    assert clf.flagged_queries() <= Q // 4 + Q // 10    # the proof may fail on zero-distance / duplicate queries only


def test_knn_exact_fallback_path(ctx):
    """Gallery made of near-identical rows: the containment proof cannot separate k-th from KC-th
    neighbour, every query is flagged and the fp64 brute-force kernel must still be bit-exact."""
    from ugaitnet_b200.knn import KNeighborsClassifier
    if not ctx.has_tcgen05:
        pytest.skip("needs sm_100")
    rng = np.random.default_rng(9)
    base = rng.normal(size=(1, 48)).astype(np.float32)
    G = (base + 1e-6 * rng.normal(size=(4000, 48))).astype(np.float32)
    y = rng.integers(0, 5, 4000).astype(np.int32)
    Qm = (base + 1e-6 * rng.normal(size=(20, 48))).astype(np.float32)
    clf = KNeighborsClassifier(n_neighbors=3).fit(G, y)
    d2, idx = clf.kneighbors_exact(Qm)
    od2, oidx = O.knn_search(G, Qm, 3)
    assert clf.flagged_queries() > 0
    assert np.array_equal(idx, oidx) and np.allclose(d2, od2, rtol=1e-12, atol=1e-18)


def test_knn_simt_scan_still_bit_exact(ctx, golden_dir, monkeypatch):
    from ugaitnet_b200.knn import KNeighborsClassifier
    monkeypatch.setenv("UGN_KNN_SIMT", "1")
    z = np.load(os.path.join(golden_dir, "knn_dups.npz"))
    k = int(z["k"])
    clf = KNeighborsClassifier(n_neighbors=k).fit(z["G"], z["y"])
    assert not clf.use_tc
    _, idx = clf.kneighbors_exact(z["Q"])
    _, oidx = O.knn_search(z["G"], z["Q"], k)
    assert np.array_equal(idx, oidx)


def test_knn_sharded_merge_equals_single(ctx, golden_dir):
    from ugaitnet_b200.knn import KNeighborsClassifier, knn_sharded_local
    z = np.load(os.path.join(golden_dir, "knn_dups.npz"))
    k = int(z["k"])
    single = KNeighborsClassifier(n_neighbors=k).fit(z["G"], z["y"])
    _, idx = single.kneighbors_exact(z["Q"])
    pred, midx = knn_sharded_local(z["G"], z["y"], z["Q"], k, shards=4)
    assert np.array_equal(midx, idx) and np.array_equal(pred, single.predict(z["Q"]))


def test_errors_are_loud(ctx):
    from ugaitnet_b200 import ops
    from ugaitnet_b200._ffi import UgnError
    with pytest.raises(UgnError, match="no CPU fallback|device"):
        ops.flatten_chw(ctx, torch.zeros(1, 2, 2, 4), torch.zeros(1, 16, device="cuda"))
    with pytest.raises(UgnError, match="shape"):
        ops.flatten_chw(ctx, torch.zeros(1, 2, 2, 4, device="cuda"), torch.zeros(1, 15, device="cuda"))


@pytest.mark.parametrize("dt", ["f32", "f16"])
def test_pack_input_expand_and_mirror(ctx, dt):
    """ugn_pack_input_expand vs the numpy restatement of the generator's expansion (:806-812) and of
    mj_mirrorsequence (data/mj_augmentation.py:12-32: LR flip + even channels negated)."""
    from ugaitnet_b200 import ops
    from ugaitnet_b200.expand import NOISE, expand_on_host, mirror_sequence
    rng = np.random.default_rng(3)
    B0, C, H, Cp = 3, 6, 10, 32 if dt == "f16" else 6
    base = rng.normal(size=(B0, C, H, H)).astype(np.float32)
    src = np.array([0, 0, 1, 2, 2, 1, 0], dtype=np.int32)
    en = np.array([1, 0, 1, 1, 0, 1, 1], dtype=np.float32)
    mir = np.array([0, 0, 1, 0, 1, 1, 0], dtype=np.uint8)
    ref = expand_on_host(base, src, en)
    for b in range(len(src)):
        if mir[b] and en[b]:
            ref[b] = mirror_sequence(ref[b])
    if dt == "f32":
        out = torch.zeros(len(src), H, H, Cp, device="cuda")
    else:
        out = torch.zeros(2, len(src), H, H, Cp, dtype=torch.float16, device="cuda")
    ops.pack_input_expand(ctx, torch.tensor(base).cuda(), torch.tensor(src).cuda(), torch.tensor(en).cuda(),
                          torch.tensor(mir).cuda(), out, noise=NOISE)
    got = (out if dt == "f32" else out.float().sum(0))[..., :C].permute(0, 3, 1, 2).cpu().numpy()
    # fp16 planes: 1e-9 is below the smallest fp16 subnormal and becomes 0 -- immaterial, the row's use-flag
    # gates the branch output to exactly 0 either way (nets/mj_uwyhNets_ba.py:51-54)
    assert np.allclose(got, ref, rtol=1e-6 if dt == "f32" else 2e-6, atol=1e-12 if dt == "f32" else 1e-7)   # fp16 lo plane: subnormal below 6e-5
    if Cp > C:
        assert float(out.float().abs()[..., C:].max()) == 0.0
    if dt == "f32":
        assert np.allclose(got[1], NOISE, rtol=1e-3)   # disabled row == the reference's noise constant


@pytest.mark.parametrize("use_avg", [True, False])
def test_video_level_pool_and_vote(ctx, use_avg):
    """a13: per-video descriptor pooling + statistics.mode vote vs the oracle's loop-for-loop restatement of
    mains/mj_testUWYHGaitNet_open_tum.py:355-420 (ragged videos, single-row videos, vote ties)."""
    from ugaitnet_b200.video import video_groups, pool_per_video, mode_per_video
    rng = np.random.default_rng(4)
    N, D = 157, 70
    vids = rng.integers(0, 23, N)
    vids[:3] = 99                                   # unsorted ids, a video with 3 rows
    codes = rng.normal(size=(N, D)).astype(np.float32)
    labels = rng.integers(0, 4, N).astype(np.int32)     # few classes -> frequent ties
    preds = rng.integers(0, 4, N).astype(np.int32)
    uv, oc_, ol, op = O.video_level(codes, labels, vids, preds, use_avg)
    u, order, offsets = video_groups(vids)
    assert np.array_equal(u, uv)
    pc = pool_per_video(ctx, codes, order, offsets, use_avg).cpu().numpy()
    assert np.allclose(pc, oc_, rtol=1e-6, atol=1e-7)
    assert np.array_equal(mode_per_video(ctx, labels, order, offsets).cpu().numpy(), ol)      # bit-exact votes
    assert np.array_equal(mode_per_video(ctx, preds, order, offsets).cpu().numpy(), op)


def test_open_world_evaluation_end_to_end(ctx):
    from ugaitnet_b200.video import evaluate_open_world
    rng = np.random.default_rng(8)
    D, ncls = 48, 12
    cent = rng.normal(size=(ncls, D)).astype(np.float32)

    def make(nv):
        labs, vids, codes = [], [], []
        for v in range(nv):
            c = v % ncls
            n = int(rng.integers(2, 7))
            codes.append(cent[c] + 0.4 * rng.normal(size=(n, D)).astype(np.float32))
            labs += [c] * n
            vids += [1000 + v] * n
        return np.vstack(codes), np.array(labs, dtype=np.int32), np.array(vids)
    cg, lg, vg = make(60)
    ct, lt, vt = make(40)
    res = evaluate_open_world(cg, lg, vg, ct, lt, vt, knn=3)
    pred_o, _ = O.knn_predict(cg, lg, ct, 3)
    assert np.array_equal(res["pred"].cpu().numpy(), pred_o)
    uv, cvg, lvg, _ = O.video_level(cg, lg, vg)
    _, cvt, lvt, pvt = O.video_level(ct, lt, vt, pred_o)
    assert np.array_equal(res["pred_vid"].cpu().numpy(), pvt) and np.array_equal(res["labs_vid"].cpu().numpy(), lvt)
    pm, _ = O.knn_predict(cvg, lvg, cvt, 3)
    # pooled descriptors agree to fp32 rounding; the merged k-NN must give the same labels
    assert np.array_equal(res["pred_vid_merged"].cpu().numpy(), pm)
    acc, acc_vid, score = res["summary"]
    assert acc > 0.9 and acc_vid >= acc - 0.05 and score > 0.9


def test_device_side_augmentation_equals_host_restatement():
    """ugn_pack_input_augment (integer shift of the random transform, optical-flow magnitude clip, mirror,
    missing-modality expansion fused into the input pack) == the numpy restatement of the generator's statements
    (data/mj_dataGeneratorMMUWYHsingle.py:718-746, data/mj_augmentation.py:12-50; the shift is pinned to
    scipy.ndimage.affine_transform on CPU)."""
    import random
    from ugaitnet_b200._ffi import TRef, check, lib, stream_ptr
    from ugaitnet_b200.expand import augment_on_host, expansion_pattern
    from ugaitnet_b200.net import OF_CLIP_HI, OF_CLIP_LO, OF_CLIP_VAL
    from ugaitnet_b200 import ops
    ctx = ops.get_ctx()
    rng = np.random.default_rng(2)
    B0, E, C, H = 5, 4, 50, 60
    base = np.clip(rng.normal(0, 0.8, size=(B0, C, H, H)), -3.3, 3.3).astype(np.float32)
    src, use = expansion_pattern(B0, E, 3, random.Random(1))
    B = len(src)
    mirror = (rng.integers(0, 2, B)).astype(np.uint8)
    shift = rng.choice([-5, -3, 0, 3, 5], size=(B, 2)).astype(np.int8)
    clip = (rng.integers(0, 2, B)).astype(np.uint8)
    ref = augment_on_host(base, src, use[:, 0], mirror, shift, clip)
    out = torch.zeros(B, H, H, 64, device="cuda")
    t = [torch.as_tensor(a).cuda() for a in (base, src.astype(np.int32), use[:, 0].astype(np.float32), mirror, shift, clip)]
    R = [TRef(x) for x in t] + [TRef(out)]
    check(lib.ugn_pack_input_augment(ctx.h, R[0].ptr, R[1].ptr, R[2].ptr, R[3].ptr, R[4].ptr, R[5].ptr, OF_CLIP_LO,
                                     OF_CLIP_HI, OF_CLIP_VAL, 1e-9, R[6].ptr, stream_ptr()))
    got = out[..., :C].permute(0, 3, 1, 2).cpu().numpy()
    assert np.array_equal(got, ref)
    assert float(out[..., C:].abs().max()) == 0.0


@pytest.mark.parametrize("B,d,dt", [(96, 2048, torch.float16), (120, 256, torch.bfloat16), (512, 2048, torch.float16)])
def test_triplet_gram_on_tensor_cores(ctx, B, d, dt):
    """ugn_triplet_all_tc (north_star: the B x B pairwise-distance matrix as a tensor-core GEMM): loss, active count and
    gradient against the fp64 oracle, with duplicated rows (an expanded batch repeats sequences) that must keep an
    exactly zero distance, and against the FFMA-Gram kernel."""
    from ugaitnet_b200 import ops
    from ugaitnet_b200._ffi import TRef, check, lib, stream_ptr
    rng = np.random.default_rng(B)
    per = 4 if B != 512 else 2
    lab = np.repeat(np.arange(B // per), per).astype(np.int32)
    e = rng.normal(size=(B, d)).astype(np.float32)
    e[1] = e[0]                                            # identical rows with the same label
    e[5] = e[4] + 1e-4 * rng.normal(size=d)                # and a near-duplicate
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    x = torch.tensor(e).cuda()
    hi = x.to(dt)
    x16 = torch.stack([hi, (x - hi.float()).to(dt)]).contiguous()
    out, de = torch.zeros(2, device="cuda"), torch.zeros(B, d, device="cuda")
    ws = torch.zeros(ops.triplet_workspace_bytes(1, B) // 4 + 16, device="cuda")
    R = [TRef(t) for t in (x, x16, torch.tensor(lab).cuda(), out, de, ws)]
    check(lib.ugn_triplet_all_tc(ctx.h, R[0].ptr, R[1].ptr, R[2].ptr, 0.2, 1.0, R[3].ptr, R[4].ptr, R[5].ptr, stream_ptr()))
    ctx.check()
    e64 = torch.tensor(e, dtype=torch.float64, requires_grad=True)
    loss, cnt = O.triplet_loss_all(torch.tensor(lab), e64, 0.2)
    loss.backward()
    assert float(out[0]) == pytest.approx(float(loss), rel=1e-4 if dt is torch.float16 else 1e-3)
    assert abs(float(out[1]) - float(cnt.sum())) <= max(2, 1e-4 * float(cnt.sum()))
    assert rel(de, e64.grad) < (2e-3 if dt is torch.float16 else 1e-2)
    out2, de2 = _run_triplet(ctx, lab.astype(np.float32), e, 0.2)          # FFMA Gram
    assert float(out[0]) == pytest.approx(float(out2[0]), rel=1e-4 if dt is torch.float16 else 1e-3)
    # no split-K: one accumulation order per element, so identical rows are at distance EXACTLY 0 (no gradient flows
    # between rows 0 and 1) and the diagonal is exactly 0; (a,b) and (b,a) add the hi*lo and lo*hi passes in swapped
    # order and agree to fp32 rounding
    accb, x2b = 256, (B * 4 + 255) // 256 * 256
    D = ws[(accb + x2b) // 4:(accb + x2b) // 4 + B * B].view(B, B)
    assert float(D[0, 1]) == 0.0 and float(D[1, 0]) == 0.0 and float(D.diagonal().abs().max()) == 0.0
    assert float((D - D.t()).abs().max()) <= 1e-5


@pytest.mark.parametrize("case", ["mixed", "all_pos", "far_negatives", "ignored"])
def test_pair_verif_loss_vs_oracle(ctx, case):
    """ugn_pair_verif_loss (UWYHNet.build's VerifLossLayer, nets/mj_loss.py:65-95) against the fp64 restatement: value
    and gradient; only positives (sqrt of an empty sum), negatives beyond the margin (inactive hinge), labels that are
    neither 0 nor 1 (the reference's tf.where(equal(labels, 1|0)) ignores them)."""
    from ugaitnet_b200._ffi import TRef, check, lib, stream_ptr
    rng = np.random.default_rng(9)
    B, d = 12, 48
    e = rng.normal(size=(2 * B, d)).astype(np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    lab = rng.integers(0, 2, B)
    margin = 6.0
    if case == "all_pos":
        lab[:] = 1
    elif case == "far_negatives":
        margin = 0.5
    elif case == "ignored":
        lab[::3] = 7
    x = torch.tensor(e).cuda()
    out, de = torch.zeros(2, device="cuda"), torch.full((2 * B, d), 9.0, device="cuda")
    ws = torch.zeros(8, device="cuda")
    R = [TRef(t) for t in (x, torch.tensor(lab).int().cuda(), out, de, ws)]
    check(lib.ugn_pair_verif_loss(ctx.h, R[0].ptr, R[1].ptr, margin, 0.5, R[2].ptr, R[3].ptr, R[4].ptr, stream_ptr()))
    ctx.check()
    e64 = torch.tensor(e, dtype=torch.float64, requires_grad=True)
    loss = O.pair_verif_loss(torch.tensor(lab), e64, margin)
    (0.5 * loss).backward()
    assert float(out[0]) == pytest.approx(float(loss.detach()), rel=1e-5, abs=1e-7)
    g = e64.grad
    if float(g.norm()) == 0.0:
        assert float(de.abs().max()) == 0.0
    else:
        assert rel(de, g) < 1e-5
    if case == "far_negatives":
        assert float(out[1]) == 0.0


@pytest.mark.parametrize("geom", [((2, 9, 20, 21, 1), 6, (3, 5, 5), (1, 2, 2)), ((3, 8, 9, 10, 8), 12, (3, 3, 3), (2, 2, 2)),
                                  ((2, 5, 6, 6, 70), 66, (3, 2, 2), (1, 1, 1)), ((4, 2, 1, 1, 16), 8, (2, 1, 1), (1, 1, 1))])
def test_conv3d_ops_vs_torch(ctx, geom):
    """ugn_conv3d_{fwd,wgrad,dgrad} (use3D branches, nets/mj_uwyhNets_ba.py:346-363: strided 'valid' channels-last Conv3D)
    against torch.nn.functional.conv3d and its autograd in fp64."""
    import torch.nn.functional as F
    from ugaitnet_b200._ffi import TRef, check, lib, stream_ptr
    (B, T, H, W, C), Co, k, s = geom
    g = torch.Generator().manual_seed(B * 7 + C)
    x = torch.randn(B, T, H, W, C, generator=g)
    w = torch.randn(Co, k[0], k[1], k[2], C, generator=g) * 0.2
    bias = torch.randn(Co, generator=g) * 0.1
    x64 = x.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
    w64 = w.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
    b64 = bias.double().requires_grad_(True)
    z = F.conv3d(x64, w64, b64, stride=s)
    y64 = F.leaky_relu(z, 0.3)
    dy = torch.randn(y64.shape, generator=g, dtype=torch.float64)
    z.backward(dy)                                                  # gradients w.r.t. the pre-activation dz = dy
    To, Ho, Wo = y64.shape[2:]
    xd, wd, bd = x.cuda(), w.cuda(), bias.cuda()
    y = torch.empty(B, To, Ho, Wo, Co, device="cuda")
    R = [TRef(t) for t in (xd, wd, bd, y)]
    check(lib.ugn_conv3d_fwd(ctx.h, R[0].ptr, R[1].ptr, R[2].ptr, R[3].ptr, s[0], s[1], s[2], 2, 0.3, stream_ptr()))
    assert rel(y.permute(0, 4, 1, 2, 3), y64.detach()) < 1e-5
    dz = dy.permute(0, 2, 3, 4, 1).contiguous().float().cuda()
    dw, db, dx = torch.full_like(wd, 7.0), torch.full_like(bd, 7.0), torch.full_like(xd, 7.0)
    R2 = [TRef(t) for t in (dz, dw, db, dx)]
    check(lib.ugn_conv3d_wgrad(ctx.h, R[0].ptr, R2[0].ptr, R2[1].ptr, R2[2].ptr, s[0], s[1], s[2], stream_ptr()))
    check(lib.ugn_conv3d_dgrad(ctx.h, R2[0].ptr, R[1].ptr, R2[3].ptr, s[0], s[1], s[2], stream_ptr()))
    ctx.check()
    assert rel(dw.permute(0, 4, 1, 2, 3), w64.grad) < 1e-5
    assert rel(db, b64.grad) < 1e-5
    assert rel(dx.permute(0, 4, 1, 2, 3), x64.grad) < 1e-5


@pytest.mark.parametrize("merge", [0, 1, 2])
@pytest.mark.parametrize("cfg", [(3, 200, 37, 1, 2, True), (2, 2048, 256, 2, 0, False), (1, 64, 5, 0, 1, True)])
def test_fusion_plus_fc1_single_kernel(ctx, merge, cfg):
    """ugn_fuse_fc1_fwd (north_star: gated fusion + l2_normalize + FC1 "code" in ONE kernel, nets/mj_uwyhNets_ba.py:1163-1203)
    against the oracle pieces and against the two-kernel path: signature, winners, 16-bit planes, code, dropped code; rows
    with every modality missing, exact |x| ties, odd nc, with and without normalisation / dropout mask."""
    from ugaitnet_b200 import ops
    from ugaitnet_b200._ffi import TRef, check, lib, ptr_array, stream_ptr
    nmods, d, nc, act, P, normalize = cfg
    B = 11
    g = torch.Generator().manual_seed(merge * 7 + d)
    br = [torch.randn(B, d, generator=g) for _ in range(nmods)]
    if nmods > 1:
        br[1][:, :20] = br[0][:, :20]
        br[1][:, 20:30] = -br[0][:, 20:30]
    fl = [(torch.rand(B, 1, generator=g) > 0.3).float() for _ in range(nmods)]
    for f in fl:
        f[0] = 0.0
    fl[0][1] = 1.0
    W = torch.randn(nc, d, generator=g) / d ** 0.5
    bias = torch.randn(nc, generator=g) * 0.1
    cm = (torch.rand(B, nc, generator=g) > 0.4).float() / 0.6
    dev = lambda t: t.cuda()
    sig, win, inv = torch.zeros(B, d, device="cuda"), torch.zeros(B, d, dtype=torch.uint8, device="cuda"), torch.zeros(B, 2, device="cuda")
    sig16 = torch.zeros(P, B, d, dtype=torch.float16, device="cuda") if P else None
    code, dropc = torch.zeros(B, nc, device="cuda"), torch.zeros(B, nc, device="cuda")
    rb, rf = [TRef(dev(t)) for t in br], [TRef(dev(f)) for f in fl]
    R = [TRef(t) if t is not None else None for t in (sig, sig16, win, inv, dev(W), dev(bias), code, dev(cm), dropc)]
    pp = lambda r: None if r is None else r.ptr
    check(lib.ugn_fuse_fc1_fwd(ctx.h, nmods, ptr_array(rb), ptr_array(rf), pp(R[0]), pp(R[1]), pp(R[2]), pp(R[3]), merge,
                               int(normalize), pp(R[4]), pp(R[5]), pp(R[6]), pp(R[7]), pp(R[8]), act, 0.3, stream_ptr()))
    ctx.check()
    b64 = [t.double() for t in br]
    ref = O.merge_modalities([t * f.double() for t, f in zip(b64, fl)], merge)
    if normalize:
        ref = O.l2_normalize(ref, 1)
    assert torch.allclose(sig.cpu().double(), ref, atol=2e-6, rtol=1e-5)
    assert float(sig[0].abs().max()) == 0.0
    z = ref @ W.double().t() + bias.double()
    want = O._act(z, act, 0.3)
    assert rel(code, want) < 1e-5
    assert rel(dropc, want * cm.double()) < 1e-5
    # the two-kernel path gives the same signature, winners and planes
    sig2, win2, inv2 = torch.zeros_like(sig), torch.zeros_like(win), torch.zeros_like(inv)
    sig16b = torch.zeros_like(sig16) if P else None
    ops.fuse_fwd(ctx, [dev(t) for t in br], [dev(f) for f in fl], sig2, sig16b, win2, inv2, merge, normalize)
    assert torch.equal(win, win2) and torch.allclose(sig, sig2, atol=1e-7) and torch.allclose(inv, inv2, rtol=1e-6)
    if P:
        assert torch.allclose(sig16.float(), sig16b.float(), atol=1e-6)
