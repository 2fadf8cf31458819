"""CPU tests: the oracle against the committed golden vectors (reference-produced where the
reference can run here) and against itself (literal vs general triplet form)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ugait_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tie_free(dist):
    d = np.sort(dist, axis=1)
    return (np.diff(d, axis=1) > 1e-7).all(axis=1)


@pytest.mark.parametrize("name", ["knn_small", "knn_dups", "knn_k7"])
def test_knn_oracle_matches_sklearn_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    k = int(z["k"])
    pred, idx = O.knn_predict(z["G"], z["y"], z["Q"], k)
    # labels: bit-exact on every query whose k-th / (k+1)-th neighbours are not exactly tied
    d2, _ = O.knn_search(z["G"], z["Q"], k + 1)
    boundary_tie = d2[:, k - 1] == d2[:, k]
    assert (pred[~boundary_tie] == z["pred"][~boundary_tie]).all()
    # indices: bit-exact on tie-free queries; as sets elsewhere (sklearn's order inside an exact
    # tie is a heap artefact)
    tf = _tie_free(z["dist"]) & ~boundary_tie
    assert tf.sum() > 0
    assert (idx[tf] == z["idx"][tf]).all()
    same_set = [set(a) == set(b) for a, b in zip(idx[~boundary_tie], z["idx"][~boundary_tie])]
    assert all(same_set)


def _load_c_oracle():
    path = os.path.join(ROOT, "oracle", "libknn_oracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    return ctypes.CDLL(path)


@pytest.mark.parametrize("name", ["knn_small", "knn_dups", "knn_k7"])
def test_c_knn_oracle_equals_numpy_oracle(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    lib = _load_c_oracle()
    G, Q, y, k = np.ascontiguousarray(z["G"]), np.ascontiguousarray(z["Q"]), np.ascontiguousarray(z["y"]), int(z["k"])
    idx = np.empty((Q.shape[0], k), dtype=np.int64)
    d2 = np.empty((Q.shape[0], k), dtype=np.float64)
    rc = lib.knn_oracle_search(G.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(G.shape[0]), ctypes.c_int64(G.shape[1]),
                               Q.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(Q.shape[0]), ctypes.c_int(k),
                               idx.ctypes.data_as(ctypes.c_void_p), d2.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    pred = np.empty(Q.shape[0], dtype=np.int32)
    lib.knn_oracle_vote(idx.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(Q.shape[0]), ctypes.c_int(k),
                        y.ctypes.data_as(ctypes.c_void_p), pred.ctypes.data_as(ctypes.c_void_p))
    p2, i2 = O.knn_predict(G, y, Q, k)
    assert (idx == i2).all()
    assert (pred == p2).all()


def test_eer_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "eer.npz"))
    for i in range(int(z["n"])):
        eer, thr = O.eer_verif_dist(z[f"y{i}"], z[f"d{i}"])
        assert eer == pytest.approx(float(z["eer"][i]), abs=1e-12)
        assert thr == pytest.approx(float(z["thr"][i]), abs=1e-12)
    # the reference's own demo known-answer (nets/mj_metrics.py:28-35)
    assert float(z["eer"][0]) == pytest.approx(0.25) and float(z["thr"][0]) == pytest.approx(0.07)


def test_triplet_general_form_equals_literal_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "triplet.npz"))
    for i in range(int(z["n"])):
        lab, emb, margin = z[f"lab{i}"], z[f"emb{i}"], float(z[f"margin{i}"])
        loss, cnt = O.triplet_loss_all(torch.tensor(lab), torch.tensor(emb, dtype=torch.float64), margin)
        assert float(loss) == pytest.approx(float(z[f"loss{i}"]), rel=1e-9)
        assert np.allclose(cnt.numpy(), z[f"cnt{i}"])


def test_reference_demo_vectors_give_zero_loss():
    # nets/triplet_loss_all.py:115-116: trivially separated classes -> 0 for margin <= 1
    logits = np.array([[1.1, 1.2, 1.4], [1.09, 1.21, 1.41], [0.25, 0.45, 0.75], [0.23, 0.43, 0.7],
                       [1.5, 2.5, 3.5], [1.55, 2.75, 3.8]], dtype=np.float32)
    labels = np.array([1, 1, 2, 2, 3, 3], dtype=np.float32)
    loss, _ = O.triplet_loss_all_literal_np(labels, logits, 0.2)
    assert loss == 0.0


def test_sign_max_and_maximum_tie_rules():
    a = torch.tensor([[1.0, -2.0, 0.0, 3.0]])
    b = torch.tensor([[-1.0, 2.0, 0.0, -4.0]])
    out = O.merge_modalities([a, b], O.MERGE_SIGNMAX)
    assert out.tolist() == [[1.0, -2.0, 0.0, -4.0]]      # ties -> lowest modality index
    out = O.merge_modalities([a, b], O.MERGE_MAX)
    assert out.tolist() == [[1.0, 2.0, 0.0, 3.0]]
    out = O.merge_modalities([a, b], O.MERGE_AVG)
    assert out.tolist() == [[0.0, 0.0, 0.0, -0.5]]


def test_l2_normalize_eps_path():
    x = torch.zeros(2, 8)
    x[1, 0] = 3.0
    y = O.l2_normalize(x)
    assert torch.all(y[0] == 0) and y[1, 0] == pytest.approx(1.0)


def test_synth_batch_expansion_contract():
    cfg = O.NetConfig(nd=16, nclasses=150)
    xs, fl, lab = O.synth_batch(cfg, base_rows=6, expand=4, seed=1)
    assert xs[0].shape == (24, 50, 60, 60) and xs[1].shape == (24, 25, 60, 60)
    F = np.concatenate(fl, 1)
    assert (F[::4] == 1).all()                       # row i*E has every modality
    assert (F.sum(1) >= 1).all()                     # never all missing
    for i in range(6):
        if i % 2 == 1:                               # odd i: exactly one modality enabled
            assert (F[i * 4 + 1:(i + 1) * 4].sum(1) == 1).all()
    for m in range(3):
        off = F[:, m] == 0
        assert np.all(xs[m][off] == np.float32(1e-9))
    assert (lab.reshape(6, 4) == lab.reshape(6, 4)[:, :1]).all()


def test_oracle_step_gradients_finite_and_adam_decreases_loss():
    torch.manual_seed(0)
    cfg = O.NetConfig(in_channels=(4, 3), filters_numbers=(8, 8, 16, 16), nd=16, nclasses=5, merge=O.MERGE_SIGNMAX,
                      wid=0.1)
    P = O.init_params(cfg, seed=3, dtype=torch.float64)
    xs, fl, lab = O.synth_batch(cfg, base_rows=4, expand=2, seed=5, kinds=("of", "gray"))
    xs = [torch.tensor(x, dtype=torch.float64) for x in xs]
    fl = [torch.tensor(f, dtype=torch.float64) for f in fl]
    lab = torch.tensor(lab % 5)
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    losses = []
    for t in range(1, 6):
        res, G = O.loss_and_grads(xs, fl, lab, P, cfg)
        assert all(torch.isfinite(g).all() for g in G.values())
        losses.append(float(res["loss"]))
        O.adam_step(P, G, M, V, t, lr=1e-3)
    assert losses[-1] < losses[0]


def test_expansion_pattern_matches_oracle_generator():
    """ugaitnet_b200.expand.expansion_pattern (host side of the device-side expansion) draws the same
    missing-modality pattern as the oracle's restatement of the reference generator (:791-803)."""
    import random
    from ugaitnet_b200.expand import expansion_pattern, expand_on_host
    oc = O.NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nclasses=10)
    for E in (2, 3, 4):
        xs, fl, lab = O.synth_batch(oc, base_rows=6, expand=E, seed=11)
        src, use = expansion_pattern(6, E, 3, random.Random(11))
        assert np.array_equal(src, np.repeat(np.arange(6), E))
        for m in range(3):
            assert np.array_equal(use[:, m], fl[m].reshape(-1))
            assert np.array_equal(expand_on_host(xs[m][::E], src, use[:, m]), xs[m])
        assert (use.sum(1) >= 1).all()          # never all modalities missing


# ---- GaitSet oracle (row a16): pin the rank-3 pieces against literal restatements of the reference ops
def _sign_max_literal_np(xs):
    """mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:169-178 op by op (stack, reshape, argmax|.|, gather_nd)."""
    dims = xs[0].shape
    cat = np.stack(xs, 0).reshape(len(xs), -1)
    pos = np.argmax(np.abs(cat), axis=0)                       # first maximum, like tf.math.argmax
    return cat[pos, np.arange(dims[0] * dims[1] * dims[2])].reshape(dims)


def test_gaitset_sign_max_rank3_matches_literal():
    rng = np.random.default_rng(0)
    xs = [rng.standard_normal((62, 5, 8)) for _ in range(3)]
    xs[1][:, :, :2] = -xs[0][:, :, :2]                         # |.| ties -> lowest modality index
    got = O.merge_modalities([torch.tensor(x) for x in xs], O.MERGE_SIGNMAX).numpy()
    assert np.array_equal(got, _sign_max_literal_np(xs))


def test_gaitset_hpp_layout_and_shapes():
    from oracle import gaitset_oracle as G
    cfg = G.GaitSetConfig(in_channels=(1,), frames=2, hw=12, nclasses=0)
    P = G.init_params(cfg, dtype=torch.float64)
    x = torch.rand(2, 2, 12, 12, 1, dtype=torch.float64)
    out, acts = G.gaitset_branch_forward(x, P, "ofBranch", cfg, return_acts=True)
    assert out.shape == (62, 2, 256) and acts["hpp"].shape == (62, 2, 128)
    # literal Keras ops on the NHWC map: Reshape((nb,-1,c)) then mean+max over axis 2, a-strips before b-strips
    a = acts["a_set"].permute(0, 2, 3, 1).numpy()
    b = acts["b_set"].permute(0, 2, 3, 1).numpy()
    feats = []
    for nb in (1, 2, 4, 8, 16):
        for m in (a, b):
            r = m.reshape(m.shape[0], nb, -1, m.shape[-1])
            feats.append(r.mean(2) + r.max(2))
    lit = np.concatenate(feats, 1).transpose(1, 0, 2)
    assert np.allclose(acts["hpp"].numpy(), lit, rtol=0, atol=1e-12)
    # axis=1 of [62,B,d] is the batch axis: every (part, feature) column of the signature has unit norm
    sig, _ = G.model_forward([x], [torch.ones(2, 1, dtype=torch.float64)], P, cfg)
    assert torch.allclose((sig ** 2).sum(1), torch.ones(62, 256, dtype=torch.float64))


def test_optimizer_variants_against_torch_optim():
    """Adam, AMSGrad, decoupled weight decay (tfa AdamW) and SGD with momentum of the oracle against torch.optim -- an
    implementation the restatement shares no code with.  torch places epsilon after the bias correction of sqrt(v)
    (Kingma & Ba, algorithm 1) where Keras uses the "epsilon hat" form (the note before their section 2.1): with
    |g| = O(1) and eps = 1e-7 the two differ by O(eps / sqrt(v)), the tolerance below."""
    g = torch.Generator().manual_seed(17)
    w0 = torch.randn(50, generator=g, dtype=torch.float64)
    grads = []
    for i in range(6):          # |g| >= 0.5: the eps-placement difference is lr * eps / sqrt(v) <= 1e-2 * 1e-7 / (0.03 * 0.5) per step
        r = torch.randn(50, generator=g, dtype=torch.float64) * (1.0 + 0.5 * (5 - i))    # shrinking: v decays, vhat holds
        grads.append(torch.sign(r) * (0.5 + r.abs()))
    lr, wd = 1e-2, 1e-3

    def run_torch(opt_cls, **kw):
        p = torch.nn.Parameter(w0.clone())
        opt = opt_cls([p], **kw)
        for gr in grads:
            p.grad = gr.clone()
            opt.step()
        return p.detach()

    def run_oracle(amsgrad=False, weight_decay=0.0, b2=0.999):
        P, G = {"w": w0.clone()}, None
        M, V = {"w": torch.zeros_like(w0)}, {"w": torch.zeros_like(w0)}
        Vh = {"w": torch.zeros_like(w0)} if amsgrad else None
        for t, gr in enumerate(grads, 1):
            O.adam_step(P, {"w": gr}, M, V, t, lr=lr, b2=b2, eps=1e-7, Vhat=Vh, weight_decay=weight_decay)
        return P["w"]
    tol = dict(rtol=0, atol=1e-6)
    assert torch.allclose(run_oracle(), run_torch(torch.optim.Adam, lr=lr, eps=1e-7), **tol)
    # (beta_2 = 0.5: with 0.999 v only grows over six steps and vhat never differs from it)
    assert torch.allclose(run_oracle(amsgrad=True, b2=0.5),
                          run_torch(torch.optim.Adam, lr=lr, betas=(0.9, 0.5), eps=1e-7, amsgrad=True), **tol)
    assert torch.allclose(run_oracle(b2=0.5), run_torch(torch.optim.Adam, lr=lr, betas=(0.9, 0.5), eps=1e-7), **tol)
    # tfa AdamW subtracts weight_decay * var (NOT scaled by lr); torch scales its coefficient by lr
    assert torch.allclose(run_oracle(weight_decay=wd), run_torch(torch.optim.AdamW, lr=lr, eps=1e-7, weight_decay=wd / lr), **tol)
    assert not torch.allclose(run_oracle(b2=0.5), run_oracle(amsgrad=True, b2=0.5), rtol=0, atol=1e-5)   # (the variants do differ)
    # SGD(momentum): Keras keeps v = mu v - lr g, torch v = mu v + g and steps by -lr v: identical for a constant lr
    P, V = {"w": w0.clone()}, {"w": torch.zeros_like(w0)}
    for t, gr in enumerate(grads, 1):
        O.sgd_step(P, {"w": gr}, V, t, lr=lr, momentum=0.9)
    assert torch.allclose(P["w"], run_torch(torch.optim.SGD, lr=lr, momentum=0.9), rtol=0, atol=1e-14)
    # ... and Keras' `decay`: lr_t = lr / (1 + decay * iterations), baked into the velocity at the step it is applied
    P, V, wn, vn = {"w": w0.clone()}, {"w": torch.zeros_like(w0)}, w0.numpy().copy(), np.zeros(50)
    for t, gr in enumerate(grads, 1):
        O.sgd_step(P, {"w": gr}, V, t, lr=lr, momentum=0.9, decay=0.1)
        vn = 0.9 * vn - lr / (1 + 0.1 * (t - 1)) * gr.numpy()
        wn = wn + vn
    assert np.allclose(P["w"].numpy(), wn, rtol=0, atol=1e-14)


def test_conv3d_branch_against_literal_channels_last_numpy():
    """build_3Dbranch / build_3DbranchLReLU (nets/mj_uwyhNets_ba.py:346-372, :385-417) restated in channels-last numpy:
    six strided 'valid' Conv3D as sums over strided sliding windows with Keras' (kt,kh,kw,cin,cout) kernels, ReLU |
    LeakyReLU(alpha), Conv3D(nd, (1,1,1)) "grayCode", Flatten -- equal to the torch oracle's branch3d_forward."""
    from numpy.lib.stride_tricks import sliding_window_view
    layers = (((3, 5, 5), (1, 2, 2)), ((3, 3, 3), (1, 2, 2)), ((3, 3, 3), (2, 2, 2)), ((3, 3, 3), (2, 2, 2)),
              ((3, 2, 2), (1, 1, 1)), ((2, 1, 1), (1, 1, 1)))                       # the literal kernel / stride list
    for act, alpha in ((O.ACT_RELU, 0.0), (O.ACT_LEAKY, 0.2)):
        cfg = O.NetConfig(in_channels=(25,), nd=6, nclasses=0, single=True, act=act, alpha=alpha, branch3d=(True,),
                          filters3d=(3, 4, 5, 6, 6, 7))
        P = O.init_params(cfg, seed=8, dtype=torch.float64)
        g = torch.Generator().manual_seed(8)
        for k in P:
            if k.endswith("/b"):
                P[k] = torch.randn(P[k].shape, generator=g, dtype=torch.float64) * 0.1
        x = torch.rand(2, 25, 60, 60, 1, generator=g, dtype=torch.float64) - 0.5
        got = O.branch3d_forward(x, P, "ofBranch", cfg).numpy()
        a = x.numpy()                                                                # [B, T, H, W, C] channels-last
        for li, (ks, st) in enumerate(layers):
            k = np.transpose(P[f"ofBranch/conv{li}/w"].numpy(), (2, 3, 4, 1, 0))     # [cout,cin,kt,kh,kw] -> (kt,kh,kw,cin,cout)
            assert k.shape[:3] == ks
            win = sliding_window_view(a, ks, axis=(1, 2, 3))[:, ::st[0], ::st[1], ::st[2]]   # [B,t,h,w,C,kt,kh,kw]
            z = np.einsum("bthwcijk,ijkco->bthwo", win, k) + P[f"ofBranch/conv{li}/b"].numpy()
            a = np.maximum(z, 0) if act == O.ACT_RELU else np.where(z > 0, z, alpha * z)
        assert a.shape[1:4] == (1, 1, 1)
        lit = a.reshape(2, -1) @ P["ofBranch/ofCode/w"].numpy().T + P["ofBranch/ofCode/b"].numpy()
        assert got.shape == lit.shape == (2, 6) and np.abs(got - lit).max() <= 1e-12 * max(1.0, np.abs(lit).max())


def _gs_branch_literal_np(x, P, bn, alpha=0.3):
    """build_gaitset_branch (nets/mj_uwyhNets_ba.py:427-482) layer by layer in numpy, CHANNELS-LAST as Keras runs it, with no
    operator shared with the torch oracle: ZeroPadding2D(2), Conv2D(k, 'same', no bias) as a sum over sliding windows with
    the kernel in Keras' (kh, kw, cin, cout) layout, LeakyReLU(), MaxPooling2D(2), reduce_max over the frame axis, Add,
    the five Reshape((num_bin, -1, c)) + mean + max strips of both maps, Concatenate(axis=1), transpose([1,0,2]), MatMul."""
    from numpy.lib.stride_tricks import sliding_window_view

    def conv_same(a, w):                                   # a [n,h,w,c]; w oracle layout [cout,cin,kh,kw] -> Keras (kh,kw,cin,cout)
        k = np.transpose(w, (2, 3, 1, 0))
        ph = k.shape[0] // 2
        ap = np.pad(a, ((0, 0), (ph, ph), (ph, ph), (0, 0)))
        win = sliding_window_view(ap, (k.shape[0], k.shape[1]), axis=(1, 2))       # [n,h,w,c,kh,kw]
        return np.einsum("nhwcij,ijco->nhwo", win, k)

    def lrelu(a):
        return np.where(a > 0, a, alpha * a)

    def pool(a):
        n, h, w, c = a.shape
        return a.reshape(n, h // 2, 2, w // 2, 2, c).max((2, 4))

    W = {k.split("/")[1]: v.numpy() for k, v in P.items() if k.startswith(bn + "/")}
    B, T = x.shape[:2]
    a = np.pad(x, ((0, 0), (0, 0), (2, 2), (2, 2), (0, 0))).reshape((B * T,) + (x.shape[2] + 4, x.shape[3] + 4, x.shape[4]))
    td_max = lambda t: t.reshape((B, T) + t.shape[1:]).max(1)                      # Lambda(reduce_max(axis=1))
    a = lrelu(conv_same(a, W["a1"]))
    a = pool(lrelu(conv_same(a, W["a2"])))
    b = td_max(a)
    b = lrelu(conv_same(b, W["b1"]))
    b = pool(lrelu(conv_same(b, W["b2"])))
    a = lrelu(conv_same(a, W["a3"]))
    a = pool(lrelu(conv_same(a, W["a4"])))
    b = b + td_max(a)
    b = lrelu(conv_same(b, W["b3"]))
    b = lrelu(conv_same(b, W["b4"]))
    a = lrelu(conv_same(a, W["a5"]))
    a = lrelu(conv_same(a, W["a6"]))
    a = td_max(a)
    b = b + a
    feats = []
    for nb in (1, 2, 4, 8, 16):
        for m in (a, b):
            r = m.reshape(m.shape[0], nb, -1, m.shape[-1])
            feats.append(r.mean(2) + r.max(2))
    f = np.transpose(np.concatenate(feats, 1), (1, 0, 2))                         # [62, B, 128]
    return np.matmul(f, W["matmul"])


def test_gaitset_branch_against_literal_channels_last_numpy():
    """The whole GaitSet branch of the torch oracle == the layer-by-layer channels-last numpy restatement (fp64)."""
    from oracle import gaitset_oracle as G
    for c, seed in ((1, 2), (2, 5)):
        cfg = G.GaitSetConfig(in_channels=(c,), frames=3, hw=12, nclasses=0)
        P = G.init_params(cfg, seed=seed, dtype=torch.float64)
        xs, _, _ = G.synth_batch(cfg, ids=2, per_id=1, seed=seed, dtype=torch.float64)
        got = G.gaitset_branch_forward(xs[0], P, "ofBranch", cfg).numpy()
        lit = _gs_branch_literal_np(xs[0].numpy(), P, "ofBranch")
        assert got.shape == lit.shape == (62, 2, 256)
        assert np.abs(got - lit).max() <= 1e-12 * max(1.0, np.abs(lit).max())
        # gradients: torch autograd of the oracle == central differences of the NUMPY forward (a linear functional of the
        # branch output, so the only non-smooth points are LeakyReLU / max kinks, measure-zero for a 1e-6 step)
        R = torch.randn(got.shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
        Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        (G.gaitset_branch_forward(xs[0], Pg, "ofBranch", cfg) * R).sum().backward()
        for name in ("ofBranch/a1/w", "ofBranch/b3/w", "ofBranch/a6/w", "ofBranch/matmul/w"):
            g = Pg[name].grad
            d = g / g.norm()
            eps = 1e-6
            Pp, Pm = dict(P), dict(P)
            Pp[name], Pm[name] = P[name] + eps * d, P[name] - eps * d
            fd = float(((_gs_branch_literal_np(xs[0].numpy(), Pp, "ofBranch") - _gs_branch_literal_np(xs[0].numpy(), Pm, "ofBranch"))
                        * R.numpy()).sum()) / (2 * eps)
            assert abs(fd - float(g.norm())) <= 1e-6 * float(g.norm()), (name, fd, float(g.norm()))


def test_gaitset_single_modality_graph_is_the_bare_branch():
    """UWYHSemiNet.build on ONE input shape with gaitset (nets/mj_uwyhNets_ba.py:890-905): ``ofout1 = ofBranch;
    outsignature = ofout1`` -- no use-flag gate, no fusion, no l2_normalize; "classprob" on transpose([1,0,2]) + Flatten;
    losses = [triplet over the 62 parts, plain categorical cross-entropy].  Gradients against central differences."""
    from oracle import gaitset_oracle as G
    cfg = G.GaitSetConfig(in_channels=(1,), frames=2, hw=12, nc=0, nclasses=5, wver=1.0, wid=0.1, single=True)
    P = G.init_params(cfg, seed=3, dtype=torch.float64)
    xs, fl, lab = G.synth_batch(cfg, ids=2, per_id=2, seed=3, dtype=torch.float64)
    sig, logits = G.model_forward(xs, None, P, cfg)
    branch = G.gaitset_branch_forward(xs[0], P, "ofBranch", cfg)
    assert torch.equal(sig, branch) and sig.shape == (62, 4, 256)
    assert not torch.allclose((sig ** 2).sum(1), torch.ones(62, 256, dtype=torch.float64))      # NOT normalised
    lit = torch.tensor(np.transpose(sig.numpy(), (1, 0, 2)).reshape(4, -1)) @ P["classprob/w"].t() + P["classprob/b"]
    assert torch.allclose(logits, lit, rtol=0, atol=1e-12)
    res, grads = G.loss_and_grads(xs, fl, lab, P, cfg)
    trip, cnt = O.triplet_loss_all(lab, sig, cfg.margin)
    ce = -torch.log_softmax(logits, 1)[torch.arange(4), lab].mean()
    assert abs(float(res["loss"]) - float(trip + 0.1 * ce)) < 1e-12 and float(res["reg"]) == 0.0
    # central differences on a few coordinates of the classifier and the per-part MatMul (smooth everywhere); the conv
    # kernels are checked through a directional derivative of the triplet + CE loss
    for name in ("classprob/w", "ofBranch/matmul/w", "ofBranch/a6/w", "ofBranch/a1/w"):
        d = grads[name] / grads[name].norm()            # steepest direction: the derivative is |grad|, not ~0
        eps = 1e-5      # the a == p diagonal distances are sqrt(rounding residue) ~ 1e-8: a noise floor of ~3e-11 on the loss
        Pp, Pm = dict(P), dict(P)
        Pp[name], Pm[name] = P[name] + eps * d, P[name] - eps * d
        fd = (float(G.total_loss(xs, fl, lab, Pp, cfg)["loss"]) - float(G.total_loss(xs, fl, lab, Pm, cfg)["loss"])) / (2 * eps)
        an = float((grads[name] * d).sum())
        assert an > 0 and abs(fd - an) <= 1e-3 * an, (name, fd, an)


def test_gaitset_postriplet2_graph_literal():
    """postriplet == 2 with GaitSet branches (nets/mj_uwyhNets_ba.py:814-832, LeakyReLU path): fusion NOT normalised ->
    Dense(nc, None, activity_regularizer l2(1e-3)) "signature" -> LeakyReLU -> l2_normalize(axis=1) "code" (outsignature) ->
    transpose + Flatten -> "classprob"; against the literal numpy statements."""
    from oracle import gaitset_oracle as G
    cfg = G.GaitSetConfig(in_channels=(2, 1), frames=2, hw=12, nc=8, nclasses=5, merge=O.MERGE_SIGNMAX, wver=1.0, wid=0.5,
                          postriplet=2)
    P = G.init_params(cfg, seed=4, dtype=torch.float64)
    xs, fl, lab = G.synth_batch(cfg, ids=2, per_id=2, seed=4, dtype=torch.float64)
    outs = G.model_forward(xs, fl, P, cfg, return_all=True)
    fus = outs["fusion"].numpy()                                                     # [62, B, 256], un-normalised
    assert not np.allclose((fus ** 2).sum(1), 1.0)
    lin = fus @ P["code/w"].numpy().T + P["code/b"].numpy()
    act = np.where(lin > 0, lin, cfg.alpha * lin)
    code = act / np.sqrt(np.maximum((act ** 2).sum(1, keepdims=True), 1e-12))        # axis 1 = the batch axis of [62, B, nc]
    assert np.allclose(outs["signature_layer"].numpy(), lin, rtol=0, atol=1e-12)
    assert np.allclose(outs["signature"].numpy(), code, rtol=0, atol=1e-12) and outs["code"] is outs["signature"]
    logits = np.transpose(code, (1, 0, 2)).reshape(4, -1) @ P["classprob/w"].numpy().T + P["classprob/b"].numpy()
    assert np.allclose(outs["logits"].numpy(), logits, rtol=0, atol=1e-12)
    res, grads = G.loss_and_grads(xs, fl, lab, P, cfg)
    trip, _ = O.triplet_loss_all(lab, torch.tensor(code), cfg.margin)
    ce = -torch.log_softmax(torch.tensor(logits), 1)[torch.arange(4), lab].mean()
    reg = 1e-3 * (lin ** 2).sum() / 62                                               # Keras: / shape(output)[0]
    assert abs(float(res["loss"]) - float(trip + 0.5 * ce + reg)) < 1e-9 and abs(float(res["reg"]) - reg) < 1e-15   # (a == p distances: sqrt of rounding residue)
    # (not code/b: lin ~ 1e-4 sits on the LeakyReLU kink; not the conv kernels: sign_max winners flip -- a discontinuity)
    for name in ("code/w", "classprob/w", "ofBranch/matmul/w"):
        d = grads[name] / grads[name].norm()
        eps = 1e-5
        Pp, Pm = dict(P), dict(P)
        Pp[name], Pm[name] = P[name] + eps * d, P[name] - eps * d
        fd = (float(G.total_loss(xs, fl, lab, Pp, cfg)["loss"]) - float(G.total_loss(xs, fl, lab, Pm, cfg)["loss"])) / (2 * eps)
        an = float((grads[name] * d).sum())
        assert an > 0 and abs(fd - an) <= 1e-3 * an, (name, fd, an)


# ---- committed step fixtures: the restatement must not drift
@pytest.mark.parametrize("name", ["step_stacked", "step_gaitset", "step_gaitset_single", "step_gaitset_post2"])
def test_oracle_reproduces_step_fixture(golden_dir, name):
    sys.path.insert(0, golden_dir)
    import make_golden
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    got = make_golden.step_case(name)
    for k in ("triplet", "ce", "count", "reg", "loss"):
        assert float(got[k]) == pytest.approx(float(z[k]), rel=1e-12), k
    assert np.allclose(got["signature"], z["signature"], rtol=0, atol=1e-12)
    assert list(got["names"]) == list(z["names"])
    assert np.allclose(got["grad_norms"], z["grad_norms"], rtol=1e-10)


# ---- second, independent restatement (oracle/numpy_restatement.py): numpy, hand-derived backward -------------------
def _np_case(name):
    from oracle import numpy_restatement as NR
    if name == "3mod_signmax":
        oc = O.NetConfig(in_channels=(3, 2, 2), filters_numbers=(4, 4, 6, 6), nd=8, nclasses=5, merge=O.MERGE_SIGNMAX,
                         wver=1.0, wid=0.1)
        sb = dict(base_rows=4, expand=4)
    elif name == "2mod_max_leaky_code":
        oc = O.NetConfig(in_channels=(3, 2), filters_numbers=(4, 4, 6, 6), nd=8, nc=4, nclasses=5, merge=O.MERGE_MAX,
                         act=O.ACT_LEAKY, wver=0.7, wid=1.0, label_smoothing=0.1)
        sb = dict(base_rows=6, expand=2, kinds=("of", "gray"))
    elif name == "3mod_avg_relu_code":
        oc = O.NetConfig(in_channels=(2, 2, 2), filters_numbers=(4, 4, 6, 6), nd=8, nc=4, nclasses=4, merge=O.MERGE_AVG,
                         wver=0.5, wid=0.5)
        sb = dict(base_rows=4, expand=3, kinds=("of", "gray", "sil"))
    else:
        oc = O.NetConfig(in_channels=(3,), filters_numbers=(4, 4, 6, 6), nd=8, nclasses=6, single=True, wver=1.0, wid=0.1)
        sb = dict(base_rows=8, expand=1, kinds=("gray",))
    xs, fl, lab = O.synth_batch(oc, seed=21, dtype=np.float64, **sb)
    lab = lab % oc.nclasses
    P = O.init_params(oc, seed=21, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    for k in P:
        if k.endswith("/b"):
            P[k] = torch.randn(P[k].shape, generator=g, dtype=torch.float64) * 0.05
    B = xs[0].shape[0]
    masks = [(torch.rand(B, 2 * oc.nd, generator=g) >= 0.3).double() / 0.7 for _ in range(oc.nmods)]
    cmask = (torch.rand(B, oc.nc, generator=g) >= 0.3).double() / 0.7 if oc.nc else None
    return NR, oc, xs, fl, lab, P, masks, cmask


@pytest.mark.parametrize("name", ["3mod_signmax", "2mod_max_leaky_code", "3mod_avg_relu_code", "1mod_gray"])
def test_torch_oracle_equals_independent_numpy_restatement(name):
    """conv / pool / gate / fusion / l2_normalize / FC1 / FC2+CE / triplet / regularisers: the autograd oracle and the
    hand-derived numpy restatement agree on every loss (1e-11) and every gradient tensor (1e-9)."""
    NR, oc, xs, fl, lab, P, masks, cmask = _np_case(name)
    res, G = O.loss_and_grads([torch.tensor(x) for x in xs], [torch.tensor(f) for f in fl], torch.tensor(lab), P, oc,
                              masks, cmask)
    Pn = {k: v.numpy() for k, v in P.items()}
    rn, Gn = NR.step(xs, fl, lab, Pn, oc, [m.numpy() for m in masks], None if cmask is None else cmask.numpy())
    # 1mod_gray: the signature is NOT normalised there (|x|^2 ~ 10), and batch_dist forms d^2 by cancellation
    # (x2a + x2b - 2ab): two summation orders differ by ~1e-16 |x|^2 / d^2 -> 1e-9 relative
    ltol, gtol = (5e-9, 1e-7) if oc.single else (1e-11, 1e-9)
    for key in ("triplet", "ce", "reg", "loss", "acc"):
        assert float(rn[key]) == pytest.approx(float(res[key]), rel=ltol, abs=1e-13), key
    assert float(rn["count"]) == float(res["count"].sum())
    assert np.allclose(rn["signature"], res["signature"].numpy(), rtol=1e-11, atol=1e-13)
    assert set(Gn) == set(G)
    for k in G:
        ref = G[k].numpy()
        assert np.abs(Gn[k] - ref).max() <= gtol * max(np.abs(ref).max(), 1e-6), k


def test_numpy_restatement_gradients_match_finite_differences():
    """The hand-derived backward against central differences of its own forward (a few coordinates per tensor)."""
    NR, oc, xs, fl, lab, P, masks, cmask = _np_case("2mod_max_leaky_code")
    Pn = {k: v.numpy().copy() for k, v in P.items()}
    mk = [m.numpy() for m in masks]
    _, Gn = NR.step(xs, fl, lab, Pn, oc, mk, cmask.numpy())
    rng = np.random.default_rng(0)
    h = 1e-7        # the graph is piecewise linear in places (LeakyReLU / pool / hinge kinks): a small step keeps both
    for k in Pn:    # probes on one piece (measured: 1e-6 already straddles a kink for one conv0 weight); fp64 noise ~1e-9
        flat = Pn[k].reshape(-1)
        for j in rng.choice(flat.size, size=min(3, flat.size), replace=False):
            w0 = flat[j]
            flat[j] = w0 + h
            lp = NR.step(xs, fl, lab, Pn, oc, mk, cmask.numpy(), want_grads=False)[0]["loss"]
            flat[j] = w0 - h
            lm = NR.step(xs, fl, lab, Pn, oc, mk, cmask.numpy(), want_grads=False)[0]["loss"]
            flat[j] = w0
            fd = (lp - lm) / (2 * h)
            an = Gn[k].reshape(-1)[j]
            assert abs(fd - an) <= 2e-6 * max(1.0, abs(an)) + 1e-8, (k, j, fd, an)


def test_adam_update_equals_independent_numpy_restatement():
    NR, oc, xs, fl, lab, P, masks, cmask = _np_case("3mod_signmax")
    Pt = {k: v.clone() for k, v in P.items()}
    Mt = {k: torch.zeros_like(v) for k, v in P.items()}
    Vt = {k: torch.zeros_like(v) for k, v in P.items()}
    Pn = {k: v.numpy().copy() for k, v in P.items()}
    Mn = {k: np.zeros_like(v) for k, v in Pn.items()}
    Vn = {k: np.zeros_like(v) for k, v in Pn.items()}
    mk = [m.numpy() for m in masks]
    for t in (1, 2, 3):
        _, G = O.loss_and_grads([torch.tensor(x) for x in xs], [torch.tensor(f) for f in fl], torch.tensor(lab), Pt, oc, masks)
        O.adam_step(Pt, G, Mt, Vt, t, lr=1e-3)
        _, Gn = NR.step(xs, fl, lab, Pn, oc, mk)
        NR.adam_update(Pn, Gn, Mn, Vn, t, lr=1e-3)
    for k in Pn:
        assert np.abs(Pn[k] - Pt[k].numpy()).max() <= 1e-9, k


def test_decision_injection_reproduces_free_run():
    """oracle.loss_and_grads(decisions=its own recorded decisions) == the free-running oracle (the machinery the GPU
    gradient-parity tests rely on, tests/test_decisions_gpu.py)."""
    NR, oc, xs, fl, lab, P, masks, cmask = _np_case("3mod_signmax")
    args = ([torch.tensor(x) for x in xs], [torch.tensor(f) for f in fl], torch.tensor(lab), P, oc, masks)
    rec = {}
    r0, G0 = O.loss_and_grads(*args)
    r1, G1 = O.loss_and_grads(*args, record=rec)
    dec = {m: {k: v for k, v in rec[m].items() if k.startswith(("pool", "act"))} for m in range(oc.nmods)}
    dec["winner"] = rec["winner"]
    r2, G2 = O.loss_and_grads(*args, decisions=dec)
    for k in G0:
        assert torch.equal(G0[k], G1[k]) and torch.allclose(G0[k], G2[k], rtol=0, atol=1e-15), k
    # a deliberately flipped arg-max reroutes the gradient of the tensors upstream of it
    dec[0]["pool1"] = (dec[0]["pool1"] + 1) % 4
    _, G3 = O.loss_and_grads(*args, decisions=dec)
    assert not torch.allclose(G3["ofBranch/conv0/w"], G0["ofBranch/conv0/w"])


def test_oracle_primitives_against_scipy_and_sklearn():
    """Third-party pins of the restated primitives (TensorFlow itself is unavailable): valid cross-correlation against
    scipy.signal.correlate, 2x2/2 max-pooling against scipy.ndimage.maximum_filter, categorical cross-entropy against
    sklearn.metrics.log_loss, l2_normalize against sklearn.preprocessing.normalize, batch_dist against
    scipy.spatial.distance.cdist, Adam against a scalar closed form of Kingma & Ba's update as Keras implements it."""
    import math
    import torch.nn.functional as F
    from scipy import ndimage, signal
    from scipy.spatial.distance import cdist
    from sklearn.metrics import log_loss
    from sklearn.preprocessing import normalize
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 3, 12, 12))
    w = rng.normal(size=(4, 3, 5, 5))
    b = rng.normal(size=4)
    y = F.conv2d(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    ref = np.stack([np.stack([sum(signal.correlate(x[n, c], w[o, c], mode="valid") for c in range(3)) + b[o]
                              for o in range(4)]) for n in range(2)])
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12)
    a = rng.normal(size=(2, 3, 9, 9))                        # odd size: the trailing row / column is dropped
    pooled = F.max_pool2d(torch.tensor(a), 2).numpy()
    mf = ndimage.maximum_filter(a, size=(1, 1, 2, 2), origin=(0, 0, -1, -1))[:, :, 0:8:2, 0:8:2]
    assert np.array_equal(pooled, mf)
    logits = rng.normal(size=(11, 7))
    lab = rng.integers(0, 7, 11)
    ce, _ = O.softmax_ce(torch.tensor(logits), F.one_hot(torch.tensor(lab), 7).double())
    p = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    assert float(ce) == pytest.approx(log_loss(lab, p, labels=list(range(7))), rel=1e-12)
    # smoothlabels (:1252-1262): tf.losses.CategoricalCrossentropy(label_smoothing=e) = targets y (1-e) + e/C -- the
    # definition torch's cross_entropy(label_smoothing=e) implements
    oh = F.one_hot(torch.tensor(lab), 7).double()
    ce_s, _ = O.softmax_ce(torch.tensor(logits), oh * 0.9 + 0.1 / 7)
    assert float(ce_s) == pytest.approx(float(F.cross_entropy(torch.tensor(logits), torch.tensor(lab), label_smoothing=0.1)),
                                        rel=1e-12)
    v = rng.normal(size=(6, 10))
    assert np.allclose(O.l2_normalize(torch.tensor(v), 1).numpy(), normalize(v, norm="l2", axis=1), rtol=1e-12)
    e = rng.normal(size=(1, 9, 5))
    assert np.allclose(O.batch_dist(torch.tensor(e))[0].numpy(), cdist(e[0], e[0]) * (1 - np.eye(9)), atol=1e-7)
    # Adam, three steps on one scalar with a constant gradient g: m_t = (1-b1^t) g, v_t = (1-b2^t) g^2 ->
    # every update is -lr * g / (|g| + eps * sqrt(1-b2^t)) exactly (Keras folds the bias correction into lr_t)
    P, G = {"w": torch.tensor([1.0], dtype=torch.float64)}, {"w": torch.tensor([0.3], dtype=torch.float64)}
    M, V = {"w": torch.zeros(1, dtype=torch.float64)}, {"w": torch.zeros(1, dtype=torch.float64)}
    wref = 1.0
    for t in (1, 2, 3):
        O.adam_step(P, G, M, V, t, lr=1e-2)
        mt, vt = (1 - 0.9 ** t) * 0.3, (1 - 0.999 ** t) * 0.09
        wref -= 1e-2 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * mt / (math.sqrt(vt) + 1e-7)
        assert float(P["w"]) == pytest.approx(wref, rel=1e-12)


def _hard_triplet_plain(lab, e, margin):
    """Batch-hard loss written directly from its definition (loops, fp64): farthest positive, nearest negative."""
    B = len(lab)
    d = np.sqrt(np.maximum(((e[:, None, :] - e[None, :, :]) ** 2).sum(2), 0.0))
    tot, act = 0.0, 0
    for a in range(B):
        pos = [d[a, b] for b in range(B) if b != a and lab[b] == lab[a]]
        neg = [d[a, b] for b in range(B) if lab[b] != lab[a]]
        hp = max(pos) if pos else 0.0
        hn = min(neg) if neg else d[a].max()          # masked_minimum without negatives leaves the row maximum
        t = hp - hn + margin
        if t > 0:
            tot += t
            act += 1
    return tot / B, act


def test_triplet_hard_oracle_against_plain_definition_and_fd():
    """compile_hard's tfa TripletHardLoss (restated masked_maximum / masked_minimum form) against the loss written from
    its definition, incl. an anchor without positives, a batch without negatives, duplicated rows; gradient by finite
    differences."""
    rng = np.random.default_rng(11)
    for lab, B in ((np.repeat(np.arange(6), 4), 24), (np.array([0, 0, 0, 1, 2, 2, 3]), 7), (np.zeros(5, dtype=int), 5)):
        e = rng.normal(size=(B, 16))
        e[1] = e[0]
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        for margin in (0.2, 1.0):
            loss, act = O.triplet_hard_loss(torch.tensor(lab), torch.tensor(e), margin)
            want, wact = _hard_triplet_plain(lab, e, margin)
            assert float(loss) == pytest.approx(want, rel=1e-12, abs=1e-15)
            assert int(act) == wact
    lab = np.repeat(np.arange(4), 3)
    e = rng.normal(size=(12, 8))
    e64 = torch.tensor(e, requires_grad=True)
    loss, _ = O.triplet_hard_loss(torch.tensor(lab), e64, 0.5)
    loss.backward()
    h = 1e-6
    for (i, j) in ((0, 0), (3, 5), (11, 7)):
        ep, em = e.copy(), e.copy()
        ep[i, j] += h
        em[i, j] -= h
        fd = (_hard_triplet_plain(lab, ep, 0.5)[0] - _hard_triplet_plain(lab, em, 0.5)[0]) / (2 * h)
        assert float(e64.grad[i, j]) == pytest.approx(fd, rel=1e-5, abs=1e-9)
