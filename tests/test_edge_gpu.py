"""Edge cases at the drop-in boundary: the exceptions scikit-learn / Keras raise for the same mistakes, tiny and odd batch
sizes, out-of-range class ids, all-missing rows -- in the fp32 validation mode and in the benchmarked tensor-core mode."""
import numpy as np
import pytest
import torch

from oracle import ugait_oracle as O

pytestmark = pytest.mark.gpu


def test_knn_input_validation_matches_sklearn():
    from sklearn.neighbors import KNeighborsClassifier as SK
    from ugaitnet_b200.knn import KNeighborsClassifier
    rng = np.random.default_rng(0)
    G = rng.normal(size=(50, 24)).astype(np.float32)
    y = rng.integers(0, 5, 50)
    bad_q = np.full((2, 24), np.nan, np.float32)
    inf_q = np.full((2, 24), np.inf, np.float32)
    cases = [
        ("k > N", lambda K: K(n_neighbors=3).fit(G[:2], y[:2]).predict(G[:4])),
        ("Q = 0", lambda K: K(n_neighbors=3).fit(G, y).predict(G[:0])),
        ("feature mismatch", lambda K: K(n_neighbors=3).fit(G, y).predict(G[:3, :10])),
        ("NaN query", lambda K: K(n_neighbors=3).fit(G, y).predict(bad_q)),
        ("inf query", lambda K: K(n_neighbors=3).fit(G, y).predict(inf_q)),
        ("NaN gallery", lambda K: K(n_neighbors=3).fit(np.where(np.arange(50)[:, None] == 7, np.nan, G), y)),
        ("1-D query", lambda K: K(n_neighbors=3).fit(G, y).predict(G[0])),
        ("len(X) != len(y)", lambda K: K(n_neighbors=3).fit(G, y[:40])),
        ("empty gallery", lambda K: K(n_neighbors=3).fit(G[:0], y[:0])),
    ]
    for name, fn in cases:
        with pytest.raises(ValueError) as sk_err:
            fn(SK)
        with pytest.raises(ValueError) as our_err:
            fn(KNeighborsClassifier)
        # same first sentence for the messages scikit-learn words deterministically
        if name in ("k > N", "Q = 0", "feature mismatch", "NaN query"):
            assert str(our_err.value).split("\n")[0].strip() == str(sk_err.value).split("\n")[0].strip(), name
    # smallest legal problems: N == k, one query, one gallery row with k = 1
    for n, k in ((3, 3), (1, 1), (5, 1)):
        q = G[:4] + 0.01
        assert np.array_equal(KNeighborsClassifier(n_neighbors=k).fit(G[:n], y[:n]).predict(q),
                              SK(n_neighbors=k).fit(G[:n], y[:n]).predict(q))
    assert np.array_equal(KNeighborsClassifier(n_neighbors=3).fit(G, y).predict(G[:1] + 0.01),
                          SK(n_neighbors=3).fit(G, y).predict(G[:1] + 0.01))


@pytest.mark.parametrize("mode", ["fp32", "f16mix"])
def test_engine_tiny_and_odd_batches(mode):
    from ugaitnet_b200.config import NetConfig
    from ugaitnet_b200.net import UGaitEngine
    oc = O.NetConfig(in_channels=(6, 4, 4), filters_numbers=(32, 32, 32, 64), nd=64, nclasses=10, merge=O.MERGE_SIGNMAX,
                     wver=1.0, wid=0.1)
    cfg = NetConfig(in_channels=oc.in_channels, filters_numbers=oc.filters_numbers, nd=64, nclasses=10, merge=oc.merge,
                    wver=1.0, wid=0.1)
    eng = UGaitEngine(cfg, math_mode=mode, lr=1e-3)
    P = {k: v.double().cpu() for k, v in eng.export_params().items()}
    g = torch.Generator().manual_seed(3)
    tol = 1e-5 if mode == "fp32" else 1e-3
    for B, labs in ((1, [4]), (2, [1, 1]), (7, [0, 1, 2, 0, 1, 2, 0]), (3, [2, 99, -1])):
        xs = [torch.randn(B, c, 60, 60, generator=g) for c in cfg.in_channels]
        fl = [torch.ones(B, 1) for _ in cfg.in_channels]
        if B == 7:
            fl[1][2] = 0.0
            fl[0][5] = fl[1][5] = fl[2][5] = 0.0                      # a row with EVERY modality missing
            for m in range(3):
                xs[m][fl[m].reshape(-1) == 0] = 1e-9
        lab = torch.tensor(labs)
        out = eng.loss_and_grad([x.cuda() for x in xs], [f.cuda() for f in fl], lab.cuda())
        sig = eng.predict([x.cuda() for x in xs], [f.cuda() for f in fl])
        assert tuple(sig.shape) == (B, 64) and bool(torch.isfinite(sig).all())
        if B == 7:
            assert float(sig[5].abs().max()) == 0.0                   # all-zero fusion -> l2_normalize eps path
        x64, f64 = [x.double() for x in xs], [f.double() for f in fl]
        outs = O.model_forward(x64, f64, P, oc, return_all=True)
        trip, cnt = O.triplet_loss_all(lab, outs["signature"], oc.margin)
        assert float(out["triplet"]) == pytest.approx(float(trip), rel=tol, abs=1e-7)
        assert float(out["count"]) == float(cnt.sum())
        # class ids outside [0, C): tf.one_hot gives an all-zero target row -> no CE contribution, no out-of-row read
        valid = (lab >= 0) & (lab < 10)
        logp = torch.log_softmax(outs["logits"], dim=1)
        ce = -(logp[valid, lab[valid]]).sum() / B
        assert float(out["ce"]) == pytest.approx(float(ce), rel=tol, abs=1e-7)
        assert all(bool(torch.isfinite(v).all()) for v in eng.export_grads().values())


@pytest.mark.parametrize("mode", ["fp32", "f16mix"])
def test_engine_rejects_empty_and_misshaped_inputs(mode):
    from ugaitnet_b200.config import NetConfig
    from ugaitnet_b200.net import UGaitEngine
    cfg = NetConfig(in_channels=(6, 4, 4), filters_numbers=(32, 32, 32, 64), nd=64, nclasses=10, merge=O.MERGE_SIGNMAX)
    eng = UGaitEngine(cfg, math_mode=mode)
    mk = lambda B, hw=60: ([torch.randn(B, c, hw, 60, device="cuda") for c in cfg.in_channels],
                           [torch.ones(B, 1, device="cuda")] * 3)
    with pytest.raises(ValueError, match="non-empty"):
        eng.predict(*mk(0))
    with pytest.raises(ValueError, match="expected shape=\\(None, 6, 60, 60\\)"):
        eng.predict(*mk(2, 50))
    xs, fl = mk(2)
    with pytest.raises(ValueError, match="use-flag"):
        eng.predict(xs, [torch.ones(3, 1, device="cuda")] * 3)
    with pytest.raises(ValueError, match="Input 1"):
        eng.train_step([xs[0], xs[0], xs[2]], fl, torch.tensor([0, 1], device="cuda"))
    assert tuple(eng.predict(xs, fl).shape) == (2, 64)                # and the engine is still usable afterwards
    if mode != "fp32":      # the activation of layer i is the K operand of layer i + 1: multiples of 32 on the tensor cores
        with pytest.raises(ValueError, match="multiples of 32"):
            UGaitEngine(NetConfig(in_channels=(6, 4), filters_numbers=(16, 16, 32, 32), nd=64, nclasses=10), math_mode=mode)
