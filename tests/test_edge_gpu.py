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


def test_decode_samples_bit_exact_vs_generator_arithmetic():
    """ugn_decode_samples against samples.decode_sample (= __load_dd, data/mj_dataGeneratorMMUWYHsingle.py:313-329):
    every int16 value with and without the raw-magnitude clip, every uint8 value for gray and silhouette."""
    from ugaitnet_b200 import ops, samples
    from ugaitnet_b200._ffi import TRef, check, lib, stream_ptr
    ctx = ops.get_ctx(0)
    i16 = np.arange(-32768, 32768, dtype=np.int16).reshape(64, 64, 16)
    u8 = np.arange(256, dtype=np.uint8).repeat(5)[:1279].reshape(1, 1279, 1)          # odd length: scalar tail
    for data, cf, sil, spec, clip in ((i16, 100, False, samples.RAW_FLOW, (0.0, 0.0)),
                                      (i16, 100, False, samples.RAW_FLOW, (50.0, 2300.0)),
                                      (u8, 1, False, samples.RAW_GRAY, (0.0, 0.0)),
                                      (u8, 1, True, samples.RAW_SILHOUETTE, (0.0, 0.0))):
        want = samples.decode_sample({"data": data, "compressFactor": cf}, silhouette=sil, ntype=2, clip_min=clip[0],
                                     clip_max=clip[1])
        raw = torch.tensor(np.ascontiguousarray(np.moveaxis(data, 2, 0))).cuda()
        out = torch.empty(raw.shape, device="cuda")
        a, b = TRef(raw), TRef(out)
        check(lib.ugn_decode_samples(ctx.h, a.ptr, spec[1], spec[2], spec[3], clip[0], clip[1], b.ptr, stream_ptr()))
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("mode", ["fp32", "f16mix"])
def test_raw_sample_batches_equal_decoded_batches(mode):
    """HostBatch(raw=...): stored int16 / uint8 volumes cross PCIe and are decoded on the device -- descriptors and step
    losses are bit-identical to feeding the host-decoded f32 volumes, plain and with the device-side expansion."""
    from ugaitnet_b200 import samples
    from ugaitnet_b200.config import NetConfig
    from ugaitnet_b200.net import UGaitEngine
    cfg = NetConfig(in_channels=(50, 25, 25), filters_numbers=(32, 32, 32, 64), nd=64, nclasses=10, merge=O.MERGE_SIGNMAX)
    eng = UGaitEngine(cfg, math_mode=mode, use_graph=False)
    rng = np.random.default_rng(4)
    B, B0 = 8, 4
    specs = [samples.RAW_FLOW, samples.RAW_GRAY, samples.RAW_SILHOUETTE]
    raws = [rng.integers(-3000, 3000, size=(B, 50, 60, 60)).astype(np.int16),
            rng.integers(0, 256, size=(B, 25, 60, 60)).astype(np.uint8),
            (rng.random((B, 25, 60, 60)) > 0.7).astype(np.uint8) * 255]
    dec = [(np.float32(r) / np.float32(s[1]) * np.float32(s[2]) - np.float32(s[3])).astype(np.float32)
           for r, s in zip(raws, specs)]
    assert int(raws[0].nbytes + raws[1].nbytes + raws[2].nbytes) * 8 == int(sum(d.nbytes for d in dec)) * 3
    # descriptor extraction
    hr, hf = eng.host_batch(B, train=False, raw=specs), eng.host_batch(B, train=False)
    assert hr.nbytes < 0.4 * hf.nbytes and hr.inputs[0].dtype == np.int16 and hr.inputs[1].dtype == np.uint8
    for m in range(3):
        hr.inputs[m][...] = raws[m]
        hf.inputs[m][...] = dec[m]
        hr.flags[m][...] = hf.flags[m][...] = (rng.random((B, 1)) > 0.3)
        hf.flags[m][...] = hr.flags[m]
    eng.prefetch_batch(hf, train=False)
    sig_f = eng.predict_prefetched("signature")
    eng.prefetch_batch(hr, train=False)
    sig_r = eng.predict_prefetched("signature")
    for m in range(3):      # the decoded volumes in the plan's input block are the generator's f32 values, bit for bit
        assert np.array_equal(eng.plan(B, False).br[m].x_in.cpu().numpy().view(np.uint32), dec[m].view(np.uint32))
    # (two runs of the fp32 validation kernels differ by their atomic accumulation order)
    assert float((sig_f - sig_r).norm() / sig_f.norm()) < 1e-6
    # training step on base rows expanded on the device
    lab = np.repeat(np.arange(B0 // 2), 2).astype(np.int32)
    losses = []
    for raw in (None, specs):
        eng2 = UGaitEngine(cfg, math_mode=mode, use_graph=False, seed=3)
        hb = eng2.host_batch(B, base_rows=B0, raw=raw)
        for m in range(3):
            hb.inputs[m][...] = (raws if raw else dec)[m][:B0]
            hb.flags[m][...] = 1.0
            hb.flags[m][1::2] = float(m == 1)
        hb.labels[...] = np.repeat(lab, B // B0)
        hb.src_row[...] = np.repeat(np.arange(B0, dtype=np.int32), B // B0)
        eng2.prefetch_batch(hb)
        out = eng2.train_step_prefetched()
        losses.append(out["losses"].cpu().clone())
    assert torch.allclose(losses[0][:4], losses[1][:4], rtol=1e-5, atol=1e-7) and float(losses[0][0]) > 0


def test_resident_input_buffers_equal_train_step():
    """input_buffers() + train_step_resident() / predict_resident(): the same step as train_step(tensors), with the batch
    written straight into the engine's input block."""
    from ugaitnet_b200.config import NetConfig
    from ugaitnet_b200.net import UGaitEngine
    cfg = NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nclasses=10, merge=O.MERGE_SIGNMAX)
    g = torch.Generator().manual_seed(1)
    B = 6
    xs = [torch.randn(B, c, 60, 60, generator=g).cuda() for c in cfg.in_channels]
    fl = [(torch.rand(B, 1, generator=g) > 0.2).float().cuda() for _ in cfg.in_channels]
    lab = torch.tensor([0, 0, 1, 1, 2, 2]).cuda()
    outs = []
    for resident in (False, True):
        eng = UGaitEngine(cfg, math_mode="fp32", lr=1e-3, seed=5, use_graph=False)
        if resident:
            xi, fi, li = eng.input_buffers(B)
            for m in range(3):
                xi[m].copy_(xs[m])
                fi[m].copy_(fl[m])
            li.copy_(lab.to(torch.int32))
            o = eng.train_step_resident(B)
            xp, fp, _ = eng.input_buffers(B, train=False)
            for m in range(3):
                xp[m].copy_(xs[m])
                fp[m].copy_(fl[m])
            sig = eng.predict_resident(B)
        else:
            o = eng.train_step(xs, fl, lab)
            sig = eng.predict(xs, fl)
        outs.append((o["losses"][:5].cpu().clone(), sig.cpu().clone(), eng.export_params()["ofBranch/dense/w"].cpu()))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-7)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-6)
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-7)


def test_piecewise_prefetch_delivers_the_bytes_of_the_single_copy():
    """prefetch_open / prefetch_upto / prefetch_close (the loader's cast of volume m + 1 overlaps the H2D copy of volume
    m): the staging block ends up byte-identical to the one prefetch_batch fills, for any cut points, and the step /
    the descriptors computed from it are those of the single copy."""
    from ugaitnet_b200.config import NetConfig
    from ugaitnet_b200.net import UGaitEngine
    cfg = NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nclasses=10, merge=O.MERGE_SIGNMAX)
    rng = np.random.default_rng(9)
    B = 6
    res = []
    for piecewise in (False, True):
        eng = UGaitEngine(cfg, math_mode="fp32", lr=1e-3, seed=5, use_graph=False)
        rng = np.random.default_rng(9)
        for train in (True, False):
            hb = eng.host_batch(B, train=train)
            for m in range(3):
                hb.inputs[m][...] = rng.standard_normal(hb.inputs[m].shape)
                hb.flags[m][...] = rng.random((B, 1)) > 0.2
            hb.labels[...] = [0, 0, 1, 1, 2, 2]
            if piecewise:
                eng.prefetch_open(hb, train)
                eng.prefetch_upto(100)                       # inside the header
                eng.prefetch_upto(hb.io.x_off[1])
                eng.prefetch_upto(hb.io.x_off[1])            # nothing new: no-op
                eng.prefetch_upto(hb.io.x_off[2] + 12345)    # mid-volume cut
                eng.prefetch_close()
            else:
                eng.prefetch_batch(hb, train)
            torch.cuda.synchronize()
            st = eng._io_stage[(B, train)]
            assert torch.equal(st["buf"][eng._io_k][:hb.nbytes].cpu(), hb.buf)
            if train:
                res.append(eng.train_step_prefetched()["losses"][:5].cpu().clone())
            else:
                res.append(eng.predict_prefetched("signature").cpu().clone())
    assert torch.allclose(res[0], res[2], rtol=1e-5, atol=1e-7) and float(res[0][0]) > 0
    assert torch.allclose(res[1], res[3], rtol=1e-5, atol=1e-6)
