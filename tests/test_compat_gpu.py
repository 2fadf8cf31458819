"""The Keras-shaped drop-in surface (reference builder signatures, fit/predict/encode protocol)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ugait_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeGen:
    """Stands in for DataGeneratorGaitMMUWYH: len(), [i] -> (X_list, y_list), on_epoch_end()."""

    def __init__(self, oc, n=3, base_rows=4, expand=4, seed=0):
        self.items = []
        for i in range(n):
            xs, fl, lab = O.synth_batch(oc, base_rows=base_rows, expand=expand, seed=seed + i)
            lab = lab % oc.nclasses
            X = []
            for x, f in zip(xs, fl):
                X += [x.astype(np.float64), f.astype(np.float64)]       # the reference feeds float64
            onehot = np.eye(oc.nclasses)[lab.reshape(-1).astype(int)]
            self.items.append((X, [lab, onehot]))
        self.epochs_ended = 0

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    def on_epoch_end(self):
        self.epochs_ended += 1


@pytest.fixture()
def compat_path(monkeypatch):
    monkeypatch.setenv("UGN_MATH_MODE", "fp32")
    monkeypatch.syspath_prepend(os.path.join(ROOT, "ugaitnet_b200", "compat"))
    for m in [k for k in sys.modules if k == "nets" or k.startswith("nets.")]:
        del sys.modules[m]
    yield
    for m in [k for k in sys.modules if k == "nets" or k.startswith("nets.")]:
        del sys.modules[m]


def test_reference_import_paths_and_training_protocol(compat_path, tmp_path):
    # exactly the imports / calls of mains/mj_trainUWYHGaitNet_DataGen_3mods.py:339-343,559-572
    from nets.mj_uwyhNets_ba import UWYHSemiNet, UWYHSemiNet3Mods
    from nets.mj_metrics import mj_eerVerifDist
    from ugaitnet_b200.compat import optimizers, sign_max
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    input_shape = [(6, 60, 60), (4, 60, 60), (4, 60, 60)]
    model = UWYHSemiNet3Mods.build_or_load(input_shape, 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16],
                                           32, 0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3), margin=0.2,
                                           nclasses=10, loss_weights=[1.0, 0.1], initnet="", fMerge=sign_max)
    model.summary()
    oc = O.NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nclasses=10)
    gen, val = FakeGen(oc, n=3), FakeGen(oc, n=1, seed=50)
    model, hist = UWYHSemiNet.fit_generator(model, 4, [], gen, val, 0, len(gen), len(val))
    assert hist.epoch == [0, 1, 2, 3] and gen.epochs_ended == 4
    for key in ("loss", "classprob_acc", "val_classprob_acc", "lr", "val_loss"):
        assert key in hist.history and len(hist.history[key]) == 4
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    assert model.get_layer("classprob").units == 10 and model.dtype == "float32"
    # predict protocol + sub-model on a named layer (mains/mj_testUWYHGaitNet_open_tum.py:139-148)
    from ugaitnet_b200.compat import Model
    X, _ = gen[0]
    sig, prob = model.predict(X)
    assert sig.shape == (16, 32) and prob.shape == (16, 10)
    assert np.allclose(prob.sum(1), 1.0, atol=1e-5) and np.allclose(np.linalg.norm(sig, axis=1), 1.0, atol=1e-5)
    sub = Model(model.input, model.get_layer("signature").output)
    assert np.array_equal(sub.predict(X), sig)
    # encode(): first two modalities, always Maximum (nets/mj_uwyhNets_ba.py:971-999)
    codes = UWYHSemiNet.encode(model, [X[0], X[2]], [X[1], X[3]])
    P = {k: v.double().cpu() for k, v in model.engine.export_params().items()}
    b0 = O.branch_forward(torch.tensor(X[0]), P, "ofBranch", oc) * torch.tensor(X[1])
    b1 = O.branch_forward(torch.tensor(X[2]), P, "grayBranch", oc) * torch.tensor(X[3])
    ref = O.l2_normalize(torch.maximum(b0, b1), 1)
    assert np.allclose(codes, ref.numpy(), atol=2e-5)
    # weights round trip by layer name
    path = str(tmp_path / "model-state-0004_weights.hdf5")
    model.save_weights(path)
    m2 = UWYHSemiNet3Mods.build(input_shape, 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16], 32, 0.00005, 0.0,
                                optimizer=optimizers.Adam(lr=1e-3), nclasses=10, loss_weights=[1.0, 0.1], fMerge=sign_max)
    m2.load_weights(path, by_name=True, skip_mismatch=True)
    assert np.array_equal(m2.predict(X)[0], sig)
    w = model.get_layer("classprob").get_weights()
    assert [a.shape for a in w] == [(32, 10), (10,)]                # Keras order: (in,out) kernel, then bias
    wb = model.get_layer("ofBranch").get_weights()                  # Sequential: kernel, bias of every sublayer in graph order
    assert [a.shape for a in wb][:4] == [(7, 7, 6, 8), (8,), (5, 5, 8, 8), (8,)] and len(wb) == 12
    # the file is a regular HDF5 file in Keras' save_weights layout
    from ugaitnet_b200 import hdf5
    assert hdf5.is_hdf5(path)
    f = hdf5.File(path)
    names = [n.decode() for n in f.attrs["layer_names"]]
    assert names[:3] == ["ofBranch", "grayBranch", "depthBranch"] and names[-1] == "classprob"
    wn = [n.decode() for n in f["grayBranch"].attrs["weight_names"]]
    assert wn[0] == "conv2d_4/kernel:0" and wn[-1] == "ofCode/bias:0"
    assert f["classprob/classprob/kernel:0"].value.shape == (32, 10)
    eer, thr = mj_eerVerifDist(np.array([1, 1, 1, 1, 1, 0, 0, 0, 0]),
                               np.array([0.01, 0.02, 0.015, 0.08, 0.05, 0.07, 0.2, 0.15, 0.18]))
    assert eer == pytest.approx(0.25) and thr == pytest.approx(0.07)   # the reference demo's known answer


def test_triplet_loss_closure_and_unsupported_options(compat_path):
    from nets.triplet_loss_all import triplet_loss
    from nets.mj_uwyhNets_ba import UWYHSemiNet3Mods
    loss = triplet_loss(0.2)
    loss.margin = np.float32(0.2)
    rng = np.random.default_rng(0)
    e = rng.normal(size=(3, 12, 16)).astype(np.float32)
    lab = np.repeat(np.arange(4), 3).reshape(-1, 1).astype(np.float32)
    got = float(loss(lab, e))
    ref, _ = O.triplet_loss_all_literal_np(lab, e, 0.2)
    assert got == pytest.approx(ref, rel=1e-5)
    with pytest.raises(ValueError, match="gaitset"):        # the reference builds GaitSet branches in its LeakyReLU path only
        UWYHSemiNet3Mods.build([(25, 60, 60, 1)] * 3, 4, [7, 5, 3, 2], [96, 192, 512, 512], gaitset=True)
    with pytest.raises(NotImplementedError, match="aux_losses with gaitset"):
        UWYHSemiNet3Mods.build([(25, 60, 60, 1)] * 3, 4, [7, 5, 3, 2], [96, 192, 512, 512], gaitset=True,
                               fActivation='lrelu', aux_losses=True, nclasses=5)
    with pytest.raises(NotImplementedError, match="use3D_with_gaitset"):
        UWYHSemiNet3Mods.build([(25, 60, 60, 1)] * 3, 4, [7, 5, 3, 2], [96, 192, 512, 512], gaitset=True,
                               fActivation='lrelu', use3D=True)


def test_gaitset_builder_protocol(compat_path, tmp_path):
    """UWYHSemiNet3Mods.build_or_load(..., gaitset=True) (mains/mj_trainUWYHGaitNet_DataGen_CasiaB.py:345-355):
    training protocol, descriptor layers, encode(gaitset=True) and the weight round trip, against the oracle."""
    from nets.mj_uwyhNets_ba import UWYHSemiNet, UWYHSemiNet3Mods
    from ugaitnet_b200.compat import Model, optimizers, sign_max
    from oracle import gaitset_oracle as G
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    shapes = [(3, 12, 12, 2), (3, 12, 12, 1), (3, 12, 12, 1)]
    model = UWYHSemiNet3Mods.build_or_load(shapes, 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [96, 192, 512, 512], [256, 16],
                                           0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3), margin=0.2, nclasses=12,
                                           loss_weights=[1.0, 0.1], fMerge=sign_max, fActivation='lrelu', gaitset=True)
    oc = G.GaitSetConfig(in_channels=(2, 1, 1), frames=3, hw=12, nc=16, nclasses=12, merge=O.MERGE_SIGNMAX, wver=1.0, wid=0.1)
    xs, fl, lab = G.synth_batch(oc, 3, 2, seed=5)

    class Gen:
        def __len__(self):
            return 2

        def __getitem__(self, i):
            X = [xs[0].numpy(), fl[0].numpy(), xs[1].numpy(), fl[1].numpy(), xs[2].numpy(), fl[2].numpy()]
            return X, [lab.numpy().reshape(-1, 1).astype(np.float32), np.eye(12, dtype=np.float32)[lab.numpy()]]

        def on_epoch_end(self):
            pass

    gen = Gen()
    X, _ = gen[0]
    P = {k: v.double().cpu() for k, v in model.engine.export_params().items()}
    outs = G.model_forward([x.double() for x in xs], [f.double() for f in fl], P, oc, return_all=True)
    sig, prob = model.predict(X)
    assert sig.shape == (62, 6, 256) and np.allclose(sig, outs["signature"].numpy(), atol=2e-5)
    assert np.allclose(prob, torch.softmax(outs["logits"], 1).numpy(), atol=2e-5)
    flat = Model(model.input, model.get_layer("flatten").output).predict(X)       # typecode 3 descriptor
    assert flat.shape == (6, 62 * 16) and np.allclose(flat, outs["code"].permute(1, 0, 2).flatten(1).numpy(), atol=2e-5)
    codes = UWYHSemiNet.encode(model, [X[0], X[2]], [X[1], X[3]], gaitset=True)
    g0 = outs["branch0"] * fl[0].double().reshape(1, -1, 1)
    g1 = outs["branch1"] * fl[1].double().reshape(1, -1, 1)
    ref = O.l2_normalize(torch.maximum(g0, g1), 1).numpy()
    # columns whose entries are all tiny are amplified by 1/norm (batch-axis normalisation): compare in norm
    assert np.linalg.norm(codes - ref) <= 1e-4 * np.linalg.norm(ref) and np.abs(codes - ref).max() < 1e-3
    model, hist = UWYHSemiNet.fit_generator(model, 3, [], gen, gen, 0, len(gen), 1)
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    path = str(tmp_path / "gs_weights.hdf5")
    model.save_weights(path)
    m2 = UWYHSemiNet3Mods.build(shapes, 4, [7, 5, 3, 2], [96, 192, 512, 512], [256, 16], nclasses=12,
                                loss_weights=[1.0, 0.1], fMerge=sign_max, fActivation='lrelu', gaitset=True)
    m2.load_weights(path, by_name=True)
    assert np.array_equal(m2.predict(X)[0], model.predict(X)[0])
    w = model.get_layer("ofBranch").layers[0].get_weights()
    assert [a.shape for a in w] == [(5, 5, 2, 32)]                               # Keras (kh,kw,cin,cout) kernel


def test_gaitset_single_modality_builder(compat_path, tmp_path):
    """The README "blsingle" recipe: mains/mj_trainUWYHGaitNet_DataGen_CasiaB_1mod.py:329-333 calls
    UWYHSemiNet.build_or_load(ONE shape (25,60,60,1), ..., [dropout0, dropout], ..., postriplet, gaitset=True); build_or_load
    switches to the LeakyReLU path (:588-589) and build() makes the 1-modality GaitSet graph (:776-777, :890-905): the
    branch output [62,B,256] is the signature, no gate / fusion / normalisation, "classprob" on transpose + Flatten."""
    from nets.mj_uwyhNets_ba import UWYHSemiNet
    from ugaitnet_b200.compat import Model, optimizers
    from oracle import gaitset_oracle as G
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    model = UWYHSemiNet.build_or_load((3, 12, 12, 1), 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [96, 192, 512, 512], [256],
                                      0.00005, [0.0, 0.4], optimizer=optimizers.Adam(lr=1e-3), margin=0.2, nclasses=12,
                                      loss_weights=[1.0, 0.1], initnet="", freeze_convs=False, use3D=False,
                                      smoothlabels=0, freeze_all=False, postriplet=1, gaitset=True)
    assert model.gaitset and not model.multimodal and model.cfg.single and model.cfg.nmods == 1
    oc = G.GaitSetConfig(in_channels=(1,), frames=3, hw=12, nc=0, nclasses=12, wver=1.0, wid=0.1, single=True)
    xs, fl, lab = G.synth_batch(oc, 3, 2, seed=5)

    class Gen:
        def __len__(self):
            return 2

        def __getitem__(self, i):       # the 1-modality generator yields the volume alone (no use-flags)
            return xs[0].numpy(), [lab.numpy().reshape(-1, 1).astype(np.float32), np.eye(12, dtype=np.float32)[lab.numpy()]]

        def on_epoch_end(self):
            pass

    gen = Gen()
    X, _ = gen[0]
    P = {k: v.double().cpu() for k, v in model.engine.export_params().items()}
    outs = G.model_forward([xs[0].double()], None, P, oc, return_all=True)
    sig, prob = model.predict(X)
    assert sig.shape == (62, 6, 256) and np.allclose(sig, outs["signature"].numpy(), atol=2e-5)
    assert np.allclose(prob, torch.softmax(outs["logits"], 1).numpy(), atol=2e-5)
    flat = Model(model.input, model.get_layer("flatten").output).predict(X)       # typecode 3 descriptor
    assert flat.shape == (6, 62 * 256) and np.allclose(flat, outs["signature"].permute(1, 0, 2).flatten(1).numpy(), atol=2e-5)
    logs = model.train_on_batch(X, gen[0][1])
    assert np.isfinite(logs["loss"]) and "classprob_acc" in logs
    model, hist = UWYHSemiNet.fit_generator(model, 3, [], gen, None, 0, len(gen), 1)
    assert min(hist.history["loss"][1:]) < hist.history["loss"][0]
    path = str(tmp_path / "gs1_weights.hdf5")
    model.save_weights(path)
    m2 = UWYHSemiNet.build((3, 12, 12, 1), 4, [7, 5, 3, 2], [96, 192, 512, 512], [256], nclasses=12,
                           loss_weights=[1.0, 0.1], fActivation='lrelu', gaitset=True)
    m2.load_weights(path, by_name=True)
    assert np.array_equal(m2.predict(X)[0], model.predict(X)[0])


def test_model_save_loadnet_initnet_freeze_and_init_branches(compat_path, tmp_path):
    """model.save -> loadnet (:1008-1029), build_or_load(initnet=..., freeze_convs / freeze_all) (:1325-1391) with
    classifier "surgery" (a different nclasses keeps every compatible layer), init_branches (:57-75) and the optimiser
    state (m, v, AMSGrad vhat, iterations, lr) in the checkpoint."""
    from nets.mj_uwyhNets_ba import UWYHSemiNet3Mods
    from ugaitnet_b200.compat import optimizers, sign_max
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    shp, fs, fn = [(6, 60, 60), (4, 60, 60), (4, 60, 60)], [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16]
    model = UWYHSemiNet3Mods.build(shp, 4, fs, fn, [32, 16], 0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3, amsgrad=True),
                                   nclasses=10, loss_weights=[1.0, 0.1], fMerge=sign_max)
    oc = O.NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nc=16, nclasses=10)
    gen = FakeGen(oc, n=2)
    X, y = gen[0]
    for _ in range(2):
        model.train_on_batch(X, y)
    model.optimizer.lr = 5e-4
    path = str(tmp_path / "model-state-0002.hdf5")
    model.save(path)
    model.save_weights(UWYHSemiNet3Mods.get_weights_filename(path))
    W = model.engine.export_params()                                              # the state in the checkpoint
    m2 = UWYHSemiNet3Mods.loadnet(path)
    assert np.array_equal(m2.predict(X)[0], model.predict(X)[0])
    e, e2 = model.engine, m2.engine
    assert e2.t == e.t == 2 and e2.lr == pytest.approx(5e-4) and e2.optimizer == "amsgrad"
    assert torch.equal(e2.m, e.m) and torch.equal(e2.v, e.v) and torch.equal(e2.vhat, e.vhat)
    l1, l2 = model.train_on_batch(X, y), m2.train_on_batch(X, y)                 # resumed training continues identically
    assert l2["loss"] == pytest.approx(l1["loss"], rel=1e-6)
    # initnet with another classifier width + freeze_convs
    m3 = UWYHSemiNet3Mods.build_or_load(shp, 4, fs, fn, [32, 16], 0.00005, 0.0, optimizer=optimizers.SGD(1e-2, 0.9),
                                        nclasses=7, loss_weights=[1.0, 0.1], initnet=path, freeze_convs=True, fMerge=sign_max)
    W3 = m3.engine.export_params()
    assert torch.equal(W3["ofBranch/conv2/w"], W["ofBranch/conv2/w"]) and torch.equal(W3["code/w"], W["code/w"])
    assert W3["classprob/w"].shape == (7, 16)                                     # skipped: built fresh
    assert sorted(m3.engine.frozen()) == sorted(k for k in W3 if "/conv" in k)
    assert m3.get_layer("ofBranch").layers[0].trainable is False and m3.get_layer("ofBranch").layers[5].trainable is True
    gen7 = FakeGen(O.NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nc=16, nclasses=7), n=1)
    m3.train_on_batch(*gen7[0])
    W3b = m3.engine.export_params()
    assert torch.equal(W3b["grayBranch/conv0/w"], W3["grayBranch/conv0/w"])       # frozen
    assert not torch.equal(W3b["grayBranch/dense/w"], W3["grayBranch/dense/w"])   # trained
    m4 = UWYHSemiNet3Mods.build_or_load(shp, 4, fs, fn, [32, 16], 0.00005, 0.0, optimizer=optimizers.SGD(1e-2, 0.9),
                                        nclasses=10, loss_weights=[1.0, 0.1], initnet=path, freeze_all=True, fMerge=sign_max)
    assert sorted(m4.engine.frozen()) == sorted(k for k in W if k.split("/")[0].endswith("Branch"))
    # init_branches: a stand-alone branch file initialises the same-named branch
    bpath = str(tmp_path / "gray_branch.hdf5")
    model.save_branch(bpath, "grayBranch")
    W = model.engine.export_params()                                              # (the model trained on after the checkpoint)
    m5 = UWYHSemiNet3Mods.build(shp, 4, fs, fn, [32, 16], 0.00005, 0.0, optimizer=optimizers.SGD(1e-2, 0.9), nclasses=10,
                                loss_weights=[1.0, 0.1], init_branches={"of": "", "gray": bpath, "depth": ""},
                                freeze_branches=False, fMerge=sign_max)
    W5 = m5.engine.export_params()
    assert all(torch.equal(W5[k], W[k]) for k in W if k.startswith("grayBranch/"))
    assert not torch.equal(W5["ofBranch/dense/w"], W["ofBranch/dense/w"])


def test_postriplet2_builder(compat_path):
    """UWYHSemiNet.build(..., postriplet=2) (2-modality builder, :819-832): 'signature' is the Dense layer, 'code' its
    l2_normalize -- the embedding of the triplet loss."""
    from nets.mj_uwyhNets_ba import UWYHSemiNet
    from ugaitnet_b200.compat import Model, optimizers
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    model = UWYHSemiNet.build([(6, 60, 60), (4, 60, 60)], 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16], [32, 8],
                              0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3), nclasses=7, loss_weights=[1.0, 0.5],
                              postriplet=2)
    oc = O.NetConfig(in_channels=(6, 4), filters_numbers=(8, 8, 16, 16), nd=32, nc=8, nclasses=7, merge=O.MERGE_MAX,
                     wver=1.0, wid=0.5, postriplet=2)
    gen = FakeGen(oc, n=1, base_rows=8, expand=2)
    X, y = gen[0]
    P = {k: v.double().cpu() for k, v in model.engine.export_params().items()}
    outs = O.model_forward([torch.tensor(X[0]), torch.tensor(X[2])], [torch.tensor(X[1]), torch.tensor(X[3])], P, oc,
                           return_all=True)
    code = Model(model.input, model.get_layer("code").output).predict(X)
    assert code.shape == (16, 8) and np.allclose(code, outs["code"].numpy(), atol=2e-5)
    assert np.allclose(np.linalg.norm(code, axis=1), 1.0, atol=1e-5)
    sig, prob = model.predict(X)                 # output 0 of the postriplet-2 model is the normalised code (:829)
    assert sig.shape == (16, 8) and np.allclose(sig, outs["code"].numpy(), atol=2e-5)
    assert np.allclose(prob, torch.softmax(outs["logits"], 1).numpy(), atol=2e-5)
    logs = model.train_on_batch(X, y)
    res, _ = O.loss_and_grads([torch.tensor(X[0]), torch.tensor(X[2])], [torch.tensor(X[1]), torch.tensor(X[3])],
                              torch.tensor(y[0]), P, oc)
    assert logs["signature_loss"] == pytest.approx(float(res["triplet"]), rel=1e-5)
    assert logs["loss"] == pytest.approx(float(res["loss"]), rel=1e-5)


def test_postriplet2_gaitset_builder(compat_path):
    """UWYHSemiNet.build(two shapes, [nd, nc], postriplet=2, gaitset=True) (:748-765 + :814-832): output 0 is the
    normalised code [62, B, nc]; losses against the oracle restatement."""
    from nets.mj_uwyhNets_ba import UWYHSemiNet
    from ugaitnet_b200.compat import Model, optimizers, sign_max
    from oracle import gaitset_oracle as G
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    model = UWYHSemiNet.build([(3, 12, 12, 2), (3, 12, 12, 1)], 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [96, 192, 512, 512],
                              [256, 16], 0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3), margin=0.2, nclasses=10,
                              loss_weights=[1.0, 0.5], postriplet=2, fMerge=sign_max, fActivation='lrelu', gaitset=True)
    assert model.engine.post2 and model.cfg.postriplet == 2
    oc = G.GaitSetConfig(in_channels=(2, 1), frames=3, hw=12, nc=16, nclasses=10, merge=O.MERGE_SIGNMAX, wver=1.0, wid=0.5,
                         postriplet=2)
    xs, fl, lab = G.synth_batch(oc, 3, 2, seed=5)
    X = [xs[0].numpy(), fl[0].numpy(), xs[1].numpy(), fl[1].numpy()]
    y = [lab.numpy().reshape(-1, 1).astype(np.float32), np.eye(10, dtype=np.float32)[lab.numpy()]]
    P = {k: v.double().cpu() for k, v in model.engine.export_params().items()}
    outs = G.model_forward([x.double() for x in xs], [f.double() for f in fl], P, oc, return_all=True)
    sig, prob = model.predict(X)
    assert sig.shape == (62, 6, 16) and np.allclose(sig, outs["code"].numpy(), atol=2e-5)
    assert np.allclose(prob, torch.softmax(outs["logits"], 1).numpy(), atol=2e-5)
    dense = Model(model.input, model.get_layer("signature").output).predict(X)    # the Dense layer named "signature"
    assert dense.shape == (62, 6, 16) and np.allclose(dense, outs["signature_layer"].numpy(), atol=2e-5)
    res, _ = G.loss_and_grads([x.double() for x in xs], [f.double() for f in fl], lab, P, oc)
    logs = model.train_on_batch(X, y)
    assert logs["signature_loss"] == pytest.approx(float(res["triplet"]), rel=1e-5)
    assert logs["classprob_loss"] == pytest.approx(float(res["ce"]), rel=1e-5)


def test_standalone_branch_builders_and_small_entry_points(compat_path, tmp_path):
    """UWYHSemiNet.build_gaitset_branch / build_3Dbranch{,LReLU} (:336-484) as stand-alone branch models, fc_loadBranch
    (:57-62), mj_buildnet_by_config (:299-330), MatMul (:23-48), UWYHNet.encode / fit_generator (:248-295)."""
    from nets.mj_uwyhNets_ba import MatMul, UWYHNet, UWYHSemiNet, UWYHSemiNet3Mods, fc_loadBranch, mj_buildnet_by_config
    from ugaitnet_b200.compat import optimizers
    from oracle import gaitset_oracle as G
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    # GaitSet branch alone: predict() is the [62, B, 256] branch output
    gb = UWYHSemiNet.build_gaitset_branch("ofBranch", None, input_shape=(3, 12, 12, 2))
    oc = G.GaitSetConfig(in_channels=(2,), frames=3, hw=12, nc=0, nclasses=0, single=True)
    xs, _, _ = G.synth_batch(oc, 2, 2, seed=3)
    P = {k: v.double().cpu() for k, v in gb.engine.export_params().items()}
    ref = G.gaitset_branch_forward(xs[0].double(), P, "ofBranch", oc)
    out = gb.predict(xs[0].numpy())
    assert out.shape == (62, 4, 256) and np.allclose(out, ref.numpy(), atol=2e-5)
    # the MatMul stand-in reproduces the last layer of that branch on the host
    mm = MatMul(bin_num=31, hidden_dim=256).set_kernel(P["ofBranch/matmul/w"].numpy())
    assert mm.get_config()["bin_num"] == 31 and mm(np.zeros((62, 4, 128), np.float32)).shape == (62, 4, 256)
    # Conv3D branches alone
    for b3 in (UWYHSemiNet.build_3Dbranch("grayBranch", ndense_units=32),
               UWYHSemiNet.build_3DbranchLReLU("grayBranch", ndense_units=32, alpha=0.2)):
        y = b3.predict(np.random.default_rng(0).random((2, 25, 60, 60, 1), dtype=np.float32) - 0.5)
        assert y.shape == (2, 32) and np.isfinite(y).all() and b3.cfg.is3d(0)
    # fc_loadBranch: the first branch of a saved model, stand-alone, with the saved tensors
    shapes, fs, fn = [(6, 60, 60), (4, 60, 60), (4, 60, 60)], [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16]
    model = UWYHSemiNet3Mods.build(shapes, 4, fs, fn, 32, 0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3), nclasses=10,
                                   loss_weights=[1.0, 0.1])
    path = str(tmp_path / "model-final.hdf5")
    model.save(path)
    br = fc_loadBranch(path)
    Pm, Pb = model.engine.export_params(), br.engine.export_params()
    assert set(Pb) == {k for k in Pm if k.startswith("ofBranch/")} and all(torch.equal(Pb[k], Pm[k]) for k in Pb)
    x = np.random.default_rng(1).random((3, 6, 60, 60), dtype=np.float32) - 0.5
    yb = br.predict(x)
    ob = O.NetConfig(in_channels=(6,), filters_numbers=(8, 8, 16, 16), nd=32, nclasses=0, single=True)
    yr, _ = O.model_forward([torch.tensor(x).double()], None, {k: v.double().cpu() for k, v in Pb.items()}, ob)
    assert yb.shape == (3, 32) and np.allclose(yb, yr.numpy(), atol=2e-5)
    # mj_buildnet_by_config through any builder
    nc = dict(filters_size=fs, filters_numbers=fn, input_shape=shapes, ndense_units=32, weight_decay=0.00005, dropout=0.0,
              nclasses=10, optimizer=optimizers.Adam(lr=1e-3), margin=0.2, loss_weights=[1.0, 0.1])
    m2 = mj_buildnet_by_config(nc, UWYHSemiNet3Mods.build)
    assert m2.cfg.nmods == 3 and m2.cfg.nd == 32 and m2.cfg.nclasses == 10
    # UWYHNet.encode == UWYHSemiNet.encode on a 2-modality model
    m3 = UWYHSemiNet.build(shapes[:2], 4, fs, fn, 32, 0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3), nclasses=10,
                           loss_weights=[1.0, 0.1])
    rng = np.random.default_rng(2)
    bd = [rng.random((4, 6, 60, 60), dtype=np.float32) - 0.5, rng.random((4, 4, 60, 60), dtype=np.float32) - 0.5]
    ud = [np.ones((4, 1), np.float32), np.array([[1], [0], [1], [0]], np.float32)]
    e1, e2 = UWYHNet.encode(m3, list(bd), list(ud)), UWYHSemiNet.encode(m3, list(bd), list(ud))
    assert e1.shape == (4, 32) and np.allclose(e1, e2, rtol=1e-5, atol=1e-6)     # (fp32 kernels accumulate with atomics)
    assert np.allclose(np.linalg.norm(e1, axis=1), 1.0, atol=1e-5)
    assert UWYHSemiNet.get_weights_filename("/a/b/model-final.hdf5") == "/a/b/model-final_weights.hdf5"


def test_compile_hard(compat_path):
    """UWYHSemiNet3Mods.compile_hard(model, optimizer, loss_weights, margin) (:1302-1306): the compiled model switches to
    tfa's TripletHardLoss (ugn_triplet_hard), keeps its CE loss and weights, and starts a fresh optimiser."""
    from nets.mj_uwyhNets_ba import UWYHSemiNet3Mods, TripletHardLoss
    from ugaitnet_b200.compat import optimizers
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    shapes = [(6, 60, 60), (4, 60, 60), (4, 60, 60)]
    model = UWYHSemiNet3Mods.build(shapes, 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16], 32, 0.00005, 0.0,
                                   optimizer=optimizers.Adam(lr=1e-3), nclasses=10, loss_weights=[1.0, 0.1])
    oc = O.NetConfig(in_channels=(6, 4, 4), filters_numbers=(8, 8, 16, 16), nd=32, nclasses=10, merge=O.MERGE_MAX,
                     wver=0.8, wid=0.3, margin=0.3, triplet_hard=True)
    gen = FakeGen(oc, n=1, base_rows=6, expand=4)
    X, y = gen[0]
    model.train_on_batch(X, y)                                    # one batch-all step first: optimiser state exists
    W = {k: v.clone() for k, v in model.engine.export_params().items()}
    ret = UWYHSemiNet3Mods.compile_hard(model, optimizers.SGD(0.01, 0.9), [0.8, 0.3], 0.3)
    assert ret is model and isinstance(model.loss[0], TripletHardLoss) and model.loss[0].margin == 0.3
    assert model.engine.t == 0 and float(model.engine.m.abs().max()) == 0.0 and float(model.engine.v.abs().max()) == 0.0
    assert all(torch.equal(W[k], v) for k, v in model.engine.export_params().items())
    P = {k: v.double().cpu() for k, v in W.items()}
    xs = [torch.tensor(X[0]), torch.tensor(X[2]), torch.tensor(X[4])]
    fl = [torch.tensor(X[1]), torch.tensor(X[3]), torch.tensor(X[5])]
    res, G = O.loss_and_grads(xs, fl, torch.tensor(y[0]), P, oc)
    logs = model.train_on_batch(X, y)
    assert logs["signature_loss"] == pytest.approx(float(res["triplet"]), rel=1e-5)
    assert logs["loss"] == pytest.approx(float(res["loss"]), rel=1e-5)
    # SGD(0.01, momentum 0.9), first step: w -= lr * g
    W2 = model.engine.export_params()
    k = "ofBranch/dense/w"
    step = (W[k].double().cpu() - W2[k].double().cpu()) / 0.01
    assert float((step - G[k]).norm() / G[k].norm()) < 1e-4
    # the loss object itself is callable like the Keras one
    sig = model.predict(X)[0]
    v = model.loss[0](y[0], sig)
    want, _ = O.triplet_hard_loss(torch.tensor(y[0]).reshape(-1), torch.tensor(np.asarray(sig), dtype=torch.float64), 0.3)
    assert float(v) == pytest.approx(float(want), rel=1e-4)


def test_pair_network_builder(compat_path):
    """UWYHNet.build (:154-245): nine inputs, the output is VerifLossLayer's value; a training step moves it down."""
    from nets.mj_uwyhNets_ba import UWYHNet
    from ugaitnet_b200.compat import optimizers
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    model = UWYHNet.build([(6, 60, 60), (4, 60, 60)], 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16], 32, 0.00005, 0.0,
                          optimizer=optimizers.SGD(0.05, 0.9), margin=0.9)
    oc = O.NetConfig(in_channels=(6, 4), filters_numbers=(8, 8, 16, 16), nd=32, nclasses=0, merge=O.MERGE_MAX, margin=0.9,
                     pair_loss=True)
    rng = np.random.default_rng(2)
    B = 5
    X = [rng.normal(size=(B, 6, 60, 60)) * 0.3, np.ones((B, 1)), rng.uniform(-0.5, 0.5, size=(B, 4, 60, 60)), np.ones((B, 1)),
         rng.normal(size=(B, 6, 60, 60)) * 0.3, np.ones((B, 1)), rng.uniform(-0.5, 0.5, size=(B, 4, 60, 60)),
         np.array([1, 1, 0, 1, 1.0]).reshape(B, 1), np.array([1, 0, 1, 0, 0]).reshape(B, 1)]
    P = {k: v.double().cpu() for k, v in model.engine.export_params().items()}
    xs = [torch.tensor(np.concatenate([X[0], X[4]])), torch.tensor(np.concatenate([X[2], X[6]]))]
    fl = [torch.tensor(np.concatenate([X[1], X[5]])), torch.tensor(np.concatenate([X[3], X[7]]))]
    outs = O.model_forward(xs, fl, P, oc, return_all=True)
    want = float(O.pair_verif_loss(torch.tensor(X[8]), outs["signature"], 0.9))
    assert float(model.predict(X)) == pytest.approx(want, rel=1e-5)
    e1 = model.embed(X[:4])
    assert np.allclose(e1, outs["signature"][:B].numpy(), atol=2e-6)
    l0 = model.train_on_batch(X)
    for _ in range(5):
        l1 = model.train_on_batch(X)
    assert l1 < l0
    with pytest.raises(ValueError, match="9 inputs"):
        model.train_on_batch(X[:8])


def test_stand_alone_branch_builders(compat_path, tmp_path):
    """UWYHNet.buildBranch / buildBranchLReLU (:67-152): one branch as a model of its own; init_branch loads a saved one."""
    from nets.mj_uwyhNets_ba import UWYHNet
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "fp32"
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.5, 0.5, size=(3, 5, 60, 60))
    for builder, act in ((UWYHNet.buildBranch, O.ACT_RELU), (UWYHNet.buildBranchLReLU, O.ACT_LEAKY)):
        br = builder("grayBranch", (5, 60, 60), 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16], 16, 0.00005, 0.0, "")
        oc = O.NetConfig(in_channels=(5,), filters_numbers=(8, 8, 16, 16), nd=16, nclasses=0, single=True, act=act)
        P = {k: v.double().cpu() for k, v in br.engine.export_params().items()}
        want = O.branch_forward(torch.tensor(x), P, "ofBranch", oc)
        got = br.predict(x)
        assert got.shape == (3, 16) and np.allclose(got, want.numpy(), atol=2e-6)
    path = str(tmp_path / "branch.hdf5")
    br.save_branch(path, "ofBranch")
    br2 = UWYHNet.buildBranchLReLU("ofBranch", (5, 60, 60), 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [8, 8, 16, 16], 16, 0.00005,
                                   0.0, init_branch=path)
    assert np.array_equal(br2.predict(x), br.predict(x))


def test_use3d_builders(compat_path, tmp_path):
    """use3D (:336-417, :721-745, :1077-1099): Conv3D branches for every modality but the 50-channel optical flow, on
    [B,25,60,60,1] inputs; forward against the oracle, a training step, Keras-layout weight round trip (conv3d_N,
    grayCode)."""
    from nets.mj_uwyhNets_ba import UWYHSemiNet, UWYHSemiNet3Mods
    from ugaitnet_b200.compat import Model, optimizers
    from ugaitnet_b200 import hdf5
    import nets.mj_uwyhNets_ba as mod
    mod.MATH_MODE = "f16mix"                 # the builder itself falls back to the fp32 engine for use3D
    rng = np.random.default_rng(3)
    model = UWYHSemiNet3Mods.build([(50, 60, 60), (25, 60, 60, 1), (25, 60, 60, 1)], 4, [(7, 7), (5, 5), (3, 3), (2, 2)],
                                   [32, 32, 32, 64], 32, 0.00005, 0.0, optimizer=optimizers.Adam(lr=1e-3), nclasses=6,
                                   loss_weights=[1.0, 0.1], use3D=True)
    assert model.engine.math_mode == "fp32" and tuple(model.cfg.branch3d) == (False, True, True)
    B = 4
    X = [rng.normal(size=(B, 50, 60, 60)) * 0.3, np.ones((B, 1)), rng.uniform(-0.5, 0.5, size=(B, 25, 60, 60, 1)),
         np.array([1, 0, 1, 1.0]).reshape(B, 1), rng.uniform(-0.5, 0.5, size=(B, 25, 60, 60, 1)), np.ones((B, 1))]
    lab = np.array([0, 0, 1, 1])
    y = [lab.reshape(B, 1), np.eye(6)[lab]]
    oc = O.NetConfig(in_channels=(50, 25, 25), filters_numbers=(32, 32, 32, 64), nd=32, nclasses=6, merge=O.MERGE_MAX,
                     wver=1.0, wid=0.1, branch3d=(False, True, True))
    P = {k: v.double().cpu() for k, v in model.engine.export_params().items()}
    assert tuple(P["grayBranch/conv0/w"].shape) == (64, 1, 3, 5, 5) and tuple(P["depthBranch/ofCode/w"].shape) == (32, 512)
    outs = O.model_forward([torch.tensor(X[0]), torch.tensor(X[2]), torch.tensor(X[4])],
                           [torch.tensor(X[1]), torch.tensor(X[3]), torch.tensor(X[5])], P, oc, return_all=True)
    sig = Model(model.input, model.get_layer("signature").output).predict(X)
    assert np.allclose(sig, outs["signature"].numpy(), atol=2e-5)
    l0 = model.train_on_batch(X, y)["loss"]
    for _ in range(4):
        l1 = model.train_on_batch(X, y)["loss"]
    assert np.isfinite(l1) and l1 < l0
    path = str(tmp_path / "w3d.hdf5")
    model.save_weights(path)
    f = hdf5.File(path)
    names = [n.decode() for n in f["grayBranch"].attrs["weight_names"]]
    assert names[0] == "conv3d/kernel:0" and names[-2:] == ["grayCode/kernel:0", "grayCode/bias:0"]
    assert tuple(f["grayBranch"]["conv3d"]["kernel:0"].value.shape) == (3, 5, 5, 1, 64)
    assert tuple(f["depthBranch"]["grayCode"]["kernel:0"].value.shape) == (1, 1, 1, 512, 32)
    W = {k: v.clone() for k, v in model.engine.export_params().items()}
    model.train_on_batch(X, y)
    model.load_weights(path)
    assert all(torch.equal(W[k], v) for k, v in model.engine.export_params().items())
    # one modality: a non-optical-flow input with use3D gets the Conv3D branch, no fusion / normalisation (:738-745)
    single = UWYHSemiNet.build((25, 60, 60, 1), 4, [(7, 7), (5, 5), (3, 3), (2, 2)], [32, 32, 32, 64], 16, 0.00005, 0.0,
                               optimizer=optimizers.SGD(0.01, 0.9), nclasses=6, loss_weights=[1.0, 0.1], use3D=True)
    assert tuple(single.cfg.branch3d) == (True,)
    oc1 = O.NetConfig(in_channels=(25,), nd=16, nclasses=6, single=True, branch3d=(True,))
    P1 = {k: v.double().cpu() for k, v in single.engine.export_params().items()}
    want = O.branch3d_forward(torch.tensor(X[2]), P1, "ofBranch", oc1)
    got = single.predict(X[2])[0]
    assert np.allclose(got, want.numpy(), atol=2e-5)


def test_shim_gpu_knn_is_what_the_test_mains_import():
    """tf_shim.install(gpu_knn=True): `from sklearn.neighbors import KNeighborsClassifier`, the statement INSIDE evalUWYHNet
    (mains/mj_testUWYHGaitNet_open_tum.py:331-341), binds the B200 classifier; fit / predict on numpy arrays as the main
    calls them, labels equal to scikit-learn's own class on tie-free data."""
    import sys
    import sklearn.neighbors as skn
    from ugaitnet_b200.compat import tf_shim
    saved, saved_path = dict(sys.modules), list(sys.path)
    real = getattr(skn, "_ugn_sklearn_knn", skn.KNeighborsClassifier)
    rng = np.random.default_rng(12)
    cent = rng.standard_normal((12, 32))
    y = rng.integers(0, 12, 600)
    G = (cent[y] + 0.4 * rng.standard_normal((600, 32))).astype(np.float32)
    yq = rng.integers(0, 12, 50)
    Q = (cent[yq] + 0.4 * rng.standard_normal((50, 32))).astype(np.float32)
    ref = real(n_neighbors=3).fit(G, y).predict(Q)
    try:
        tf_shim.install(None, gpu_knn=True)
        from sklearn.neighbors import KNeighborsClassifier
        from ugaitnet_b200.knn import KNeighborsClassifier as GpuKNN
        assert KNeighborsClassifier is GpuKNN
        clf = KNeighborsClassifier(n_neighbors=3)
        clf.fit(G, y)
        pred = clf.predict(Q)
        assert isinstance(pred, np.ndarray) and np.array_equal(pred.astype(np.int64), ref.astype(np.int64))
    finally:
        skn.KNeighborsClassifier = real
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
        sys.modules.update(saved)            # (install() re-imports `nets.*`: put the module objects of this session back)
        sys.path[:] = saved_path
