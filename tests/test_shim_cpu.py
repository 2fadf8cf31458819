"""The drop-in claim on an UNMODIFIED reference main (north_star: "drops into the existing mains/ training and test
scripts").  Runs here, where /root/reference exists (skipped on the GPU box, which has no reference checkout): the
import shim resolves every import of mains/mj_trainUWYHGaitNet_DataGen_3mods.py, `nets.mj_uwyhNets_ba` is the B200
drop-in, and the main's own `mj_computeDistMetrics` (:103-180) runs unchanged on a generator-protocol object, calling
`UWYHSemiNet.encode` + `mj_eerVerifDist` through the drop-in surface.  The engine itself needs a GPU, so the branch
codes come from the CPU oracle here (tests may use it); tests/test_compat_gpu.py covers encode() on the device."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ugait_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mains")), reason="needs the reference checkout")


@pytest.fixture()
def shim():
    saved = dict(sys.modules)
    saved_path = list(sys.path)
    from ugaitnet_b200.compat import tf_shim
    tf_shim.install(REF)
    yield tf_shim
    for k in list(sys.modules):
        if k not in saved:
            del sys.modules[k]
    sys.path[:] = saved_path


def test_unmodified_reference_main_imports_and_runs_its_metric_loop(shim, monkeypatch):
    main = importlib.import_module("mains.mj_trainUWYHGaitNet_DataGen_3mods")
    import ugaitnet_b200.compat.nets.mj_uwyhNets_ba as drop_in
    assert main.UWYHSemiNet3Mods.__module__ == "nets.mj_uwyhNets_ba"
    assert os.path.samefile(sys.modules["nets.mj_uwyhNets_ba"].__file__, drop_in.__file__)
    assert main.mj_eerVerifDist.__module__ == "nets.mj_metrics"
    # the reference's own helper modules outside the hot path still come from the reference checkout
    assert sys.modules["nets.mj_utils"].__file__.startswith(REF)
    # optimizers / callbacks the main instantiates resolve to working stand-ins
    opt = main.optimizers.Adam(lr=1e-4)
    assert opt.name == "adam" and opt.lr == pytest.approx(1e-4)
    cb = main.ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=1, min_lr=1e-6)

    class M:
        class optimizer:
            lr = 1e-3
    cb.set_model(M)
    for e, v in enumerate([1.0, 1.0, 1.0]):
        cb.on_epoch_end(e, {"val_loss": v})
    assert M.optimizer.lr == pytest.approx(2.5e-4)

    # ---- the main's own mj_computeDistMetrics, unchanged, on a generator-protocol object
    oc = O.NetConfig(in_channels=(6, 4, 4), filters_numbers=(4, 4, 8, 8), nd=16, nclasses=10, merge=O.MERGE_MAX)
    P = O.init_params(oc, seed=3, dtype=torch.float64)

    class Gen:
        def __init__(self):
            self.items = []
            for i in range(2):
                xs, fl, lab = O.synth_batch(oc, base_rows=8, expand=2, seed=i)
                self.items.append(([xs[0], fl[0], xs[1], fl[1], xs[2], fl[2]], [(lab % 10), None]))

        def __len__(self):
            return len(self.items)

        def __getitem__(self, i):
            return self.items[i]

    def oracle_encode(model, batch_data, use_data, gaitset=False):
        # what compat UWYHSemiNet.encode computes on the device (first two modalities, gated, Maximum, l2_normalize)
        b0 = O.branch_forward(torch.tensor(batch_data[0], dtype=torch.float64), P, "ofBranch", oc) * torch.tensor(use_data[0])
        b1 = O.branch_forward(torch.tensor(batch_data[1], dtype=torch.float64), P, "grayBranch", oc) * torch.tensor(use_data[1])
        return O.l2_normalize(torch.maximum(b0, b1), 1).numpy()
    monkeypatch.setattr(main.UWYHSemiNet, "encode", staticmethod(oracle_encode))
    np.random.seed(0)
    distances, eer, chance, (codes, labs) = main.mj_computeDistMetrics(object(), Gen(), True, None)
    assert codes.shape == (32, 16) and len(labs) == 32 and len(distances) > 0
    assert 0.0 <= eer <= 1.0 and 0.0 < chance < 1.0
    # the EER the main computed == our restatement of mj_eerVerifDist on the same pairs (a14)
    # (re-derive the pairs with the same numpy seed)
    np.random.seed(0)
    d2, eer2, _, _ = main.mj_computeDistMetrics(object(), Gen(), True, None)
    assert eer2 == eer and np.array_equal(d2, distances)


@pytest.mark.parametrize("main", ["mj_trainUWYHGaitNet_DataGen_3mods", "mj_trainUWYHGaitNet_DataGen_CasiaB",
                                  "mj_trainUWYHGaitNet_DataGen_CasiaB_1mod", "mj_trainUWYHGaitNet_DataGen_1mod",
                                  "mj_testUWYHGaitNet_open_tum", "mj_testUWYHGaitNet_open_casiab"])
def test_every_in_scope_main_imports_through_the_shim(shim, main):
    """Every training / open-world test main of the reference (the both-datasets fork is out of scope) resolves its
    imports -- tensorflow, tensorflow_addons, deepdish, tensorboard, nets.* -- and binds the drop-in builders."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")              # (the reference's own "\\d" escape warnings)
        mod = importlib.import_module("mains." + main)
    import ugaitnet_b200.compat.nets.mj_uwyhNets_ba as drop_in
    nets_ba = importlib.import_module("nets.mj_uwyhNets_ba")      # (the test mains import it under __main__ only)
    assert os.path.samefile(nets_ba.__file__, drop_in.__file__)
    if hasattr(mod, "UWYHSemiNet"):
        assert mod.UWYHSemiNet is nets_ba.UWYHSemiNet


def test_shim_routes_sklearn_knn_to_the_gpu_search_on_request():
    """The open-world test mains import KNeighborsClassifier from sklearn.neighbors inside evalUWYHNet
    (mains/mj_testUWYHGaitNet_open_tum.py:331): install(gpu_knn=True) makes that import yield the B200 classifier."""
    import sklearn.neighbors
    from ugaitnet_b200.compat import tf_shim
    from ugaitnet_b200.knn import KNeighborsClassifier as GpuKNN
    saved = dict(sys.modules)
    saved_path = list(sys.path)
    orig = sklearn.neighbors.KNeighborsClassifier
    try:
        tf_shim.install(REF, gpu_knn=True)
        from sklearn.neighbors import KNeighborsClassifier as K1
        assert K1 is GpuKNN
        for name in ("fit", "predict", "kneighbors"):
            assert callable(getattr(K1, name))
        tf_shim.install(REF)
        from sklearn.neighbors import KNeighborsClassifier as K2
        assert K2 is orig
    finally:
        sklearn.neighbors.KNeighborsClassifier = orig
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
        sys.path[:] = saved_path
