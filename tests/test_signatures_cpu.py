"""Drop-in check of the Python surface: the builder / training / loss entry points of the
compat package must keep the reference's names, positional order and defaults.  Both sides are compared as SOURCE
(ast), so nothing is imported -- no TensorFlow on the reference side, no GPU on ours.  The reference tree exists only
in the build container: skipped elsewhere."""
import ast
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")


def _functions(path):
    out = {}
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            for f in node.body:
                if isinstance(f, ast.FunctionDef):
                    out[f"{node.name}.{f.name}"] = f
        elif isinstance(node, ast.FunctionDef):
            out[node.name] = node
    return out


def _sig(f):
    a = f.args
    names = [x.arg for x in a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return list(zip(names, defaults))


# defaults that are Keras objects on the reference side and plain stand-ins on ours
_EQUIV = {"optimizers.SGD(0.001, 0.9)": "None"}


@pytest.mark.parametrize("rel,names", [
    ("nets/mj_uwyhNets_ba.py", ["UWYHSemiNet3Mods.build", "UWYHSemiNet3Mods.build_or_load", "UWYHSemiNet3Mods.compile_hard",
                                "UWYHSemiNet.build", "UWYHSemiNet.build_or_load", "UWYHSemiNet.fit_generator",
                                "UWYHSemiNet.encode", "UWYHSemiNet.loadnet", "mj_tensor_times_scalar",
                                "UWYHSemiNet3Mods.loadnet", "UWYHSemiNet.build_by_config", "UWYHSemiNet.get_weights_filename",
                                "UWYHSemiNet.get_netconfig_filename", "UWYHSemiNet.build_3Dbranch",
                                "UWYHSemiNet.build_3DbranchLReLU", "UWYHSemiNet.build_gaitset_branch",
                                "UWYHNet.buildBranchLReLU", "UWYHNet.build", "UWYHNet.fit_generator", "UWYHNet.encode",
                                "fc_loadBranch", "mj_buildnet_by_config", "MatMul.__init__", "MatMul.call",
                                "MatMul.get_config"]),
    ("nets/triplet_loss_all.py", ["triplet_loss"]),
    ("nets/mj_loss.py", ["mj_l2normalize", "mj_smoothL1", "mj_smoothL1bis", "PairLossLayer.pair_loss", "PairLossLayer.call",
                         "VerifLossLayer.pair_loss", "VerifLossLayer.call", "TripletLossLayer.__init__",
                         "TripletLossLayer.triplet_loss", "TripletLossLayer.call"]),
    ("nets/mj_metrics.py", ["mj_eerVerifDist"]),
])
def test_entry_points_keep_reference_signatures(rel, names):
    ref = _functions(os.path.join(REF, rel))
    ours = _functions(os.path.join(ROOT, "ugaitnet_b200", "compat", rel))
    for n in names:
        assert n in ref, f"{n} not found in the reference (survey drift?)"
        assert n in ours, f"{n} missing from the compat package"
        rs, os_ = _sig(ref[n]), _sig(ours[n])
        assert [a for a, _ in rs] == [a for a, _ in os_], (n, rs, os_)
        for (a, dr), (_, do) in zip(rs, os_):
            assert _EQUIV.get(dr, dr) == do or dr == do, (n, a, dr, do)


def test_mj_loss_layers_values_against_literal_restatements():
    """nets/mj_loss.py entry points (a15): values against op-by-op numpy restatements of the reference's backend calls
    (K.abs / K.switch / K.sum, tf.where + tf.gather, K.square / K.sqrt / K.maximum; :11-15, :17-21, :46-50, :73-93,
    :115-119)."""
    import numpy as np
    from ugaitnet_b200.compat.nets import mj_loss as L
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(9, 6)).astype(np.float32), rng.normal(size=(9, 6)).astype(np.float32)
    # mj_l2normalize: K.l2_normalize(x, axis) = x / sqrt(max(sum(x^2), 1e-12))
    ref = a / np.sqrt(np.maximum((a * a).sum(1, keepdims=True), 1e-12))
    assert np.allclose(L.mj_l2normalize(a, axis=1).numpy(), ref, rtol=1e-6)
    z = np.zeros((2, 4), np.float32)
    assert np.array_equal(L.mj_l2normalize(z).numpy(), z)                      # the eps branch: 0 * rsqrt(1e-12) = 0
    # mj_smoothL1 / mj_smoothL1bis / PairLossLayer(alpha): Huber with switch at delta, SUM over everything
    def huber(x, d):
        x = np.abs(x)
        return np.where(x < d, 0.5 * x ** 2, d * (x - 0.5 * d)).sum()
    assert float(L.mj_smoothL1(a, b)) == pytest.approx(huber(a - b, 0.5), rel=1e-6)
    assert float(L.mj_smoothL1bis(None, [a, b])) == pytest.approx(huber(a - b, 0.5), rel=1e-6)
    pl = L.PairLossLayer(alpha=0.3)
    assert float(pl([a, b])) == pytest.approx(huber(a - b, 0.3), rel=1e-6) and len(pl.losses) == 1
    assert pl.get_config() == {"alpha": 0.3}
    # VerifLossLayer: 0.5*sum(pos rows (a-b)^2) + 0.5*max(0, m - sqrt(sum over ALL negative rows (a-b)^2))^2
    lab = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1], np.float32).reshape(-1, 1)
    res2 = (a - b) ** 2
    for m in (0.5, 40.0):                                                         # margin inactive / active
        xpos = 0.5 * res2[lab[:, 0] == 1].sum()
        xneg = 0.5 * max(0.0, m - np.sqrt(res2[lab[:, 0] == 0].sum())) ** 2
        assert float(L.VerifLossLayer(alpha=m)([a, b, lab])) == pytest.approx(xpos + xneg, rel=1e-6)
    # TripletLossLayer(alpha): sum over the batch of max(|a-p|^2 - |a-n|^2 + alpha, 0)  (SQUARED distances)
    n = rng.normal(size=(9, 6)).astype(np.float32)
    ref = np.maximum(((a - b) ** 2).sum(-1) - ((a - n) ** 2).sum(-1) + 0.2, 0).sum(0)
    assert float(L.TripletLossLayer(0.2)([a, b, n])) == pytest.approx(ref, rel=1e-6)
