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
                                "UWYHSemiNet.encode", "UWYHSemiNet.loadnet", "mj_tensor_times_scalar"]),
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
