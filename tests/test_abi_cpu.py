"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "ugaitnet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ugn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    from ugaitnet_b200 import build
    path = build.build(verbose=False)
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ugaitnet_b200.h but not exported"


def test_python_binding_covers_header():
    from ugaitnet_b200 import _ffi
    assert set(_header_symbols()) == set(_ffi.EXPORTED_SYMBOLS)
    assert _ffi.lib.ugn_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from ugaitnet_b200 import _ffi
    with pytest.raises(_ffi.UgnError, match="no CPU fallback"):
        _ffi.Ctx(0)
    from ugaitnet_b200.net import UGaitEngine
    from ugaitnet_b200.config import NetConfig
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        UGaitEngine(NetConfig(nd=16))


def test_config_geometry_matches_reference_shapes():
    # nets/mj_uwyhNets_ba.py:67-107 with filters [7,5,3,2]: 60->54->27->23->11->9->4->3, flatten 4608
    from ugaitnet_b200.config import NetConfig
    cfg = NetConfig()
    L = cfg.layers(0, 32)
    assert [(l["h"], l["ho"], l["hp"]) for l in L] == [(60, 54, 27), (27, 23, 11), (11, 9, 4), (4, 3, 3)]
    assert L[0]["cp"] == 64 and cfg.layers(1, 32)[0]["cp"] == 32
    assert cfg.flat == 4608


def test_gaitset_input_layout_matches_generator_restating():
    """to_gaitset_layout against the literal statements of the reference generator
    (data/mj_dataGeneratorMMUWYHsingle_repetitions.py:426-434)."""
    import numpy as np
    from ugaitnet_b200.expand import to_gaitset_layout
    rng = np.random.default_rng(0)
    of = rng.normal(size=(50, 6, 5)).astype(np.float32)
    x_new = np.zeros((25, of.shape[1], of.shape[2], 2), dtype=of.dtype)
    x_new[:, :, :, 0] = of[::2, :, :]
    x_new[:, :, :, 1] = of[1::2, :, :]
    assert np.array_equal(to_gaitset_layout(of), x_new)
    gray = rng.normal(size=(25, 6, 5)).astype(np.float32)
    g_new = np.zeros((25, gray.shape[1], gray.shape[2], 1), dtype=gray.dtype)
    g_new[:, :, :, 0] = gray
    assert np.array_equal(to_gaitset_layout(gray), g_new)
    assert to_gaitset_layout(np.stack([of, of])).shape == (2, 25, 6, 5, 2)
