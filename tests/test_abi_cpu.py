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


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under ugaitnet_b200/ may import, load or execute anything from
    oracle/ (only tests/, __graft_entry__.smoke() and bench.py's CPU legs do)."""
    import pathlib
    import re
    root = pathlib.Path(__file__).resolve().parents[1] / "ugaitnet_b200"
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|importlib\.import_module\(\s*['\"]oracle|oracle/_ref|liboracle", re.M)
    offenders = [str(p) for p in root.rglob("*.py") if pat.search(p.read_text())]
    inc = re.compile(r'#\s*include\s*[<"][^>"]*oracle')
    offenders += [str(p) for p in list(root.rglob("*.cu")) + list(root.rglob("*.cuh")) if inc.search(p.read_text())]
    assert offenders == []


def test_conv3d_branch_geometry_matches_reference_chain():
    # build_3Dbranch (nets/mj_uwyhNets_ba.py:346-363) on [25,60,60,1]: (23,28,28,64) (21,13,13,128) (10,6,6,256)
    # (4,2,2,512) (2,1,1,512) (1,1,1,512), then the 1x1x1 "grayCode"
    from ugaitnet_b200.config import NetConfig
    cfg = NetConfig(in_channels=(50, 25, 25), branch3d=(False, True, True))
    L = cfg.layers3d(1)
    assert [(l["to"], l["ho"], l["co"]) for l in L] == [(23, 28, 64), (21, 13, 128), (10, 6, 256), (4, 2, 512), (2, 1, 512),
                                                        (1, 1, 512)]
    assert [l["cin"] for l in L] == [1, 64, 128, 256, 512, 512] and cfg.is3d(1) and not cfg.is3d(0)
    with pytest.raises(ValueError, match="1x1x1"):
        NetConfig(in_channels=(30,), branch3d=(True,)).layers3d(0)


def test_io_block_layouts_are_aligned_and_disjoint():
    """IOBlock (the single-copy input block): every field 256-byte aligned, fields disjoint, the base-row and raw-sample
    layouts are prefixes that never exceed the full block; raw int16 / uint8 volumes take 0.375 of the f32 bytes."""
    from ugaitnet_b200.net import IOBlock
    B, B0 = 96, 24
    io = IOBlock(torch.device("cpu"), B, [(50, 60, 60), (25, 60, 60), (25, 60, 60)])
    v = io.views(io.dev_buf)
    spans = []
    for name in ("labels", "src_row", "mirror", "shift", "clip"):
        t = v[name]
        spans.append((t.data_ptr() - io.dev_buf.data_ptr(), t.numel() * t.element_size()))
    for t in v["flags"] + v["x"]:
        spans.append((t.data_ptr() - io.dev_buf.data_ptr(), t.numel() * t.element_size()))
    spans.sort()
    assert all(o % IOBlock.ALIGN == 0 for o, _ in spans)
    assert all(a + n <= b for (a, n), (b, _) in zip(spans, spans[1:])) and spans[-1][0] + spans[-1][1] <= io.nbytes_full
    assert [tuple(t.shape) for t in v["x"]] == [(B, 50, 60, 60), (B, 25, 60, 60), (B, 25, 60, 60)]
    vb = io.views(io.dev_buf, B0)
    assert [tuple(t.shape) for t in vb["x"]] == [(B0, 50, 60, 60), (B0, 25, 60, 60), (B0, 25, 60, 60)]
    assert io.header < io.nbytes_base(B0) < io.nbytes_full
    raw = [(torch.int16, 100.0, 0.1, 0.0, 0.0, 0.0), (torch.uint8, 255.0, 1.0, 0.5, 0.0, 0.0), (torch.uint8, 255.0, 1.0, 0.5, 0.0, 0.0)]
    vr = io.views(io.dev_buf, None, raw=raw)
    assert [t.dtype for t in vr["x"]] == [torch.int16, torch.uint8, torch.uint8]
    _, nraw = io.raw_offsets(B, [2, 1, 1])
    vol_f32 = io.nbytes_full - io.header
    assert abs((nraw - io.header) / vol_f32 - 0.375) < 1e-3
    last = vr["x"][-1]
    assert last.data_ptr() - io.dev_buf.data_ptr() + last.numel() <= nraw <= io.nbytes_full
