"""ugaitnet_b200.hdf5: pure-Python HDF5 subset (Keras checkpoint interchange).  No HDF5 library exists in
the image, so the reader is tested against the writer and the writer's bytes against the field offsets of the HDF5
File Format Specification (superblock 0, v1 object headers, symbol-table groups, contiguous datasets, v1 attributes)."""
import struct

import numpy as np
import pytest

from ugaitnet_b200 import hdf5


def _sample():
    rng = np.random.default_rng(0)
    w = hdf5.Writer()
    w.root.attrs["layer_names"] = [b"ofBranch", b"classprob"]
    w.root.attrs["backend"] = b"tensorflow"
    w.root.attrs["keras_version"] = "2.4.0"
    g = w.group("ofBranch")
    g.attrs["weight_names"] = [b"ofBranch/conv2d/kernel:0", b"ofBranch/conv2d/bias:0"]
    k = rng.normal(size=(7, 7, 5, 8)).astype(np.float32)
    b = rng.normal(size=(8,)).astype(np.float32)
    w.dataset("ofBranch/ofBranch/conv2d/kernel:0", k)
    w.dataset("ofBranch/ofBranch/conv2d/bias:0", b)
    w.group("classprob").attrs["weight_names"] = [b"classprob/kernel:0"]
    c = rng.normal(size=(16, 3)).astype(np.float64)
    d = w.dataset("classprob/classprob/kernel:0", c)
    d.attrs["note"] = np.arange(5, dtype=np.int32)
    w.dataset("ints/u8", np.arange(12, dtype=np.uint8).reshape(3, 4))
    w.dataset("ints/i64", np.array([-5, 7], dtype=np.int64))
    return w, k, b, c


def test_round_trip_groups_datasets_attributes(tmp_path):
    w, k, b, c = _sample()
    path = tmp_path / "weights.hdf5"
    w.save(path)
    assert hdf5.is_hdf5(path)
    f = hdf5.File(str(path))
    assert sorted(f.keys()) == ["classprob", "ints", "ofBranch"]
    assert [x.decode() for x in f.attrs["layer_names"]] == ["ofBranch", "classprob"]
    assert f.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.4.0"
    g = f["ofBranch"]
    assert [x.decode() for x in g.attrs["weight_names"]] == ["ofBranch/conv2d/kernel:0", "ofBranch/conv2d/bias:0"]
    assert np.array_equal(g["ofBranch/conv2d/kernel:0"].value, k)
    assert np.array_equal(f["ofBranch/ofBranch/conv2d/bias:0"].value, b)
    d = f["classprob/classprob/kernel:0"]
    assert d.value.dtype == np.float64 and np.array_equal(d.value, c)
    assert np.array_equal(d.attrs["note"], np.arange(5, dtype=np.int32))
    assert np.array_equal(f["ints/u8"].value, np.arange(12, dtype=np.uint8).reshape(3, 4))
    assert np.array_equal(f["ints/i64"].value, np.array([-5, 7]))
    assert dict(f.visit()).keys() == {"classprob/classprob/kernel:0", "ints/i64", "ints/u8",
                                      "ofBranch/ofBranch/conv2d/bias:0", "ofBranch/ofBranch/conv2d/kernel:0"}
    with pytest.raises(KeyError):
        f["nope"]


def test_many_links_span_several_symbol_nodes():
    w = hdf5.Writer()
    for i in range(37):                                  # > 8 entries per SNOD: several leaves under one TREE node
        w.dataset(f"g/d{i:02d}", np.full((3,), i, dtype=np.float32))
    f = hdf5.File(w.tobytes())
    assert f["g"].keys() == [f"d{i:02d}" for i in range(37)]
    for i in (0, 8, 36):
        assert np.array_equal(f[f"g/d{i:02d}"].value, np.full((3,), i, dtype=np.float32))


def test_emitted_bytes_follow_the_specification_offsets():
    w = hdf5.Writer()
    w.dataset("x", np.array([1.5, -2.0], dtype=np.float32))
    b = w.tobytes()
    # superblock version 0 (spec III.A): signature, versions, offset / length sizes, K values, addresses
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    assert struct.unpack_from("<HH", b, 16) == (4, 16)
    base, _, eof, _ = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and eof == len(b)
    root = struct.unpack_from("<Q", b, 56 + 8)[0]
    # root object header (version 1, spec IV.A.1.a) carries a symbol table message (0x0011): B-tree + heap
    assert b[root] == 1
    mtype, msize = struct.unpack_from("<HH", b, root + 16)
    assert mtype == 0x11 and msize == 16
    bt, heap = struct.unpack_from("<QQ", b, root + 24)
    assert b[bt:bt + 4] == b"TREE" and b[bt + 4] == 0 and b[bt + 5] == 0          # group node, leaf level
    assert b[heap:heap + 4] == b"HEAP"
    snod = struct.unpack_from("<Q", b, bt + 24 + 8)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 1
    name_off, ohdr = struct.unpack_from("<QQ", b, snod + 8)
    data_seg = struct.unpack_from("<Q", b, heap + 24)[0]
    assert b[data_seg + name_off:data_seg + name_off + 2] == b"x\0"
    # dataset object header: dataspace (1), datatype (3), fill value (5), layout (8)
    nmsg = struct.unpack_from("<H", b, ohdr + 2)[0]
    p, types = ohdr + 16, []
    for _ in range(nmsg):
        t, s = struct.unpack_from("<HH", b, p)
        types.append(t)
        if t == 0x03:      # IEEE little-endian float32: class 1 version 1, size 4, exponent 8 bits at 23, bias 127
            assert b[p + 8] == 0x11 and struct.unpack_from("<I", b, p + 12)[0] == 4
            assert struct.unpack_from("<HHBBBBI", b, p + 16) == (0, 32, 23, 8, 0, 23, 127)
        if t == 0x08:      # layout version 3, contiguous
            assert b[p + 8] == 3 and b[p + 9] == 1
            addr, size = struct.unpack_from("<QQ", b, p + 10)
            assert size == 8 and np.array_equal(np.frombuffer(b, "<f4", 2, addr), [1.5, -2.0])
        p += 8 + s
    assert types == [0x01, 0x03, 0x05, 0x08]


def test_rejects_non_hdf5(tmp_path):
    p = tmp_path / "x.npz"
    np.savez(p, a=np.zeros(3))
    assert not hdf5.is_hdf5(p)
    with pytest.raises(hdf5.HDF5Error):
        hdf5.File(str(p))
