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


def test_chunked_deflate_shuffle_dataset_is_read():
    """A chunked dataset with the shuffle + deflate pipeline (pytables / deepdish compression='zlib'), assembled by hand
    following the specification (layout class 2, filter pipeline message 0x000B version 1, v1 chunk B-tree)."""
    import zlib
    a = (np.arange(6 * 5, dtype=np.int16).reshape(6, 5) * 37 - 400)
    cd = (4, 5)                                            # two chunks along axis 0, the second one partial
    buf = bytearray(2048)

    def put(off, b):
        buf[off:off + len(b)] = b
        return off
    chunks = []
    pos = 1024
    for r0 in (0, 4):
        c = np.zeros(cd, dtype=np.int16)
        c[:min(4, 6 - r0)] = a[r0:r0 + 4]
        raw = c.tobytes()
        sh = np.frombuffer(raw, np.uint8).reshape(-1, 2).T.tobytes()          # shuffle
        z = zlib.compress(sh)
        put(pos, z)
        chunks.append((pos, len(z), r0))
        pos += (len(z) + 7) & ~7
    bt = 512
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, 2, hdf5.UNDEF, hdf5.UNDEF)
    for addr, sz, r0 in chunks:
        node += struct.pack("<IIQQQ", sz, 0, r0, 0, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, 8, 0, 0)
    put(bt, node)
    # dataset object header at 200
    ds = struct.pack("<BBBB4x", 1, 2, 0, 0) + struct.pack("<QQ", 6, 5)
    dt = struct.pack("<B3BI", 0x10, 0x08, 0, 0, 2) + struct.pack("<HH", 0, 16)
    lay = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", 4, 5, 2)
    flt = struct.pack("<BB6x", 1, 2)
    flt += struct.pack("<HHHH", 2, 8, 0, 1) + b"shuffle\0" + struct.pack("<II", 2, 0)
    flt += struct.pack("<HHHH", 1, 8, 0, 1) + b"deflate\0" + struct.pack("<II", 6, 0)
    msgs = [(1, ds), (3, dt), (8, lay), (0x0B, flt)]
    blob = b"".join(struct.pack("<HHB3x", t, len(hdf5._pad8(d)), 0) + hdf5._pad8(d) for t, d in msgs)
    put(200, struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(blob)) + blob)
    # wrap into a file: root group with one link "data" through the writer's machinery is simpler -> patch a Writer file
    w = hdf5.Writer()
    w.dataset("placeholder", np.zeros(1, np.uint8))
    base = bytearray(w.tobytes())
    shift = len(base)
    # relocate: append our hand-made region and point the symbol table entry of "placeholder" at the new header
    region = bytearray(buf)
    # fix absolute addresses inside the region (B-tree address in the layout message, chunk addresses in the node)
    off_lay = region.find(struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt))
    region[off_lay + 3:off_lay + 11] = struct.pack("<Q", bt + shift)
    p = bt + 24
    for addr, sz, r0 in chunks:
        region[p + 32:p + 40] = struct.pack("<Q", addr + shift)
        p += 40
    base += region
    f0 = hdf5.File(bytes(base))
    old_hdr = f0._links["placeholder"]
    i = bytes(base).find(struct.pack("<Q", old_hdr), 96)
    base[i:i + 8] = struct.pack("<Q", 200 + shift)
    f = hdf5.File(bytes(base))
    assert np.array_equal(f["placeholder"].value, a)


def test_sample_files_round_trip_and_decode(tmp_path):
    """data/generateOFData.py:136-148 layout -> ugaitnet_b200.samples.load_sample / decode_sample (= __load_dd)."""
    from ugaitnet_b200 import samples
    rng = np.random.default_rng(3)
    of = rng.integers(-3000, 3000, size=(60, 60, 50)).astype(np.int16)
    s = {"data": of, "label": np.uint16(17), "videoId": np.uint16(3), "gait": np.uint8(1),
         "frames": np.arange(25, dtype=np.uint16), "compressFactor": np.uint8(100)}
    p = tmp_path / "p017-n01-03.h5"
    samples.save_sample(p, s)
    got = samples.load_sample(p)
    assert np.array_equal(got["data"], of) and int(got["label"]) == 17 and int(got["compressFactor"]) == 100
    x = samples.decode_sample(got, ntype=2, clip_max=2300, clip_min=50)
    # literal __load_dd statements (data/mj_dataGeneratorMMUWYHsingle.py:315-324) + moveaxis (:333)
    ref = np.float32(of)
    ref[np.abs(ref) > 2300] = 1e-8
    ref[np.abs(ref) < 50] = 1e-8
    ref = ref / 100
    ref = ref * 0.1
    assert x.shape == (50, 60, 60) and x.dtype == np.float32 and np.allclose(x, np.moveaxis(ref, 2, 0), rtol=1e-6, atol=0)
    gray = rng.integers(0, 256, size=(60, 60, 25)).astype(np.uint8)
    g = samples.decode_sample({"data": gray, "compressFactor": np.uint8(1)})
    assert np.allclose(g, np.moveaxis(np.float32(gray) / 255.0 - 0.5, 2, 0))
    sil = samples.decode_sample({"data": gray, "compressFactor": np.uint8(1)}, silhouette=True)
    assert np.allclose(sil, np.moveaxis(np.float32(gray) / 255.0, 2, 0))
    assert samples.decode_sample({"data": np.zeros((0,), np.uint8), "compressFactor": 1}) is None
