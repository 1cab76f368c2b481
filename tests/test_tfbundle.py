"""TensorFlow tensor-bundle checkpoint files written / read without TensorFlow (lb_wavenet_b200/tfbundle.py).

The reference saves through tf.train.Saver (ckpt.py:41,54-62); TF is not installable here, so the format is pinned
by (1) the RFC 3720 CRC-32C vectors and TF's crc mask constants, (2) a bundle assembled BY HAND in this file from
the published LevelDB table format and tensor_bundle.proto, byte for byte, (3) structural invariants (footer magic,
block trailers, restart points, index separators) and (4) round trips across block and restart boundaries.
"""
import struct

import numpy as np
import pytest

from lb_wavenet_b200 import tfbundle as tb


def test_crc32c_known_answers(lib):
    assert tb.crc32c(b"123456789") == 0xE3069283          # the classic check value
    assert tb.crc32c(bytes(32)) == 0x8A9136AA              # RFC 3720 B.4
    assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    assert tb.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # running value == one shot; odd lengths exercise the byte-wise tail
    blob = np.random.default_rng(0).integers(0, 256, 100003).astype(np.uint8).tobytes()
    assert tb.crc32c(blob[50001:], tb.crc32c(blob[:50001])) == tb.crc32c(blob)
    # tensorflow/core/lib/hash/crc32c.h: Mask rotates right by 15 and adds 0xa282ead8
    c = tb.crc32c(b"foo")
    assert tb.crc_mask(c) != c and tb.crc_unmask(tb.crc_mask(c)) == c
    assert tb.crc_mask(0) == 0xA282EAD8 and tb.crc_mask(1 << 15) == (1 + 0xA282EAD8)


def test_varint_and_proto_records():
    for v in (0, 1, 127, 128, 300, 2 ** 31 - 1, 2 ** 32, 2 ** 63 - 1):
        enc = tb.put_varint(v)
        assert tb.get_varint(enc, 0) == (v, len(enc))
    assert tb.put_varint(300) == b"\xac\x02"
    assert tb.put_varint(-1) == b"\xff" * 9 + b"\x01"   # protobuf int64
    # BundleHeaderProto {num_shards: 1, endianness: LITTLE (default), version {producer: 1}}
    assert tb.encode_header(1) == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    # BundleEntryProto for a float32 [2, 3] tensor at offset 0 (not emitted), 24 bytes, masked crc 0x11223344
    e = tb.encode_entry(np.float32, (2, 3), 0, 24, 0x11223344)
    assert e == bytes([0x08, 0x01,                                  # dtype = DT_FLOAT
                       0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x03,   # shape {dim {size 2} dim {size 3}}
                       0x28, 0x18,                                  # size = 24
                       0x35, 0x44, 0x33, 0x22, 0x11])               # crc32c fixed32
    d = tb.decode_entry(e)
    assert (d["dtype"], d["shape"], d["offset"], d["size"], d["crc32c"]) == (1, [2, 3], 0, 24, 0x11223344)
    # scalar int32 at offset 24: empty shape message is present, offset is emitted
    e = tb.encode_entry(np.int32, (), 24, 4, 7)
    assert e == bytes([0x08, 0x03, 0x12, 0x00, 0x20, 0x18, 0x28, 0x04, 0x35, 0x07, 0, 0, 0])
    assert tb.decode_entry(e)["shape"] == []


def _hand_block(entries):
    """One table block written out longhand: every entry a restart-relative record, one restart point at 0."""
    out = bytearray()
    last = b""
    for k, v in entries:
        shared = 0
        while shared < min(len(k), len(last)) and k[shared] == last[shared]:
            shared += 1
        out += bytes([shared, len(k) - shared, len(v)]) + k[shared:] + v   # all lengths < 128 here: 1-byte varints
        last = k
    out += struct.pack("<II", 0, 1)
    return bytes(out)


def _trailer(block):
    return b"\x00" + struct.pack("<I", tb.crc_mask(tb.crc32c(block + b"\x00")))


def test_writer_matches_hand_assembled_bundle(lib, tmp_path):
    """Two tensors, one data block.  Expected bytes assembled here from the format documents, not by the writer."""
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    g = np.array(41, dtype=np.int32)
    tb.write_bundle(str(tmp_path / "m"), {"PRE": a, "GLOBAL_STEP": g})
    data = (tmp_path / "m.data-00000-of-00001").read_bytes()
    assert data == g.tobytes() + a.tobytes()                      # key order: "GLOBAL_STEP" < "PRE"
    e_g = tb.encode_entry(np.int32, (), 0, 4, tb.crc_mask(tb.crc32c(g.tobytes())))
    e_a = tb.encode_entry(np.float32, (2, 3), 4, 24, tb.crc_mask(tb.crc32c(a.tobytes())))
    blk = _hand_block([(b"", tb.encode_header(1)), (b"GLOBAL_STEP", e_g), (b"PRE", e_a)])
    meta = struct.pack("<II", 0, 1)                               # empty metaindex block
    off_meta = len(blk) + 5
    off_idx = off_meta + len(meta) + 5
    # index block: one entry, key = short successor of "PRE" = "Q", value = handle(offset 0, size len(blk))
    idx = _hand_block([(b"Q", tb.put_varint(0) + tb.put_varint(len(blk)))])
    footer = tb.put_varint(off_meta) + tb.put_varint(len(meta)) + tb.put_varint(off_idx) + tb.put_varint(len(idx))
    footer += b"\x00" * (40 - len(footer)) + bytes([0x57, 0xFB, 0x80, 0x8B, 0x24, 0x75, 0x47, 0xDB])
    expect = blk + _trailer(blk) + meta + _trailer(meta) + idx + _trailer(idx) + footer
    assert (tmp_path / "m.index").read_bytes() == expect
    back = tb.read_bundle(str(tmp_path / "m"))
    assert back["PRE"].dtype == np.float32 and np.array_equal(back["PRE"], a)
    assert back["GLOBAL_STEP"].shape == () and int(back["GLOBAL_STEP"]) == 41


@pytest.mark.parametrize("block_size", [64, 400, 262144])
def test_round_trip_across_blocks_and_restarts(lib, tmp_path, block_size):
    rng = np.random.default_rng(1)
    tensors = {}
    for b in range(3):
        for bl in range(10):                                       # 30 layers x 6 keys: > 16 keys per block -> restarts
            sfx = "_%d_%d" % (b, bl)
            tensors["SIGNAL" + sfx] = rng.normal(size=(2, 4, 4)).astype(np.float32)
            tensors["SIGNAL_BIAS" + sfx] = rng.normal(size=(4,)).astype(np.float32)
            tensors["SAVE_%d%s" % (2 ** bl, sfx)] = rng.normal(size=(2, 2 ** bl, 4)).astype(np.float32)
    tensors["GLOBAL_STEP"] = np.array(7, np.int32)
    tensors["ckpt_position"] = np.array(2 ** 40 + 5, np.int64)
    tb.write_bundle(str(tmp_path / "r"), tensors, block_size=block_size)
    raw = (tmp_path / "r.index").read_bytes()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and tb.is_table_file(str(tmp_path / "r.index"))
    items = tb.read_table(raw)
    keys = [k for k, _ in items]
    assert keys == sorted(keys) and keys[0] == b"" and len(keys) == len(tensors) + 1
    back = tb.read_bundle(str(tmp_path / "r"))
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    # corruption is detected: flip one payload byte
    p = tmp_path / "r.data-00000-of-00001"
    blob = bytearray(p.read_bytes())
    blob[10] ^= 0x40
    p.write_bytes(bytes(blob))
    with pytest.raises(IOError):
        tb.read_bundle(str(tmp_path / "r"))


def test_index_separators_follow_leveldb_comparator():
    assert tb._shortest_separator(b"abcd", b"abzz") == b"abd"
    assert tb._shortest_separator(b"abc", b"abd") == b"abc"        # adjacent: cannot shorten
    assert tb._shortest_separator(b"ab", b"abc") == b"ab"          # prefix: unchanged
    assert tb._short_successor(b"PRE") == b"Q"
    assert tb._short_successor(b"\xff\xffa") == b"\xff\xffb"


def test_checkpoint_class_writes_tf_bundles_and_reads_legacy_json(lib, tmp_path):
    """reference ckpt.py:54-62,65-81 through the TF container; JSON-index checkpoints of earlier builds still load."""
    import json
    import zlib
    from lb_wavenet_b200 import ckpt
    store = {"PRE": np.arange(12, dtype=np.float32).reshape(3, 4), "GLOBAL_STEP": np.array(3, np.int32)}

    def var(k):
        return ckpt.Variable(k, store[k].shape, store[k].dtype, lambda: store[k], lambda v: store.__setitem__(k, v))

    c = ckpt.Checkpoint(str(tmp_path / "run.net"), 2, 0)
    c.add_saveable_objects({k: var(k) for k in store})
    pfx = c.save(10)
    assert tb.is_table_file(pfx + ".index")
    saved = {k: v.copy() for k, v in store.items()}
    store["PRE"] = np.zeros((3, 4), np.float32)
    c.restore(pfx)
    assert np.array_equal(store["PRE"], saved["PRE"]) and int(store["GLOBAL_STEP"]) == 3
    # legacy container
    leg = str(tmp_path / "old.net-5")
    raw = saved["PRE"].tobytes()
    with open(leg + ".data-00000-of-00001", "wb") as f:
        f.write(raw)
    with open(leg + ".index", "w") as f:
        json.dump(dict(format="lb-wavenet-b200/1", tensors={"PRE": dict(dtype="float32", shape=[3, 4], offset=0,
                  size=len(raw), crc32=zlib.crc32(raw) & 0xFFFFFFFF)}), f)
    assert np.array_equal(ckpt.read_checkpoint(leg)["PRE"], saved["PRE"])


# ---- independent statements of the pieces, written by the TensorFlow team ---------------------------------------------
# TensorFlow itself cannot be installed here (Python 3.12, no network), but TensorBoard -- which is -- ships Google's own
# copies of the TensorFlow protos (tensor_shape.proto, types.proto, versions.proto) and a pure-Python twin of TF's
# crc32c / masked crc (tensorboard/compat/tensorflow_stub/pywrap_tensorflow.py).  tensor_bundle.proto is not among
# them, so BundleHeaderProto / BundleEntryProto are declared below from the published .proto and compiled by Google's
# protobuf runtime; what is pinned is (a) the wire encoding of this file's hand-rolled protobuf writer, byte for byte,
# against that runtime, (b) the nested messages and enums against TensorBoard's generated classes, (c) the checksum.
# What this does NOT prove: that a file written by tf.train.Saver parses (the SSTable container has no second
# implementation in this image); INTEGRATION.md says so.

def _bundle_proto_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2  # noqa: F401  (register dependencies)
    pool = descriptor_pool.Default()
    name = "lbw_test/tensor_bundle.proto"
    try:
        fd = pool.FindFileByName(name)
    except KeyError:
        f = descriptor_pb2.FileDescriptorProto(name=name, package="lbw_test", syntax="proto3")
        f.dependency.extend(["tensorboard/compat/proto/tensor_shape.proto", "tensorboard/compat/proto/types.proto",
                             "tensorboard/compat/proto/versions.proto"])
        F = descriptor_pb2.FieldDescriptorProto
        h = f.message_type.add(name="BundleHeaderProto")
        e_ = h.enum_type.add(name="Endianness")
        e_.value.add(name="LITTLE", number=0)
        e_.value.add(name="BIG", number=1)
        h.field.add(name="num_shards", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
        h.field.add(name="endianness", number=2, type=F.TYPE_ENUM, label=F.LABEL_OPTIONAL,
                    type_name=".lbw_test.BundleHeaderProto.Endianness")
        h.field.add(name="version", number=3, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".tensorboard.VersionDef")
        e = f.message_type.add(name="BundleEntryProto")
        e.field.add(name="dtype", number=1, type=F.TYPE_ENUM, label=F.LABEL_OPTIONAL, type_name=".tensorboard.DataType")
        e.field.add(name="shape", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL,
                    type_name=".tensorboard.TensorShapeProto")
        e.field.add(name="shard_id", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
        e.field.add(name="offset", number=4, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
        e.field.add(name="size", number=5, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
        e.field.add(name="crc32c", number=6, type=F.TYPE_FIXED32, label=F.LABEL_OPTIONAL)
        fd = pool.Add(f)
    return (message_factory.GetMessageClass(fd.message_types_by_name["BundleHeaderProto"]),
            message_factory.GetMessageClass(fd.message_types_by_name["BundleEntryProto"]))


def test_proto_wire_format_against_googles_runtime_and_tensorboards_tf_protos():
    pytest.importorskip("tensorboard")
    from tensorboard.compat.proto import types_pb2
    Header, Entry = _bundle_proto_classes()
    # enum values this file hard-codes (types.proto)
    assert tb.DT_OF[np.dtype(np.float32)] == types_pb2.DT_FLOAT and tb.DT_OF[np.dtype(np.int32)] == types_pb2.DT_INT32
    assert tb.DT_OF[np.dtype(np.int64)] == types_pb2.DT_INT64
    h = Header.FromString(tb.encode_header(1))
    assert h.num_shards == 1 and h.endianness == 0 and h.version.producer == 1
    ref = Header(num_shards=1)
    ref.version.producer = 1
    assert ref.SerializeToString(deterministic=True) == tb.encode_header(1)
    for dtype, shape, off, size, crc, shard in ((np.float32, (2, 32, 32), 4096, 8192, 0xdeadbeef, 0),
                                                (np.int32, (), 0, 4, 1, 0), (np.int64, (10,), 123456789012, 80, 0, 0),
                                                (np.float32, (377, 17), 1 << 33, 25636, 0xffffffff, 0)):
        mine = tb.encode_entry(np.dtype(dtype), shape, off, size, crc, shard)
        e = Entry.FromString(mine)
        assert e.dtype == tb.DT_OF[np.dtype(dtype)] and [d.size for d in e.shape.dim] == list(shape)
        assert (e.offset, e.size, e.crc32c, e.shard_id) == (off, size, crc, shard)
        ref = Entry(dtype=tb.DT_OF[np.dtype(dtype)], offset=off, size=size, crc32c=crc, shard_id=shard)
        ref.shape.SetInParent()   # TF calls mutable_shape(): the (possibly empty) message is always present
        for n in shape:
            ref.shape.dim.add(size=n)
        assert ref.SerializeToString(deterministic=True) == mine, (dtype, shape)
        assert tb.decode_entry(ref.SerializeToString())["shape"] == list(shape)


def test_crc32c_and_mask_against_tensorboards_twin_of_tfs_checksum():
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 64, 1000, 4097):
        buf = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tb.crc32c(buf) == pw.crc32c(buf), n
        assert tb.crc_mask(tb.crc32c(buf)) == pw.masked_crc32c(buf), n
        assert tb.crc_unmask(pw.masked_crc32c(buf)) == pw.crc32c(buf)
