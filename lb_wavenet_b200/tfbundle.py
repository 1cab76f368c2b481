"""TensorFlow tensor-bundle ("V2 checkpoint") files without TensorFlow.

The reference saves and restores through ``tf.train.Saver`` (reference ckpt.py:41,54-62,65-81), whose on-disk
product per prefix is

    <prefix>.index                  an SSTable (LevelDB table format): key "" -> BundleHeaderProto, every tensor
                                    name -> BundleEntryProto (dtype, shape, shard, offset, size, masked crc32c)
    <prefix>.data-00000-of-00001    the raw little-endian tensor bytes, back to back, in key order
    <prefix>.meta                   a MetaGraphDef (only its presence is checked, ckpt.py:70-76)

TensorFlow itself is not installable here, so this module restates the published formats:
LevelDB's table_format.md (blocks with prefix-compressed entries and a restart array, 5-byte block trailer, 48-byte
footer with the magic 0xdb4775248b80fb57), TensorFlow's tensor_bundle.proto / tensor_shape.proto / versions.proto and
crc32c masking (rotate right 15, add 0xa282ead8).  The writer emits exactly what TF's BundleWriter emits for a
single-shard bundle (uncompressed blocks, restart interval 16, index-block separators shortened like LevelDB's
BytewiseComparator); the reader accepts any uncompressed table.  tests/test_tfbundle.py pins the byte stream of a
hand-assembled bundle, the RFC 3720 CRC-32C vectors and round trips across block boundaries.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, List, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER = 5
RESTART_INTERVAL = 16
BLOCK_SIZE = 262144  # tensorflow/core/lib/io/table_options.h
MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DT_OF = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("uint8"): 4,
         np.dtype("int16"): 5, np.dtype("int8"): 6, np.dtype("int64"): 9, np.dtype("bool"): 10,
         np.dtype("uint16"): 17, np.dtype("float16"): 19, np.dtype("uint32"): 22, np.dtype("uint64"): 23}
NP_OF = {v: k for k, v in DT_OF.items()}


# ---- crc32c ----------------------------------------------------------------------------------------
def crc32c(data, crc: int = 0) -> int:
    """CRC-32C of a bytes-like object through the C ABI (wn_crc32c, slicing-by-8)."""
    from . import _lib
    b = bytes(data)
    return int(_lib.load().wn_crc32c(crc, b, len(b))) & 0xFFFFFFFF


def crc_mask(crc: int) -> int:
    """tensorflow/core/lib/hash/crc32c.h Mask: rotate right by 15 bits, add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + MASK_DELTA) & 0xFFFFFFFF


def crc_unmask(masked: int) -> int:
    rot = (masked - MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format ----------------------------------------------------------------
def put_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64  # two's complement, as protobuf encodes negative int64
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _fields(buf: bytes):
    """Yield (field number, wire type, value) of a serialized message; value is int or bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = get_varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield fn, wt, v


def encode_header(num_shards: int = 1) -> bytes:
    """BundleHeaderProto {num_shards = 1; endianness = LITTLE (0, default: not emitted); version {producer: 1}}."""
    return b"\x08" + put_varint(num_shards) + b"\x1a\x02\x08\x01"


def encode_entry(dtype: np.dtype, shape: Iterable[int], offset: int, size: int, masked_crc: int, shard_id: int = 0) -> bytes:
    """BundleEntryProto: dtype = 1, shape = 2 (TensorShapeProto: repeated dim = 2 {size = 1}), shard_id = 3,
    offset = 4, size = 5, crc32c = 6 (fixed32).  proto3: zero-valued scalars are not emitted, the (possibly empty)
    shape message always is because TF calls mutable_shape()."""
    dims = b"".join(b"\x12" + put_varint(len(d)) + d for d in (b"\x08" + put_varint(int(n)) for n in shape))
    out = b"\x08" + put_varint(DT_OF[np.dtype(dtype)]) + b"\x12" + put_varint(len(dims)) + dims
    if shard_id:
        out += b"\x18" + put_varint(shard_id)
    if offset:
        out += b"\x20" + put_varint(offset)
    if size:
        out += b"\x28" + put_varint(size)
    if masked_crc:  # proto3: a zero fixed32 is not emitted either (1 in 2^32 for a masked crc)
        out += b"\x35" + struct.pack("<I", masked_crc)
    return out


def decode_entry(buf: bytes) -> dict:
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=0, slices=0)
    for fn, wt, v in _fields(buf):
        if fn == 1:
            e["dtype"] = v
        elif fn == 2:
            for f2, _, dim in _fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, x in _fields(dim):
                        if f3 == 1:
                            size = x - (1 << 64) if x >= (1 << 63) else x
                    e["shape"].append(size)
        elif fn == 3:
            e["shard_id"] = v
        elif fn == 4:
            e["offset"] = v
        elif fn == 5:
            e["size"] = v
        elif fn == 6:
            e["crc32c"] = v
        elif fn == 7:
            e["slices"] += 1
    return e


# ---- LevelDB table ------------------------------------------------------------------------------------
class _BlockBuilder:
    def __init__(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.counter < RESTART_INTERVAL:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value))
        self.buf += key[shared:] + value
        self.last_key = key
        self.counter += 1

    def size_estimate(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self) -> bool:
        return not self.buf

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start: bytes, limit: bytes) -> bytes:
    """LevelDB BytewiseComparator::FindShortestSeparator."""
    m = min(len(start), len(limit))
    d = 0
    while d < m and start[d] == limit[d]:
        d += 1
    if d < m and start[d] < 0xFF and start[d] + 1 < limit[d]:
        return start[:d] + bytes([start[d] + 1])
    return start


def _short_successor(key: bytes) -> bytes:
    """LevelDB BytewiseComparator::FindShortSuccessor."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _handle(offset: int, size: int) -> bytes:
    return put_varint(offset) + put_varint(size)


def build_table(items: List[Tuple[bytes, bytes]], block_size: int = BLOCK_SIZE) -> bytes:
    """items: (key, value) in strictly increasing bytewise key order -> the bytes of an (uncompressed) table file."""
    out = bytearray()
    index = _BlockBuilder()

    def write_block(contents: bytes) -> Tuple[int, int]:
        off = len(out)
        out.extend(contents)
        out.append(0)  # kNoCompression
        out.extend(struct.pack("<I", crc_mask(crc32c(contents + b"\x00"))))
        return off, len(contents)

    blk = _BlockBuilder()
    pending = None  # (last key of the finished block, its handle): the index entry waits for the next key
    for key, value in items:
        if pending is not None:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        blk.add(key, value)
        if blk.size_estimate() >= block_size:
            off, size = write_block(blk.finish())
            pending = (blk.last_key, _handle(off, size))
            blk = _BlockBuilder()
    if not blk.empty():
        off, size = write_block(blk.finish())
        pending = (blk.last_key, _handle(off, size))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    meta_off, meta_size = write_block(_BlockBuilder().finish())
    idx_off, idx_size = write_block(index.finish())
    footer = _handle(meta_off, meta_size) + _handle(idx_off, idx_size)
    footer += b"\x00" * (FOOTER_LEN - 8 - len(footer))
    footer += struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    return bytes(out)


def _read_block(data: bytes, off: int, size: int, verify: bool = True) -> bytes:
    contents = data[off:off + size]
    ctype = data[off + size]
    crc = struct.unpack_from("<I", data, off + size + 1)[0]
    if verify and crc_unmask(crc) != crc32c(data[off:off + size + 1]):
        raise IOError("table block checksum mismatch at offset %d" % off)
    if ctype != 0:
        raise NotImplementedError("compressed table block (type %d): TF's BundleWriter writes uncompressed blocks" % ctype)
    return contents


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = get_varint(block, pos)
        non_shared, pos = get_varint(block, pos)
        vlen, pos = get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(data: bytes) -> List[Tuple[bytes, bytes]]:
    if len(data) < FOOTER_LEN or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("not a LevelDB/TensorFlow table file (bad magic)")
    footer = data[len(data) - FOOTER_LEN:]
    _, p = get_varint(footer, 0)
    _, p = get_varint(footer, p)
    idx_off, p = get_varint(footer, p)
    idx_size, p = get_varint(footer, p)
    out = []
    for _, h in _block_entries(_read_block(data, idx_off, idx_size)):
        off, q = get_varint(h, 0)
        size, q = get_varint(h, q)
        out.extend(_block_entries(_read_block(data, off, size)))
    return out


def is_table_file(path: str) -> bool:
    try:
        with open(path, "rb") as f:
            f.seek(-8, 2)
            return struct.unpack("<Q", f.read(8))[0] == TABLE_MAGIC
    except (OSError, struct.error):
        return False


# ---- bundles ----------------------------------------------------------------------------------------
def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = BLOCK_SIZE, suffix: str = "") -> None:
    """<prefix>.index + <prefix>.data-00000-of-00001 (``suffix`` is appended to both file names: atomic rename by the
    caller)."""
    items = [(b"", encode_header(1))]
    off = 0
    with open(prefix + ".data-00000-of-00001" + suffix, "wb") as f:
        for key in sorted(tensors, key=lambda k: k.encode()):
            arr = np.asarray(tensors[key])
            arr = np.array(arr, dtype=arr.dtype.newbyteorder("<"), order="C")
            raw = arr.tobytes()
            f.write(raw)
            items.append((key.encode(), encode_entry(arr.dtype.newbyteorder("="), arr.shape, off, len(raw), crc_mask(crc32c(raw)))))
            off += len(raw)
    with open(prefix + ".index" + suffix, "wb") as f:
        f.write(build_table(items, block_size))


def read_bundle(prefix: str, verify: bool = True) -> Dict[str, np.ndarray]:
    with open(prefix + ".index", "rb") as f:
        entries = read_table(f.read())
    if not entries or entries[0][0] != b"":
        raise ValueError("tensor bundle %s.index lacks the header entry" % prefix)
    hdr = {fn: v for fn, _, v in _fields(entries[0][1])}
    n_shards = hdr.get(1, 0)
    if hdr.get(2, 0) != 0:
        raise NotImplementedError("big-endian tensor bundle")
    shards = {}
    out = {}
    for key, val in entries[1:]:
        e = decode_entry(val)
        if e["slices"]:
            raise NotImplementedError("partitioned (sliced) variable %r" % key)
        if e["dtype"] not in NP_OF:
            raise NotImplementedError("tensor %r has unsupported dtype enum %d" % (key, e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            with open("%s.data-%05d-of-%05d" % (prefix, sid, n_shards), "rb") as f:
                shards[sid] = f.read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise IOError("tensor %r: data shard truncated" % key)
        if verify and crc_unmask(e["crc32c"]) != crc32c(raw):
            raise IOError("tensor %r: crc32c mismatch" % key)
        out[key.decode()] = np.frombuffer(raw, dtype=NP_OF[e["dtype"]].newbyteorder("<")).reshape(e["shape"]).astype(NP_OF[e["dtype"]])
    return out
