# coding=utf-8
"""Reader for the reference's TensorFlow checkpoints (models/AcousticModel.py:483-527 saves /
restores with tf.train.Saver, i.e. the "tensor bundle" format):

  <prefix>.index                 an SSTable (LevelDB table format, uncompressed blocks) mapping the empty key to a
                                 BundleHeaderProto and every variable name to a BundleEntryProto
                                 {dtype, shape, shard_id, offset, size, crc32c}
  <prefix>.data-0000k-of-0000N   the raw little-endian tensor bytes
  <prefix>.meta                  the MetaGraphDef.  The checkpoint shipped with the reference
                                 (trained_models/english/acoustic) embeds the trained values as
                                 `<variable>/initial_value` Const nodes, and its .data file is a git-lfs pointer, so
                                 the .meta is the only place the shipped weights can be read from.

No TensorFlow needed: the table and the two protos are parsed by hand; the .meta goes through the protobuf
definitions that ship with tensorboard (`tensorboard.compat.proto`).  Host-side glue, no arithmetic.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(buf, offset, size):
    """(key, value) pairs of one table block (prefix-compressed keys, restart array at the end)."""
    block = buf[offset:offset + size]
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path):
    """All (key, value) pairs of an uncompressed LevelDB-format table file, in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = buf[-48:]
    _, pos = _varint(footer, 0)             # metaindex handle: offset, size
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    out = []
    for _key, handle in _block_entries(buf, idx_off, idx_size):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        if buf[off + size] != 0:
            raise ValueError("%s: compressed table blocks are not supported" % path)
        out.extend(_block_entries(buf, off, size))
    return out


def _parse_entry(value):
    """BundleEntryProto -> dict(dtype, shape, shard, offset, size)."""
    ent = {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0}
    pos = 0
    while pos < len(value):
        tag, pos = _varint(value, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(value, pos)
            if field == 1:
                ent["dtype"] = v
            elif field == 3:
                ent["shard"] = v
            elif field == 4:
                ent["offset"] = v
            elif field == 5:
                ent["size"] = v
        elif wire == 2:
            n, pos = _varint(value, pos)
            sub = value[pos:pos + n]
            pos += n
            if field == 2:                   # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
                dims, q = [], 0
                while q < len(sub):
                    t2, q = _varint(sub, q)
                    if t2 & 7 == 2:
                        m, q = _varint(sub, q)
                        dim, r = sub[q:q + m], 0
                        q += m
                        size = 0
                        while r < len(dim):
                            t3, r = _varint(dim, r)
                            if t3 & 7 == 0:
                                v3, r = _varint(dim, r)
                                if t3 >> 3 == 1:
                                    size = v3
                            elif t3 & 7 == 2:
                                k, r = _varint(dim, r)
                                r += k
                        if t2 >> 3 == 2:
                            dims.append(size)
                    elif t2 & 7 == 0:
                        _, q = _varint(sub, q)
                ent["shape"] = tuple(dims)
        elif wire == 5:
            pos += 4
        elif wire == 1:
            pos += 8
        else:
            raise ValueError("unexpected wire type %d in a BundleEntryProto" % wire)
    return ent


def read_bundle_index(prefix):
    """name -> {dtype, shape, shard, offset, size} for every tensor of `<prefix>.index`; plus '' -> number of shards."""
    entries = {}
    num_shards = 1
    for key, value in read_table(prefix + ".index"):
        if key == b"":
            pos = 0
            while pos < len(value):                # BundleHeaderProto: int32 num_shards = 1
                tag, pos = _varint(value, pos)
                if tag & 7 == 0:
                    v, pos = _varint(value, pos)
                    if tag >> 3 == 1:
                        num_shards = v
                elif tag & 7 == 2:
                    n, pos = _varint(value, pos)
                    pos += n
                else:
                    break
            continue
        entries[key.decode("utf-8")] = _parse_entry(value)
    return entries, num_shards


def read_bundle(prefix, names=None):
    """name -> ndarray from `<prefix>.index` + `<prefix>.data-*`.  Raises a clear error when the data shard is missing
    or shorter than the index says (the reference's shipped .data file is a 134-byte git-lfs pointer)."""
    entries, num_shards = read_bundle_index(prefix)
    out = {}
    for name, ent in entries.items():
        if names is not None and name not in names:
            continue
        if ent["dtype"] not in _DTYPES:
            raise ValueError("tensor %s has unsupported dtype enum %d" % (name, ent["dtype"]))
        shard = "%s.data-%05d-of-%05d" % (prefix, ent["shard"], num_shards)
        if not os.path.exists(shard) or os.path.getsize(shard) < ent["offset"] + ent["size"]:
            raise IOError("%s does not hold tensor %s (%d bytes at offset %d): missing or truncated data shard"
                          % (shard, name, ent["size"], ent["offset"]))
        with open(shard, "rb") as fh:
            fh.seek(ent["offset"])
            raw = fh.read(ent["size"])
        out[name] = np.frombuffer(raw, dtype=_DTYPES[ent["dtype"]]).reshape(ent["shape"]).copy()
    return out


def read_meta_initial_values(meta_path):
    """name -> ndarray for every `<name>/initial_value` Const node of a MetaGraphDef (how the reference's shipped
    checkpoint carries its weights)."""
    try:
        from tensorboard.compat.proto import meta_graph_pb2
        from tensorboard.util import tensor_util
    except ImportError as exc:                   # pragma: no cover
        raise ImportError("reading a .meta file needs the protobuf definitions shipped with tensorboard: %s" % exc)
    meta = meta_graph_pb2.MetaGraphDef()
    with open(meta_path, "rb") as fh:
        meta.ParseFromString(fh.read())
    out = {}
    for node in meta.graph_def.node:
        if node.op == "Const" and node.name.endswith("/initial_value"):
            out[node.name[:-len("/initial_value")]] = np.array(tensor_util.make_ndarray(node.attr["value"].tensor))
    return out


def load_reference_checkpoint(prefix):
    """The variables of a reference checkpoint, from the tensor bundle when its data is there, else from the
    constants embedded in `<prefix>.meta`.  Returns (dict name -> ndarray, source string)."""
    try:
        return read_bundle(prefix), "bundle"
    except (IOError, OSError, ValueError) as bundle_error:
        meta = prefix + ".meta"
        if not os.path.exists(meta):
            raise
        values = read_meta_initial_values(meta)
        if not values:
            raise bundle_error
        return values, "meta"


# ------------------------------------------------------------------------------------------------------------------
# Writer: the same tensor-bundle layout, so that the reference (tf.train.Saver.restore, models/AcousticModel.py:489-499)
# can read what this library trained.  Checksums are CRC-32C, stored masked as TensorFlow / LevelDB do
# (rot-right 15 plus 0xa282ead8); the convention is pinned against the block trailers of the index file shipped
# with the reference (tests/test_tf_checkpoint.py).  TensorFlow itself is not available here: what is verified is the
# byte layout (round trip through the reader above) and every checksum rule against that TF-written file.
_CRC_TABLE = None


def crc32c(data, crc=0, accel=None):
    """CRC-32C of bytes-like `data` continuing from `crc`.  `accel(data_ptr, n, crc) -> crc` (the library's
    rs_crc32c) is used for large buffers when given; the pure-Python table walk otherwise."""
    global _CRC_TABLE
    if accel is not None and len(data) >= 4096:
        buf = np.frombuffer(data, dtype=np.uint8)
        return int(accel(buf.ctypes.data, buf.size, crc))
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        c = (c >> 8) ^ tab[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


_DTYPE_ENUM = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def _table_block(items, restart_interval=16):
    """One LevelDB table block: prefix-compressed entries, restart offsets, restart count."""
    out, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(key), len(prev)) and key[shared] == prev[shared]:
                shared += 1
        out += _enc_varint(shared) + _enc_varint(len(key) - shared) + _enc_varint(len(value)) + key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block):
    """block + compression type (0 = none) + masked CRC-32C of both."""
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


def write_bundle(prefix, tensors, accel=None):
    """Write `tensors` (name -> ndarray; float32 / float64 / int32 / int64) as `<prefix>.index` +
    `<prefix>.data-00000-of-00001`.  Returns the list of names written, in file order."""
    names = sorted(tensors)
    items = [(b"", b"\x08\x01\x1a\x02\x08\x01")]           # BundleHeaderProto: num_shards 1, version.producer 1
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in names:
            arr = np.asarray(tensors[name])
            if arr.ndim and not arr.flags.c_contiguous:      # (ascontiguousarray would turn a scalar into shape (1,))
                arr = np.ascontiguousarray(arr)
            if arr.dtype not in _DTYPE_ENUM:
                raise ValueError("tensor %s: dtype %s is not supported by the bundle writer" % (name, arr.dtype))
            raw = arr.tobytes()
            shape = b"".join(b"\x12" + _enc_varint(len(b"\x08" + _enc_varint(d))) + b"\x08" + _enc_varint(d)
                             for d in arr.shape)
            ent = b"\x08" + _enc_varint(_DTYPE_ENUM[arr.dtype]) + b"\x12" + _enc_varint(len(shape)) + shape
            if offset:
                ent += b"\x20" + _enc_varint(offset)
            ent += b"\x28" + _enc_varint(len(raw)) + b"\x35" + struct.pack("<I", mask_crc(crc32c(raw, accel=accel)))
            items.append((name.encode("utf-8"), ent))
            fh.write(raw)
            offset += len(raw)
    data_block = _table_block(items)
    body = _with_trailer(data_block)
    meta_off = len(body)
    meta_block = _table_block([])
    body += _with_trailer(meta_block)
    # index block: one entry whose key is >= the last key of the data block
    index_block = _table_block([(items[-1][0] + b"\x00", _enc_varint(0) + _enc_varint(len(data_block)))], 1)
    idx_off = len(body)
    body += _with_trailer(index_block)
    footer = _enc_varint(meta_off) + _enc_varint(len(meta_block)) + _enc_varint(idx_off) + _enc_varint(len(index_block))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    with open(prefix + ".index", "wb") as fh:
        fh.write(body + footer)
    return names


def verify_table_checksums(path):
    """Check the masked CRC-32C trailer of every block of a table file; returns the number of blocks checked."""
    buf = open(path, "rb").read()
    footer = buf[-48:]
    meta_off, pos = _varint(footer, 0)
    meta_size, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    handles = [(meta_off, meta_size), (idx_off, idx_size)]
    for _key, handle in _block_entries(buf, idx_off, idx_size):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        handles.append((off, size))
    for off, size in handles:
        want = struct.unpack_from("<I", buf, off + size + 1)[0]
        got = mask_crc(crc32c(buf[off:off + size + 1]))
        if want != got:
            raise ValueError("%s: block at %d: checksum %08x, expected %08x" % (path, off, got, want))
    return len(handles)
