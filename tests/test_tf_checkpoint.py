"""CPU: the reader of the reference's TensorFlow checkpoints (rnn-speech_b200/tf_checkpoint.py).

* tests/golden/tf_bundle/acousticmodel.ckpt.index is the index file of the checkpoint shipped with the reference
  (trained_models/english/acoustic/, 585 bytes; its .data file is a git-lfs pointer and its weights live in the .meta);
* a synthetic bundle written here with the same table layout exercises the data path;
* when /root/reference is present the shipped .meta is read too (its 12 variables, the 3x1024 model).
"""
import importlib.util
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("tf_checkpoint", os.path.join(ROOT, "rnn-speech_b200", "tf_checkpoint.py"))
tfck = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tfck)

SHIPPED = "/root/reference/trained_models/english/acoustic/acousticmodel.ckpt"


def test_shipped_index_lists_the_twelve_variables():
    entries, shards = tfck.read_bundle_index(os.path.join(ROOT, "tests", "golden", "tf_bundle", "acousticmodel.ckpt"))
    assert shards == 1 and len(entries) == 12
    assert entries["Input_Layer/input_w"]["shape"] == (120, 1024)
    assert entries["Output_layer/output_w"]["shape"] == (1024, 80)
    for l in range(3):
        k = entries["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % l]
        assert k["shape"] == (2048, 4096) and k["size"] == 2048 * 4096 * 4 and k["dtype"] == 1
    assert entries["global_step"]["dtype"] == 3 and entries["global_step"]["shape"] == ()
    # offsets tile the data shard without gaps, in key order
    pos = 0
    for name in sorted(entries):
        assert entries[name]["offset"] == pos
        pos += entries[name]["size"]
    assert pos == 101536072                    # the size the git-lfs pointer announces for the data shard


def _varint(v):
    out = b""
    while True:
        b = v & 0x7F
        v >>= 7
        out += bytes([b | (0x80 if v else 0)])
        if not v:
            return out


def _write_bundle(prefix, tensors):
    """Minimal tensor-bundle writer (one uncompressed data block, no checksums): same layout TF writes."""
    dt = {np.dtype("float32"): 1, np.dtype("int32"): 3, np.dtype("int64"): 9}
    data, items = b"", [(b"", b"\x08\x01\x1a\x02\x08\x01")]
    for name in sorted(tensors):
        arr = np.ascontiguousarray(tensors[name])
        shape = b"".join(b"\x12" + _varint(len(b"\x08" + _varint(d))) + b"\x08" + _varint(d) for d in arr.shape)
        ent = b"\x08" + _varint(dt[arr.dtype]) + b"\x12" + _varint(len(shape)) + shape
        ent += b"\x20" + _varint(len(data)) + b"\x28" + _varint(arr.nbytes) + b"\x35" + b"\0\0\0\0"
        items.append((name.encode(), ent))
        data += arr.tobytes()

    def block(kvs):
        out, prev = b"", b""
        for k, v in kvs:
            shared = 0
            while shared < min(len(k), len(prev)) and k[shared] == prev[shared]:
                shared += 1
            out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
            prev = k
        return out + struct.pack("<II", 0, 1)
    blk = block(items)
    meta = block([])
    body = blk + b"\0" * 5
    meta_off = len(body)
    body += meta + b"\0" * 5
    idx = block([(items[-1][0] + b"\xff", _varint(0) + _varint(len(blk)))])
    idx_off = len(body)
    body += idx + b"\0" * 5
    footer = _varint(meta_off) + _varint(len(meta)) + _varint(idx_off) + _varint(len(idx))
    footer += b"\0" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)
    open(prefix + ".index", "wb").write(body + footer)
    open(prefix + ".data-00000-of-00001", "wb").write(data)


def test_bundle_round_trip_and_truncated_shard(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"Input_Layer/input_w": rng.standard_normal((5, 7)).astype(np.float32),
               "Input_Layer/input_b": rng.standard_normal(7).astype(np.float32),
               "rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel": rng.standard_normal((14, 28)).astype(np.float32),
               "global_step": np.array(67600, np.int32), "learning_rate": np.array(1.0781e-5, np.float32)}
    prefix = str(tmp_path / "acousticmodel.ckpt")
    _write_bundle(prefix, tensors)
    got = tfck.read_bundle(prefix)
    assert sorted(got) == sorted(tensors)
    for k in tensors:
        np.testing.assert_array_equal(got[k], tensors[k])
    values, source = tfck.load_reference_checkpoint(prefix)
    assert source == "bundle" and int(values["global_step"].reshape(-1)[0]) == 67600
    # a data shard shorter than the index says (the shipped git-lfs pointer) is reported, not mis-read
    open(prefix + ".data-00000-of-00001", "wb").write(b"version https://git-lfs.github.com/spec/v1\n")
    with pytest.raises(IOError):
        tfck.read_bundle(prefix)


def test_meta_initial_values_from_a_synthetic_graph(tmp_path):
    meta_graph_pb2 = pytest.importorskip("tensorboard.compat.proto.meta_graph_pb2")
    from tensorboard.util import tensor_util
    meta = meta_graph_pb2.MetaGraphDef()
    w = np.arange(12, dtype=np.float32).reshape(3, 4)
    for name, arr in (("Output_layer/output_w", w), ("global_step", np.array(7, np.int32))):
        node = meta.graph_def.node.add()
        node.name, node.op = name + "/initial_value", "Const"
        node.attr["value"].tensor.CopyFrom(tensor_util.make_tensor_proto(arr))
    other = meta.graph_def.node.add()
    other.name, other.op = "Output_layer/output_w", "VariableV2"
    path = str(tmp_path / "m.ckpt.meta")
    open(path, "wb").write(meta.SerializeToString())
    got = tfck.read_meta_initial_values(path)
    assert sorted(got) == ["Output_layer/output_w", "global_step"]
    np.testing.assert_array_equal(got["Output_layer/output_w"], w)
    values, source = tfck.load_reference_checkpoint(str(tmp_path / "m.ckpt"))        # no bundle at all -> .meta
    assert source == "meta" and int(values["global_step"].reshape(-1)[0]) == 7


@pytest.mark.skipif(not os.path.exists(SHIPPED + ".meta"), reason="the reference tree is not mounted")
def test_shipped_model_is_readable_from_its_meta_file():
    values, source = tfck.load_reference_checkpoint(SHIPPED)
    assert source == "meta"                                   # the shipped data shard is a git-lfs pointer
    assert values["rnn/multi_rnn_cell/cell_2/basic_lstm_cell/kernel"].shape == (2048, 4096)
    assert int(values["global_step"].reshape(-1)[0]) == 67600
    # gate order i, j, f, o: the forget-gate quarter of a trained bias is the most negative... the j quarter ~ 0
    bias = values["rnn/multi_rnn_cell/cell_0/basic_lstm_cell/bias"].reshape(4, 1024)
    assert abs(float(bias[1].mean())) < 0.01
    assert sum(int(np.prod(v.shape)) for k, v in values.items() if k not in ("global_step", "learning_rate")) == 25384016


def test_crc32c_known_answers_and_tf_written_block_trailers():
    # RFC 3720 check value, and the 32 zero bytes vector
    assert tfck.crc32c(b"123456789") == 0xE3069283
    assert tfck.crc32c(bytes(32)) == 0x8A9136AA
    assert tfck.crc32c(b"6789", tfck.crc32c(b"12345")) == 0xE3069283          # continuation
    # the masking rule and "block + type byte" coverage are TensorFlow's: every block of the index file that TF wrote
    # for the reference's shipped checkpoint verifies
    assert tfck.verify_table_checksums(os.path.join(ROOT, "tests", "golden", "tf_bundle", "acousticmodel.ckpt.index")) == 3


def test_written_bundle_reads_back_and_verifies(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {"Input_Layer/input_w": rng.standard_normal((5, 7)).astype(np.float32),
               "rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel": rng.standard_normal((14, 28)).astype(np.float32),
               "rnn/multi_rnn_cell/cell_0/basic_lstm_cell/bias": rng.standard_normal(28).astype(np.float32),
               "global_step": np.array(123, np.int32), "learning_rate": np.array(3e-4, np.float32)}
    for i in range(40):                                     # enough keys for several restart points
        tensors["extra/var_%02d" % i] = np.full((i % 3 + 1,), i, np.int64)
    prefix = str(tmp_path / "acousticmodel.ckpt-123")
    names = tfck.write_bundle(prefix, tensors)
    assert names == sorted(tensors)
    assert tfck.verify_table_checksums(prefix + ".index") == 3
    entries, shards = tfck.read_bundle_index(prefix)
    assert shards == 1 and sorted(entries) == sorted(tensors)
    got = tfck.read_bundle(prefix)
    for k in tensors:
        assert got[k].shape == np.asarray(tensors[k]).shape and got[k].dtype == np.asarray(tensors[k]).dtype
        np.testing.assert_array_equal(got[k], tensors[k])
    # the per-tensor checksum in the entry is the masked CRC-32C of the payload
    raw = open(prefix + ".index", "rb").read()
    w = tensors["Input_Layer/input_w"].tobytes()
    assert struct.pack("<I", tfck.mask_crc(tfck.crc32c(w))) in raw
