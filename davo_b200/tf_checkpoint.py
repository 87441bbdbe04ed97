"""Reader for TensorFlow checkpoints in the tensor-bundle ("V2") format -- what
``tf.train.Saver.save`` wrote for the reference's models (``model-<step>.index`` +
``model-<step>.data-00000-of-00001``; restored at reference ``test_kitti_pose.py:129-131``).

No TensorFlow needed.  The format, restated from TensorFlow's public sources
(``tensorflow/core/util/tensor_bundle``, ``tensorflow/core/lib/io/format.{h,cc}``,
``tensorflow/core/protobuf/tensor_bundle.proto``):

* ``<prefix>.index`` is a LevelDB-style sorted table: data blocks, a metaindex block, an index
  block and a 48-byte footer (two varint BlockHandles ``(offset, size)``, padding, the magic
  ``0xdb4775248b80fb57`` little-endian).  Every block is followed by a 5-byte trailer
  (compression type, masked crc32c).  A block is a run of prefix-compressed entries
  ``(shared, unshared, value_len, key_delta, value)`` plus a restart array.
* The entry with the empty key holds a ``BundleHeaderProto`` (num_shards, endianness, version);
  every other key is a variable name whose value is a ``BundleEntryProto``: dtype, shape,
  shard_id, offset, size, crc32c of the bytes in ``<prefix>.data-<shard>-of-<shards>``.

Only what the pose path needs is handled: uncompressed blocks (TensorFlow writes bundle indexes
without compression), little-endian float32 tensors, unsliced variables.  Anything else raises
``ValueError`` naming what was found.  This module could not be tried on a checkpoint produced
by TensorFlow in the build container (no TensorFlow, no network): it is checked against a writer
of the same format in ``tests/test_host.py`` and says so.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DT_FLOAT = 1


# ------------------------------------------------------------------ crc32c (Castagnoli), masked --
def _crc32c_table():
    poly = 0x82F63B78
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = _crc32c_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TAB
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------- varints, protobuf --
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("tf_checkpoint: truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise ValueError("tf_checkpoint: varint too long")


def _proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """(field number, wire type, value) of one protobuf message; value is int or bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("tf_checkpoint: protobuf wire type %d not expected here" % wt)
        yield field, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf: bytes) -> List[int]:
    dims = []
    for field, _wt, v in _proto_fields(buf):
        if field == 2:                                   # repeated Dim dim = 2
            size = 0
            for f2, _w2, v2 in _proto_fields(v):
                if f2 == 1:                              # int64 size = 1
                    size = _signed64(v2)
            dims.append(size)
        elif field == 3 and v:                           # unknown_rank
            raise ValueError("tf_checkpoint: tensor of unknown rank")
    return dims


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, _wt, v in _proto_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            e["shape"] = _parse_shape(v)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = _signed64(v)
        elif field == 5:
            e["size"] = _signed64(v)
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["sliced"] = True
    return e


def _parse_header(buf: bytes) -> dict:
    h = {"num_shards": 1, "endianness": 0}
    for field, _wt, v in _proto_fields(buf):
        if field == 1:
            h["num_shards"] = v
        elif field == 2:
            h["endianness"] = v
    return h


# --------------------------------------------------------------------------------------- the table --
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + 5 > len(data):
        raise ValueError("tf_checkpoint: block handle (%d, %d) runs past the end of the index file" % (offset, size))
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("tf_checkpoint: compressed index block (type %d); only uncompressed tables are read" % ctype)
    if verify:
        want = struct.unpack_from("<I", data, offset + size + 1)[0]
        if masked_crc32c(data[offset:offset + size + 1]) != want:
            raise ValueError("tf_checkpoint: index block at %d fails its crc32c" % offset)
    return block


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise ValueError("tf_checkpoint: block too short")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise ValueError("tf_checkpoint: bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > limit:
            raise ValueError("tf_checkpoint: corrupt block entry")
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(index_path: str, verify_crc: bool = True) -> Tuple[dict, Dict[str, dict]]:
    """(bundle header, {variable name: entry}) of ``<prefix>.index``."""
    with open(index_path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("tf_checkpoint: %s is not a tensor-bundle index (bad table magic)" % index_path)
    footer = data[len(data) - 48:]
    _mo, p = _varint(footer, 0)
    _ms, p = _varint(footer, p)
    io, p = _varint(footer, p)
    isz, p = _varint(footer, p)
    header, entries = None, {}
    for _sep, handle in _block_entries(_read_block(data, io, isz, verify_crc)):
        bo, q = _varint(handle, 0)
        bs, q = _varint(handle, q)
        for key, value in _block_entries(_read_block(data, bo, bs, verify_crc)):
            if key == b"":
                header = _parse_header(value)
            else:
                entries[key.decode("utf-8")] = _parse_entry(value)
    if header is None:
        raise ValueError("tf_checkpoint: %s has no bundle header entry" % index_path)
    if header["endianness"] != 0:
        raise ValueError("tf_checkpoint: big-endian bundle")
    return header, entries


def read_checkpoint(prefix: str, name_filter=None, verify_crc: bool = True,
                    verify_data_crc: bool = False) -> Dict[str, np.ndarray]:
    """{variable name: float32 ndarray} of the checkpoint ``prefix`` (the string given to
    ``saver.restore``, e.g. ``.../model-1600000``).  Non-float variables (``global_step`` ...)
    are skipped; ``name_filter(name) -> bool`` selects variables.  ``verify_crc`` checks the index
    blocks; ``verify_data_crc`` also the tensor bytes (pure Python: seconds per MB)."""
    header, entries = read_index(prefix + ".index", verify_crc)
    shards: Dict[int, bytes] = {}
    out: Dict[str, np.ndarray] = {}
    for name, e in sorted(entries.items()):
        if name_filter is not None and not name_filter(name):
            continue
        if e["dtype"] != DT_FLOAT:
            continue
        if e["sliced"]:
            raise ValueError("tf_checkpoint: variable %s is stored in slices (partitioned variable)" % name)
        sid = e["shard_id"]
        if sid not in shards:
            path = "%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"])
            with open(path, "rb") as f:
                shards[sid] = f.read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if len(raw) != e["size"] or e["size"] != 4 * n:
            raise ValueError("tf_checkpoint: variable %s: %d bytes for shape %s" % (name, len(raw), e["shape"]))
        if verify_data_crc and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise ValueError("tf_checkpoint: variable %s fails its crc32c" % name)
        out[name] = np.frombuffer(raw, dtype="<f4").reshape(e["shape"]).copy()
    return out


def is_checkpoint_prefix(path: str) -> bool:
    return os.path.isfile(path + ".index")


def pose_variables(name: str) -> bool:
    """The trainable variables of the pose path (reference test_kitti_pose.py:129 restores
    ``tf.trainable_variables()``): everything under pose_exp_net/, without optimizer slots."""
    return name.startswith("pose_exp_net/") and "/Adam" not in name and not name.endswith("/ExponentialMovingAverage")
