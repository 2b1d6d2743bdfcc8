# -*-coding:utf-8 -*-
"""Reader (and minimal writer) of TensorFlow's tensor-bundle checkpoint format — the files the reference trains from and
serves with: `pretrain_model/ch_google/bert_model.ckpt.{index,data-00000-of-00001}` loaded by `load_bert_checkpoint`
(reference tools/train_utils.py:91-102) and `serving_model/<model>/1/variables/variables.{index,data-*}`.

TensorFlow is not installable here, so its on-disk format (tensorflow/core/util/tensor_bundle, third party) is restated:

  `<prefix>.index`   a LevelDB-style sorted string table (tensorflow/core/lib/io/table): data blocks of prefix-compressed
                     (shared, non_shared, value_len varint32 | key suffix | value) entries followed by a restart array,
                     each block trailed by 1 compression byte + masked crc32c; a 48-byte footer holds the metaindex and
                     index block handles (varint64 offset, size) and the magic 0xdb4775248b80fb57.  Key "" maps to a
                     BundleHeaderProto (num_shards, endianness, version), every other key is a variable name mapping to a
                     BundleEntryProto {1: dtype, 2: TensorShapeProto{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size,
                     6: crc32c (fixed32)}.
  `<prefix>.data-SSSSS-of-NNNNN`   raw little-endian tensor bytes at [offset, offset + size).

Only uncompressed blocks are read (TF's BundleWriter never compresses); snappy blocks raise.
"""
import os
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = 1, 2, 3, 4, 9, 10, 14, 19
_NP_OF = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_UINT8: np.uint8, DT_INT64: np.int64, DT_BOOL: np.bool_,
          DT_HALF: np.float16}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}


# ----------------------------------------------------------------------------- varints / protobuf wire format
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """[(field number, wire type, value)] of one protobuf message; value = int (varint / fixed) or bytes (length-delimited)."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = struct.unpack_from('<Q', buf, pos)[0], pos + 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val, pos = bytes(buf[pos:pos + n]), pos + n
        elif wt == 5:
            val, pos = struct.unpack_from('<I', buf, pos)[0], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type {}'.format(wt))
        out.append((field, wt, val))
    return out


def _parse_entry(buf):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': 0, 'sliced': False}
    for field, wt, val in _proto_fields(buf):
        if field == 1:
            e['dtype'] = val
        elif field == 2:                       # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1; string name = 2}
            for f2, _, dim in _proto_fields(val):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _proto_fields(dim):
                        if f3 == 1:
                            size = v3
                    e['shape'].append(size)
        elif field == 3:
            e['shard_id'] = val
        elif field == 4:
            e['offset'] = val
        elif field == 5:
            e['size'] = val
        elif field == 6:
            e['crc32c'] = val
        elif field == 7:
            e['sliced'] = True
    return e


# ----------------------------------------------------------------------------- crc32c (Castagnoli), TF's masked form
_CRC_TABLE = None


def crc32c(data, crc=0):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            tbl.append(c)
        _CRC_TABLE = tbl
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- table reader
def _read_block(buf, offset, size, verify=True):
    raw = buf[offset:offset + size]
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError('compressed table block (type {}): TF bundle writers do not compress'.format(ctype))
    if verify:
        stored = struct.unpack_from('<I', buf, offset + size + 1)[0]
        if stored != masked_crc32c(buf[offset:offset + size + 1]):
            raise ValueError('table block checksum mismatch at offset {}'.format(offset))
    return raw


def _block_entries(block):
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b'', []
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def read_bundle_index(index_path, verify=True):
    """`<prefix>.index` -> (header dict, OrderedDict name -> entry dict(dtype, shape, shard_id, offset, size, crc32c)),
    names in the table's (sorted) order."""
    with open(index_path, 'rb') as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError('{}: not a TensorFlow bundle index (bad table magic)'.format(index_path))
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    header, entries = {}, OrderedDict()
    for _, handle in _block_entries(_read_block(buf, idx_off, idx_size, verify)):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, val in _block_entries(_read_block(buf, off, size, verify)):
            if key == b'':
                for field, _, v in _proto_fields(val):
                    header[{1: 'num_shards', 2: 'endianness', 3: 'version'}.get(field, field)] = v
            else:
                entries[key.decode('utf-8')] = _parse_entry(val)
    header.setdefault('num_shards', 1)
    return header, entries


def load_tf_checkpoint(prefix, names=None, verify=False):
    """-> OrderedDict name -> numpy array, for `names` (default: every non-sliced tensor)."""
    header, entries = read_bundle_index(prefix + '.index')
    n = header.get('num_shards', 1)
    shards, out = {}, OrderedDict()
    for name, e in entries.items():
        if (names is not None and name not in names) or e['sliced']:
            continue
        if e['dtype'] not in _NP_OF:
            continue
        mm = shards.get(e['shard_id'])
        if mm is None:
            mm = shards[e['shard_id']] = np.memmap('{}.data-{:05d}-of-{:05d}'.format(prefix, e['shard_id'], n), dtype=np.uint8, mode='r')
        raw = mm[e['offset']:e['offset'] + e['size']]
        if len(raw) != e['size']:
            raise ValueError('{}: data shard shorter than the index says (an un-fetched LFS pointer?)'.format(name))
        if verify and masked_crc32c(bytes(raw)) != e['crc32c']:
            raise ValueError('{}: tensor checksum mismatch'.format(name))
        out[name] = np.frombuffer(bytes(raw), dtype=_NP_OF[e['dtype']]).reshape(e['shape'])
    return out


# ----------------------------------------------------------------------------- writer (tests + export of our checkpoints)
def _entry_proto(dtype, shape, offset, size, crc):
    dims = b''.join(b'\x12' + _put_varint(len(d)) + d for d in (b'\x08' + _put_varint(int(s)) for s in shape))
    out = b'\x08' + _put_varint(dtype) + b'\x12' + _put_varint(len(dims)) + dims
    if offset:
        out += b'\x20' + _put_varint(offset)
    out += b'\x28' + _put_varint(size) + b'\x35' + struct.pack('<I', crc)
    return out


def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (key, val) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last = key
    for r in restarts or [0]:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts) or 1)
    return bytes(out)


def save_tf_checkpoint(prefix, tensors, block_entries=64):
    """Write `tensors` (name -> numpy array) as a one-shard tensor bundle readable by TensorFlow and by the reader above."""
    names = sorted(tensors, key=lambda s: s.encode('utf-8'))
    data, items = bytearray(), [(b'', b'\x08\x01\x1a\x02\x08\x01')]        # BundleHeaderProto{num_shards: 1, version{producer: 1}}
    for n in names:
        a = np.asarray(tensors[n], order='C')
        raw = a.tobytes()
        items.append((n.encode('utf-8'), _entry_proto(_DT_OF[a.dtype], a.shape, len(data), len(raw), masked_crc32c(raw))))
        data += raw
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))
    out, index_items = bytearray(), []

    def emit(block):
        off = len(out)
        out.extend(block + b'\x00')
        out.extend(struct.pack('<I', masked_crc32c(block + b'\x00')))
        return _put_varint(off) + _put_varint(len(block))

    for s in range(0, len(items), block_entries):
        chunk = items[s:s + block_entries]
        index_items.append((chunk[-1][0] + b'\x00', emit(_build_block(chunk))))     # separator key >= last key of the block
    meta = emit(_build_block([]))
    index = emit(_build_block(index_items, restart_interval=1))
    footer = meta + index
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))


def find_checkpoint(directory, stem='bert_model.ckpt'):
    """`<directory>/<stem>` if its .index exists, else None."""
    prefix = os.path.join(directory or '', stem)
    return prefix if directory and os.path.exists(prefix + '.index') else None
