"""Pure-Python reader for TensorFlow `Saver` checkpoints (scope-table row f2; the files reference common/model.py:180-234
writes with `tf.train.Saver(...).save(sess, path)`), so that trained MultINN weights load here WITHOUT TensorFlow:

    variables = read_checkpoint('models/MultINN/generators')      # directory (uses its `checkpoint` file) or prefix
    load_tf_variables(model, variables)                           # utils/tf_import.py

Format (TensorFlow "tensor bundle", V2 checkpoints, the default since TF 1.0 - restated from its published layout;
NOT VERIFIED AGAINST A REAL FILE here: TensorFlow cannot be installed in this image and the reference ships no
checkpoint; tests/test_tf_import.py round-trips files produced by a writer that follows the same description):
  <prefix>.index                 a leveldb-style sorted string table, written uncompressed:
      data blocks | metaindex block | index block | 48-byte footer (two block handles, padding, magic 0xdb4775248b80fb57)
      block = entries (varint shared-key-bytes, varint unshared, varint value-length, key suffix, value)
              + uint32 restart offsets + uint32 restart count, followed by a 5-byte trailer (compression type, masked crc32c)
      key ""            -> BundleHeaderProto  {1: num_shards, 2: endianness, 3: version}
      key <tensor name> -> BundleEntryProto   {1: dtype, 2: shape{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size,
                                               6: masked crc32c of the bytes, 7: slices (partitioned variables)}
  <prefix>.data-SSSSS-of-NNNNN   raw little-endian tensor bytes at [offset, offset + size) of shard SSSSS.
"""
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}


class CheckpointFormatError(ValueError):
    pass


# ----------------------------------------------------------------------------- primitives
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointFormatError('truncated varint')
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise CheckpointFormatError('varint longer than 64 bits')


_CRC_TABLE = None


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), bytewise table version (checkpoint verification only: slow for large tensors)."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    c = crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    """leveldb / TensorFlow store crcs rotated and offset so that a crc of data holding crcs stays well distributed."""
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _proto_fields(buf):
    """Minimal protobuf wire decoder: yields (field number, wire type, value) with bytes for length-delimited fields."""
    pos = 0
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
            raise CheckpointFormatError(f'unsupported protobuf wire type {wt}')
        yield field, wt, val


# ----------------------------------------------------------------------------- sorted string table
def _read_block(data, offset, size, verify):
    if offset + size + 5 > len(data):
        raise CheckpointFormatError('block handle points outside the index file')
    body, kind = data[offset:offset + size], data[offset + size]
    if verify:
        want = struct.unpack_from('<I', data, offset + size + 1)[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != want:
            raise CheckpointFormatError('index block checksum mismatch')
    if kind != 0:
        raise CheckpointFormatError('compressed index block (TensorFlow writes checkpoint indices uncompressed)')
    return body


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointFormatError('block too small')
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise CheckpointFormatError('bad restart array')
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > end:
            raise CheckpointFormatError('corrupt block entry')
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a leveldb-format table file, in key order."""
    data = open(path, 'rb').read()
    if len(data) < 48:
        raise CheckpointFormatError(f'{path}: too small for a table footer')
    footer = data[-48:]
    if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
        raise CheckpointFormatError(f'{path}: not a TensorFlow checkpoint index (bad table magic)')
    pos = 0
    _, pos = _varint(footer, pos)            # metaindex handle (unused)
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _varint(handle, 0)
        size, _ = _varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ----------------------------------------------------------------------------- tensor bundle
def _parse_entry(value):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'slices': 0}
    for field, _, val in _proto_fields(value):
        if field == 1:
            e['dtype'] = val
        elif field == 2:
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
            e['slices'] += 1
    return e


def resolve_prefix(path):
    """A checkpoint prefix from a prefix, an `.index` path, or a directory holding the Saver's `checkpoint` state file
    (`model_checkpoint_path: "<name>"`, as common/model.py:226-231 reads it through tf.train.get_checkpoint_state)."""
    if os.path.isdir(path):
        state = os.path.join(path, 'checkpoint')
        if os.path.isfile(state):
            m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', open(state).read(), re.M)
            if m:
                p = m.group(1)
                return p if os.path.isabs(p) else os.path.join(path, p)
        cands = sorted(f[:-6] for f in os.listdir(path) if f.endswith('.index'))
        if len(cands) == 1:
            return os.path.join(path, cands[0])
        raise FileNotFoundError(f'{path}: no `checkpoint` state file and {len(cands)} *.index files')
    return path[:-6] if path.endswith('.index') else path


def read_checkpoint(path, names=None, verify_index=True, verify_data=False):
    """{variable name: ndarray} of a TF V2 checkpoint. `names`: optional filter (iterable of names or a predicate).
    verify_data checks every tensor's crc32c (pure Python: about a second per 10 MB)."""
    prefix = resolve_prefix(path)
    if not os.path.isfile(prefix + '.index'):
        raise FileNotFoundError(f'{prefix}.index not found (V1 `.ckpt` single-file checkpoints are not supported)')
    entries = read_table(prefix + '.index', verify=verify_index)
    if not entries or entries[0][0] != b'':
        raise CheckpointFormatError('missing bundle header entry')
    num_shards, endianness = 1, 0
    for field, _, val in _proto_fields(entries[0][1]):
        if field == 1:
            num_shards = val
        elif field == 2:
            endianness = val
    if endianness != 0:
        raise CheckpointFormatError('big-endian checkpoints are not supported')
    want = names if callable(names) or names is None else set(names).__contains__
    shards, out = {}, {}
    for key, value in entries[1:]:
        name = key.decode()
        if want is not None and not want(name):
            continue
        e = _parse_entry(value)
        if e['slices']:
            raise CheckpointFormatError(f'{name}: partitioned (sliced) variables are not supported')
        if e['dtype'] not in DTYPES:
            raise CheckpointFormatError(f'{name}: unsupported dtype enum {e["dtype"]}')
        dt = np.dtype(DTYPES[e['dtype']])
        count = int(np.prod(e['shape'])) if e['shape'] else 1
        if count * dt.itemsize != e['size']:
            raise CheckpointFormatError(f'{name}: {e["size"]} bytes do not hold shape {e["shape"]} of {dt}')
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap(f'{prefix}.data-{sid:05d}-of-{num_shards:05d}', dtype=np.uint8, mode='r')
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if raw.size != e['size']:
            raise CheckpointFormatError(f'{name}: data shard {sid} is shorter than offset + size')
        if verify_data and e['crc32c'] is not None and mask_crc(crc32c(raw.tobytes())) != e['crc32c']:
            raise CheckpointFormatError(f'{name}: tensor checksum mismatch')
        out[name] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e['shape']).copy()
    return out


# ----------------------------------------------------------------------------- writer (export back to the reference)
def _put_varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_varint(field, n):
    return _put_varint(field << 3) + _put_varint(n)


def _pb_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b'', restart_interval

    def add(self, key, value):
        shared = 0
        if self.count and self.count % self.interval == 0:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def _emit_block(out, body):
    off = len(out)
    out += body + b'\x00' + struct.pack('<I', mask_crc(crc32c(body + b'\x00')))
    return _put_varint(off) + _put_varint(len(body))          # block handle


def write_table(path, items, block_size=4096):
    """Sorted (key, value) pairs -> leveldb-format table file (uncompressed, like TensorFlow's checkpoint index)."""
    out, index, blk = bytearray(), _BlockBuilder(1), _BlockBuilder()
    items = sorted(items)
    for i, (key, value) in enumerate(items):
        blk.add(key, value)
        if len(blk.buf) >= block_size or i == len(items) - 1:
            index.add(key, _emit_block(out, blk.finish()))      # index key = last key of the block
            blk = _BlockBuilder()
    meta = _emit_block(out, _BlockBuilder().finish())
    idx = _emit_block(out, index.finish())
    footer = meta + idx
    out += footer + bytes(40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    with open(path, 'wb') as f:
        f.write(out)


def write_checkpoint(prefix, variables, write_state_file=True, block_size=4096):
    """{name: ndarray} -> `<prefix>.index` + `<prefix>.data-00000-of-00001` (+ the `checkpoint` state file of the
    directory), the layout `tf.train.Saver().restore` / common/model.py:216-234 read. With
    utils/tf_import.export_tf_variables this hands weights trained here back to the TF reference."""
    inv = {np.dtype(v): k for k, v in DTYPES.items()}
    items = [(b'', _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))]           # header: 1 shard, little endian, producer 1
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(f'{prefix}.data-00000-of-00001', 'wb') as f:
        offset = 0
        for name in sorted(variables):
            a = np.asarray(variables[name])                       # (ascontiguousarray would turn scalars into [1])
            if a.dtype not in inv:
                raise CheckpointFormatError(f'{name}: dtype {a.dtype} has no TensorFlow enum here')
            raw = a.tobytes()
            shape = b''.join(_pb_bytes(2, _pb_varint(1, int(d))) for d in a.shape)
            entry = _pb_varint(1, inv[a.dtype]) + _pb_bytes(2, shape)
            if offset:
                entry += _pb_varint(4, offset)
            entry += _pb_varint(5, len(raw)) + _put_varint((6 << 3) | 5) + struct.pack('<I', mask_crc(crc32c(raw)))
            items.append((name.encode(), entry))
            f.write(raw)
            offset += len(raw)
    write_table(prefix + '.index', items, block_size)
    if write_state_file:
        base = os.path.basename(prefix)
        with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), 'checkpoint'), 'w') as f:
            f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')
