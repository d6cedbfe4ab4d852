"""TensorFlow checkpoint (V2 "tensor bundle") reader / writer without TensorFlow -- SURVEY.md section 8 row f3.

nabu saves and restores its models with `tf.train.Saver(variables, sharded=True)` (reference:
components/hooks.py:6-52; `model/network.ckpt` in trainers/trainer.py:617 and recognizer.py:107,
`logdir/validated.ckpt` in trainers/trainer.py:622).  TF 1.8's Saver writes the V2 format: `<prefix>.index` plus
`<prefix>.data-SSSSS-of-NNNNN`.  The format lives in TensorFlow (external, recalled from the 1.8 tree:
core/util/tensor_bundle/tensor_bundle.{h,cc}, core/protobuf/tensor_bundle.proto, core/lib/io/{table,block,format}.cc
-- the LevelDB table format); nothing under /root/reference pins it, so the codec is checked by round trips, by
hand-assembled bytes and by the format's own checksums ("parity unpinned" until a real TF-1.8 file is read with it).

`.index`: an immutable sorted string table
    data blocks | metaindex block | index block | footer(48 bytes)
    block   = entries, uint32 restart offsets[], uint32 num_restarts ; followed by a 5-byte trailer
              (compression type, 0 = none / 1 = snappy; uint32 masked crc32c of block + type byte)
    entry   = varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value
    index block entry: key >= last key of a data block, value = BlockHandle(varint64 offset, varint64 size)
    footer  = metaindex BlockHandle, index BlockHandle, zero padding to 40 bytes, magic 0xdb4775248b80fb57 (LE)
  key ""  -> BundleHeaderProto {1: num_shards, 2: endianness (0 little), 3: VersionDef{1: producer}}
  key var -> BundleEntryProto  {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id, 4: offset, 5: size,
                                6: fixed32 masked crc32c of the tensor bytes, 7: slices (partitioned variables)}
`.data-*`: the raw little-endian tensor bytes at [offset, offset + size).
"""
import os
import struct

import numpy as np

from .tfrecord import _enc_varint, _fields, _ld, _varint, crc32c, masked_crc32c

TABLE_MAGIC = 0xdb4775248b80fb57
RESTART_INTERVAL = 16            # table::Options::block_restart_interval
BLOCK_SIZE = 262144              # table::Options::block_size

# tensorflow/core/framework/types.proto
DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'), 5: np.dtype('<i2'),
          6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'), 17: np.dtype('<u2'), 19: np.dtype('<f2'),
          22: np.dtype('<u4'), 23: np.dtype('<u8')}
DTYPE_ENUM = {v: k for k, v in DTYPES.items()}


def _mask(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- sorted string table ---------------------------------------------------------------------------------------

def _snappy_uncompress(buf):
    """Raw snappy block format (the index of a bundle is normally uncompressed; kept for files written with it)."""
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                       # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], 'little')
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise IOError('corrupted snappy block')
        for _ in range(ln):                                 # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise IOError('corrupted snappy block (length)')
    return bytes(out)


def _read_block(buf, offset, size, check_crc=True):
    block = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if len(block) != size or len(trailer) != 5:
        raise IOError('truncated table block')
    if check_crc and _mask(crc32c(block + trailer[:1])) != struct.unpack('<I', trailer[1:])[0]:
        raise IOError('table block checksum mismatch')
    if trailer[0] == 1:
        block = _snappy_uncompress(block)
    elif trailer[0] != 0:
        raise IOError('unknown block compression %d' % trailer[0])
    return block


def _block_entries(block):
    """(key, value) pairs of one block, undoing the shared-prefix key compression"""
    num_restarts = struct.unpack('<I', block[-4:])[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key, out = 0, b'', []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path, check_crc=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order"""
    with open(path, 'rb') as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack('<Q', buf[-8:])[0] != TABLE_MAGIC:
        raise IOError('%s is not a table file (bad magic number)' % path)
    footer = buf[-48:]
    pos = 0
    _, pos = _varint(footer, pos)               # metaindex handle (unused: bundles store no filter / properties)
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, check_crc)):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        out.extend(_block_entries(_read_block(buf, boff, bsize, check_crc)))
    return out


class _BlockBuilder(object):
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b''

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            n = min(len(key), len(self.last))
            while shared < n and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _enc_varint(shared) + _enc_varint(len(key) - shared) + _enc_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + \
            struct.pack('<I', len(self.restarts))


def write_table(path, items, block_size=BLOCK_SIZE):
    """items: (key bytes, value bytes) in strictly increasing key order"""
    out = bytearray()
    index = _BlockBuilder(1)

    def emit(block):
        handle = _enc_varint(len(out)) + _enc_varint(len(block))
        out.extend(block + b'\x00' + struct.pack('<I', _mask(crc32c(block + b'\x00'))))
        return handle

    builder, last = _BlockBuilder(RESTART_INTERVAL), None
    for key, value in items:
        if last is not None and key <= last:
            raise ValueError('table keys must be strictly increasing')
        builder.add(key, value)
        last = key
        if builder.size() >= block_size:
            index.add(last, emit(builder.finish()))         # TF shortens the separator key; any key >= last works
            builder = _BlockBuilder(RESTART_INTERVAL)
    if builder.count or last is None:
        index.add(last if last is not None else b'', emit(builder.finish()))
    meta = emit(_BlockBuilder(RESTART_INTERVAL).finish())
    idx = emit(index.finish())
    footer = meta + idx
    out.extend(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC))
    with open(path, 'wb') as f:
        f.write(bytes(out))


# ---- bundle ----------------------------------------------------------------------------------------------------

def _parse_entry(value):
    entry = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'slices': 0}
    for num, wire, val in _fields(value):
        if num == 1:
            entry['dtype'] = val
        elif num == 2:
            for n2, _, dim in _fields(val):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _fields(dim):
                        if n3 == 1:
                            size = v3
                    entry['shape'].append(size)
        elif num == 3:
            entry['shard_id'] = val
        elif num == 4:
            entry['offset'] = val
        elif num == 5:
            entry['size'] = val
        elif num == 6:
            entry['crc32c'] = val if isinstance(val, int) else struct.unpack('<I', val)[0]
        elif num == 7:
            entry['slices'] += 1
    return entry


def _shard_name(prefix, shard, num_shards):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


def list_variables(prefix):
    """[(name, shape, numpy dtype)] of a checkpoint, like tf.train.list_variables"""
    out = []
    for key, value in read_table(prefix + '.index'):
        if key == b'':
            continue
        e = _parse_entry(value)
        out.append((key.decode('utf-8'), tuple(e['shape']), DTYPES.get(e['dtype'])))
    return out


def read_checkpoint(prefix, names=None, check_crc=True):
    """{variable name: ndarray} of the checkpoint `<prefix>.index` + `<prefix>.data-*`.  `names`: read only these."""
    items = read_table(prefix + '.index', check_crc)
    if not items or items[0][0] != b'':
        raise IOError('%s.index has no bundle header' % prefix)
    num_shards, endianness = 1, 0
    for num, _, val in _fields(items[0][1]):
        if num == 1:
            num_shards = val
        elif num == 2:
            endianness = val
    if endianness != 0:
        raise IOError('big-endian bundles are not supported')
    shards, out = {}, {}
    try:
        for key, value in items[1:]:
            name = key.decode('utf-8')
            if names is not None and name not in names:
                continue
            e = _parse_entry(value)
            if e['slices']:
                raise IOError('%s: partitioned variables are not supported' % name)
            if e['dtype'] not in DTYPES:
                raise IOError('%s: unsupported dtype enum %d' % (name, e['dtype']))
            if e['shard_id'] not in shards:
                shards[e['shard_id']] = open(_shard_name(prefix, e['shard_id'], num_shards), 'rb')
            f = shards[e['shard_id']]
            f.seek(e['offset'])
            raw = f.read(e['size'])
            dt = DTYPES[e['dtype']]
            if len(raw) != e['size'] or e['size'] != int(np.prod(e['shape'], dtype=np.int64)) * dt.itemsize:
                raise IOError('%s: truncated or inconsistent tensor data' % name)
            if check_crc and e['crc32c'] is not None and _mask(crc32c(raw)) != e['crc32c']:
                raise IOError('%s: tensor checksum mismatch' % name)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e['shape']).copy()
    finally:
        for f in shards.values():
            f.close()
    if names is not None:
        missing = [n for n in names if n not in out]
        if missing:
            raise KeyError('not in checkpoint %s: %s' % (prefix, ', '.join(missing)))
    return out


def write_checkpoint(prefix, arrays, shard_of=None, num_shards=1, producer=1, state_file=True):
    """Write {name: ndarray} as a V2 bundle.  `shard_of(name) -> shard id` spreads the variables over `num_shards`
    data files the way a sharded Saver spreads them over parameter-server devices (default: one shard).  `producer`
    1 = kTensorBundleVersion, what TF's BundleWriter puts in the header's VersionDef (tensor_bundle.cc; min_consumer 0)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    files = [open(_shard_name(prefix, s, num_shards), 'wb') for s in range(num_shards)]
    offsets = [0] * num_shards
    header = b'\x08' + _enc_varint(num_shards) + _ld(3, b'\x08' + _enc_varint(producer))
    items = [(b'', header)]
    try:
        for name in sorted(arrays, key=lambda n: n.encode('utf-8')):
            arr = np.asarray(arrays[name], order='C')           # (ascontiguousarray would turn a scalar into [1])
            if arr.dtype.byteorder == '>':
                arr = arr.astype(arr.dtype.newbyteorder('<'))
            enum = DTYPE_ENUM.get(arr.dtype)                 # native little-endian dtypes compare equal to '<..'
            if enum is None:
                raise ValueError('%s: dtype %s cannot be stored' % (name, arr.dtype))
            raw = arr.tobytes()
            shard = shard_of(name) if shard_of else 0
            files[shard].write(raw)
            shape = b''.join(_ld(2, b'\x08' + _enc_varint(d)) for d in arr.shape)
            value = b'\x08' + _enc_varint(enum) + _ld(2, shape)
            if shard:
                value += b'\x18' + _enc_varint(shard)
            if offsets[shard]:
                value += b'\x20' + _enc_varint(offsets[shard])
            value += b'\x28' + _enc_varint(len(raw)) + b'\x35' + struct.pack('<I', _mask(crc32c(raw)))
            offsets[shard] += len(raw)
            items.append((name.encode('utf-8'), value))
    finally:
        for f in files:
            f.close()
    write_table(prefix + '.index', items)
    if state_file:
        write_state_file(prefix)


def write_state_file(prefix):
    """the `checkpoint` state file tf.train.latest_checkpoint reads, replaced atomically"""
    path = os.path.join(os.path.dirname(os.path.abspath(prefix)), 'checkpoint')
    base = os.path.basename(prefix)
    with open(path + '.tmp', 'w') as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    os.replace(path + '.tmp', path)


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix the directory's `checkpoint` state file names, or None"""
    path = os.path.join(directory, 'checkpoint')
    if not os.path.isfile(path):
        return None
    for line in open(path):
        if line.startswith('model_checkpoint_path:'):
            name = line.split(':', 1)[1].strip().strip('"')
            prefix = name if os.path.isabs(name) else os.path.join(directory, name)
            return prefix if os.path.isfile(prefix + '.index') else None
    return None


if __name__ == '__main__':                      # python -m nabu_b200.processing.tfcheckpoint <prefix> [name]
    import sys
    if len(sys.argv) < 2:
        raise SystemExit('usage: python -m nabu_b200.processing.tfcheckpoint <checkpoint prefix> [variable name]')
    if len(sys.argv) == 2:
        total = 0
        for var_name, var_shape, var_dtype in list_variables(sys.argv[1]):
            total += int(np.prod(var_shape, dtype=np.int64))
            print('%-100s %-18s %s' % (var_name, var_shape, var_dtype))
        print('%d values' % total)
    else:
        print(read_checkpoint(sys.argv[1], names={sys.argv[2]})[sys.argv[2]])
