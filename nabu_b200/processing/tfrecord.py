"""TFRecord framing + tf.train.Example (de)serialisation without TensorFlow.

nabu stores ONE serialized `tf.train.Example` per file (reference: processing/tfwriters/tfwriter.py:34-55,
`tf.python_io.TFRecordWriter`).  TFRecord framing (tensorflow/core/lib/io/record_writer.cc, TF 1.8, external):
    uint64 length | uint32 masked_crc32c(length) | byte data[length] | uint32 masked_crc32c(data)
with masked_crc = ((crc >> 15) | (crc << 17)) + 0xa282ead8 (mod 2^32), CRC-32C (Castagnoli, reflected 0x82F63B78).
`tf.train.Example` (tensorflow/core/example/{example,feature}.proto): Example{1: Features{1: map<string, Feature>}},
Feature{oneof 1: BytesList{repeated bytes 1}, 2: FloatList{repeated float 1 [packed]}, 3: Int64List{repeated int64 1
[packed]}}.  The parser accepts packed and unpacked repeated scalars.
"""
import struct

import numpy as np

_CRC_TABLE = None


def _table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    return _CRC_TABLE


_NATIVE = None


def _native():
    """nabu_crc32c of libnabu_b200.so (include/nabu_b200.h; slicing-by-8, ~1 GB/s) -- the byte loop below is the
    restatement it is tested against and what runs when the library has not been built (host-side IO only: no
    arithmetic of the hot path lives here)."""
    global _NATIVE
    if _NATIVE is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'libnabu_b200.so')
        try:
            fn = ctypes.CDLL(path).nabu_crc32c
            fn.restype, fn.argtypes = ctypes.c_uint, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint]
            _NATIVE = fn
        except (OSError, AttributeError):
            _NATIVE = False
    return _NATIVE


def crc32c_py(data, crc=0):
    t, c = _table(), crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data, crc=0):
    fn = _native()
    if fn:
        data = bytes(data)
        return int(fn(data, len(data), crc))
    return crc32c_py(data, crc)


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def read_records(path, check_crc=True):
    """yields the payload of every record in a TFRecord file"""
    with open(path, 'rb') as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError('%s: truncated record header' % path)
            length, lcrc = struct.unpack('<QI', head)
            if check_crc and masked_crc32c(head[:8]) != lcrc:
                raise IOError('%s: corrupted record length' % path)
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise IOError('%s: truncated record' % path)
            if check_crc and masked_crc32c(data) != struct.unpack('<I', tail)[0]:
                raise IOError('%s: corrupted record data' % path)
            yield data


def write_records(path, records):
    with open(path, 'wb') as f:
        for data in records:
            head = struct.pack('<Q', len(data))
            f.write(head + struct.pack('<I', masked_crc32c(head)) + data + struct.pack('<I', masked_crc32c(data)))


# ---- protobuf wire format (just what Example needs) ---------------------------------------------
def _varint(buf, pos):
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    """yields (field number, wire type, value) of one message; value is int (varint / fixed) or bytes"""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield num, wt, val


def parse_example(data):
    """serialized tf.train.Example -> {name: list of bytes | float32 array | int64 array}"""
    out = {}
    for num, _, features in _fields(data):
        if num != 1:
            continue
        for fnum, _, entry in _fields(features):
            if fnum != 1:
                continue
            key, feature = None, b''
            for enum, _, v in _fields(entry):
                if enum == 1:
                    key = bytes(v).decode('utf-8')
                elif enum == 2:
                    feature = v
            value = None
            for kind, _, lst in _fields(feature):
                if kind == 1:                                   # BytesList
                    value = [bytes(v) for n_, _, v in _fields(lst) if n_ == 1]
                elif kind == 2:                                 # FloatList
                    vals = []
                    for n_, wt, v in _fields(lst):
                        if n_ == 1:
                            vals.append(np.frombuffer(bytes(v), '<f4'))
                    value = np.concatenate(vals) if vals else np.zeros(0, np.float32)
                elif kind == 3:                                 # Int64List
                    vals = []
                    for n_, wt, v in _fields(lst):
                        if n_ != 1:
                            continue
                        if wt == 0:
                            vals.append(v)
                        else:
                            p = 0
                            while p < len(v):
                                x, p = _varint(v, p)
                                vals.append(x)
                    value = np.array([x - (1 << 64) if x >= (1 << 63) else x for x in vals], np.int64)
            out[key] = value
    return out


def _enc_varint(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def make_example(features):
    """{name: bytes | list of bytes | float array | int array} -> serialized tf.train.Example (sorted keys, packed)"""
    entries = b''
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, list) and all(isinstance(x, (bytes, bytearray)) for x in v):
            feat = _ld(1, b''.join(_ld(1, bytes(x)) for x in v))
        else:
            a = np.asarray(v)
            if a.dtype.kind == 'f':
                feat = _ld(2, _ld(1, a.astype('<f4').tobytes()))
            else:
                feat = _ld(3, _ld(1, b''.join(_enc_varint(int(x)) for x in a.reshape(-1))))
        entries += _ld(1, _ld(1, key.encode('utf-8')) + _ld(2, feat))
    return _ld(1, entries)
