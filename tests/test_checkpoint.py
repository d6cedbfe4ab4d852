"""Row f3: TF checkpoint (V2 tensor bundle) interop -- CPU tests of the codec and of the
parameter store's name map.  No TF-1.8 file exists in the reference tree or in this image, so the format is pinned
by bytes assembled by hand from the format description (independent of the writer), by the published CRC-32C
check value, and by round trips."""
import os
import struct

import numpy as np
import pytest
import torch

from nabu_b200.processing import tfcheckpoint as C
from nabu_b200.processing import tfrecord as R
from tests.util import make_conf


def test_crc32c_native_matches_the_byte_loop():
    assert R.crc32c_py(b'123456789') == 0xE3069283            # the CRC-32C check value (RFC 3720 B.4)
    assert R._native(), 'libnabu_b200.so does not export nabu_crc32c'
    rng = np.random.default_rng(0)
    for n in [0, 1, 7, 8, 9, 31, 64, 1001, 4096 + 3]:
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert R.crc32c(data) == R.crc32c_py(data)
        k = n // 3                                            # continuation from a running value
        assert R.crc32c(data[k:], R.crc32c(data[:k])) == R.crc32c_py(data)
    assert R.crc32c(b'\x00' * 32) == 0x8A9136AA and R.crc32c(b'\xff' * 32) == 0x62A8AB43   # RFC 3720 B.4


def _block(entries):
    """one table block without key sharing, a single restart point"""
    body = b''
    for key, value in entries:
        body += bytes([0, len(key), len(value)]) + key + value
    body += struct.pack('<II', 0, 1)
    return body + b'\x00' + struct.pack('<I', C._mask(R.crc32c_py(body + b'\x00')))


def test_reader_on_a_hand_assembled_bundle(tmp_path):
    w = np.arange(6, dtype='<f4').reshape(2, 3)
    step = np.array(5, dtype='<i8')
    data = step.tobytes() + w.tobytes()                      # BundleWriter lays the tensors out in key order
    # BundleHeaderProto{num_shards=1, version{producer=1 = kTensorBundleVersion}}; BundleEntryProto per tensor
    header = bytes([0x08, 1, 0x1a, 2, 0x08, 1])
    e_w = bytes([0x08, 1, 0x12, 8, 0x12, 2, 0x08, 2, 0x12, 2, 0x08, 3, 0x20, 8, 0x28, 24, 0x35]) + \
        struct.pack('<I', C._mask(R.crc32c_py(w.tobytes())))
    e_s = bytes([0x08, 9, 0x12, 0, 0x28, 8, 0x35]) + struct.pack('<I', C._mask(R.crc32c_py(step.tobytes())))
    b0 = _block([(b'', header), (b'global_step', e_s), (b'w', e_w)])
    meta = _block([])
    index = _block([(b'w', bytes([0, len(b0) - 5]))])
    footer = bytes([len(b0), len(meta) - 5, len(b0) + len(meta), len(index) - 5])
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    prefix = str(tmp_path / 'hand.ckpt')
    with open(prefix + '.index', 'wb') as f:
        f.write(b0 + meta + index + footer)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(data)
    assert C.list_variables(prefix) == [('global_step', (), np.dtype('int64')), ('w', (2, 3), np.dtype('float32'))]
    out = C.read_checkpoint(prefix)
    assert out['global_step'].shape == () and int(out['global_step']) == 5
    np.testing.assert_array_equal(out['w'], w)
    # the writer produces the same index bytes for the same content
    C.write_checkpoint(str(tmp_path / 'ours.ckpt'), {'w': w, 'global_step': step})
    assert open(str(tmp_path / 'ours.ckpt.index'), 'rb').read() == open(prefix + '.index', 'rb').read()
    assert open(str(tmp_path / 'ours.ckpt.data-00000-of-00001'), 'rb').read() == data
    # a flipped data byte is caught by the tensor checksum, a flipped index byte by the block checksum
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(data[:3] + bytes([data[3] ^ 1]) + data[4:])
    with pytest.raises(IOError, match='checksum'):
        C.read_checkpoint(prefix)
    raw = bytearray(open(prefix + '.index', 'rb').read())
    raw[10] ^= 1
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(raw))
    with pytest.raises(IOError, match='checksum'):
        C.read_table(prefix + '.index')


def test_table_round_trip_with_prefix_compression_and_many_blocks(tmp_path):
    rng = np.random.default_rng(1)
    keys = sorted(set('Listener/features/layer%d/%s/%d' % (rng.integers(0, 9), rng.choice(['fw', 'bw']), i)
                      for i in range(3000)))
    items = [(k.encode(), rng.integers(0, 256, size=rng.integers(0, 60), dtype=np.uint8).tobytes()) for k in keys]
    path = str(tmp_path / 'table')
    C.write_table(path, items, block_size=700)
    assert C.read_table(path) == items
    with pytest.raises(ValueError):
        C.write_table(path, [(b'b', b''), (b'a', b'')])


def test_snappy_block_decoder():
    # literal "abcd", copy-1 (offset 4, length 8, overlapping), literal "xy", copy-2 (offset 2, length 3)
    comp = bytes([17, (4 - 1) << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4]) + bytes([(2 - 1) << 2]) + b'xy' + \
        bytes([((3 - 1) << 2) | 2, 2, 0])
    assert C._snappy_uncompress(comp) == b'abcdabcdabcdxyxyx'


def test_sharded_bundle_and_dtypes(tmp_path):
    rng = np.random.default_rng(2)
    arrays = {'a/kernel': rng.standard_normal((5, 7)).astype(np.float32), 'a/bias': np.zeros(7, np.float32),
              'global_step': np.array(1234, np.int64), 'ids': rng.integers(0, 9, (3, 2, 2)).astype(np.int32),
              'empty': np.zeros((0, 4), np.float32), 'd': rng.standard_normal(3)}
    prefix = str(tmp_path / 'model' / 'network.ckpt')
    C.write_checkpoint(prefix, arrays, shard_of=lambda n: len(n) % 3, num_shards=3)
    assert sorted(os.listdir(str(tmp_path / 'model'))) == ['checkpoint', 'network.ckpt.data-00000-of-00003',
                                                           'network.ckpt.data-00001-of-00003',
                                                           'network.ckpt.data-00002-of-00003', 'network.ckpt.index']
    out = C.read_checkpoint(prefix)
    assert set(out) == set(arrays)
    for k, v in arrays.items():
        assert out[k].dtype == v.dtype and out[k].shape == v.shape and np.array_equal(out[k], v)
    assert set(C.read_checkpoint(prefix, names={'ids'})) == {'ids'}
    with pytest.raises(KeyError):
        C.read_checkpoint(prefix, names={'nope'})


def test_store_restores_by_the_reference_variable_names(tmp_path):
    from nabu_b200.neuralnetworks.models.model import Model
    conf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = 30\n[encoder]\nencoder = listener\n'
                     'num_units = 16\nnum_layers = 2\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = speller\n'
                     'num_layers = 2\nnum_units = 16\nattention = location_aware\nnumfilt = 3\nfiltersize = 5\n')
    a = Model(conf, 1, seed=1).build({'features': 40}, 'cpu')
    a.store.m.uniform_()
    a.store.v.uniform_()
    prefix = str(tmp_path / 'logdir' / 'validated.ckpt')
    a.store.save_tf_checkpoint(prefix, with_adam=True, global_step=17)
    names = dict((n, s) for n, s, _ in C.list_variables(prefix))
    # the names a nabu-trained checkpoint carries
    key = 'Listener/features/layer0/BLSTM/bidirectional_rnn/fw/layer_norm_basic_lstm_cell/kernel'
    assert names[key] == (40 + 16, 64) and names[key + '/Adam_1'] == (40 + 16, 64)
    assert names['Speller/decoder/attention_wrapper/location_aware_attention/conv1d/kernel'] == (5, 1, 3)
    assert names['global_step'] == ()
    b = Model(conf, 1, seed=2).build({'features': 40}, 'cpu')
    assert not torch.equal(a.store.theta, b.store.theta)
    assert b.store.load_tf_checkpoint(prefix, with_adam=True) == 17
    for var in a.store.order:
        sl = slice(var.offset, var.offset + var.numel)
        for buf in ('theta', 'm', 'v'):
            assert torch.equal(getattr(a.store, buf)[sl], getattr(b.store, buf)[sl])
    # Saver.restore semantics: a missing variable or another shape is an error
    other = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = 30\n[encoder]\nencoder = listener\n'
                      'num_units = 16\nnum_layers = 3\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = speller\n'
                      'num_layers = 2\nnum_units = 16\nattention = location_aware\nnumfilt = 3\nfiltersize = 5\n')
    c = Model(other, 1).build({'features': 40}, 'cpu')
    with pytest.raises(KeyError, match='layer3'):
        c.store.load_tf_checkpoint(prefix)
    d = Model(conf, 1).build({'features': 39}, 'cpu')
    with pytest.raises(ValueError, match='shape'):
        d.store.load_tf_checkpoint(prefix)
