"""Row f1: nabu's on-disk data format and bucketed input pipeline, without TensorFlow.

Pins: CRC-32C check value and TFRecord framing; tf.train.Example parsing on hand-assembled protobuf bytes (packed
and unpacked repeated scalars); `bucket_boundaries` against vectors produced by the reference's own function
(tests/golden/make_bucket_golden.py); reader semantics (float32 [T, dim], EOS = alphabet size appended, length + 1)
and the batch plan of input_pipeline.py:131-160 restated independently in the test.
"""
import json
import os

import numpy as np
import pytest

from nabu_b200.processing import input_pipeline as ip
from nabu_b200.processing import tfreaders, tfrecord, tfwriters

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_crc32c_and_framing(tmp_path):
    assert tfrecord.crc32c(b'123456789') == 0xE3069283            # the CRC-32C check value
    assert tfrecord.crc32c(b'') == 0
    path = str(tmp_path / 'a.rec')
    tfrecord.write_records(path, [b'hello', b'', b'x' * 1000])
    assert list(tfrecord.read_records(path)) == [b'hello', b'', b'x' * 1000]
    raw = bytearray(open(path, 'rb').read())
    assert raw[:8] == (5).to_bytes(8, 'little')                   # uint64 length first
    raw[13] ^= 1                                                  # flip a payload bit
    open(path, 'wb').write(raw)
    with pytest.raises(IOError):
        list(tfrecord.read_records(path))


def test_example_parser_on_hand_assembled_protobuf():
    # Example{features{feature{"length": Int64List[3]}}}, packed (what protobuf emits) and unpacked
    packed = bytes.fromhex('0a110a0f0a066c656e67746812051a030a0103')
    unpacked = bytes.fromhex('0a100a0e0a066c656e67746812041a020803')
    for blob in (packed, unpacked):
        p = tfrecord.parse_example(blob)
        assert list(p) == ['length'] and p['length'].tolist() == [3]
    # bytes + float features, negative int64 (10-byte varint)
    ex = tfrecord.make_example({'data': b'a b', 'w': np.array([1.5, -2.0], np.float32), 'n': [-1]})
    p = tfrecord.parse_example(ex)
    assert p['data'] == [b'a b'] and p['w'].tolist() == [1.5, -2.0] and p['n'].tolist() == [-1]
    assert tfrecord.make_example({'length': [3]}) == packed


def test_bucket_boundaries_match_the_reference_function():
    cases = json.load(open(os.path.join(GOLD, 'bucket_boundaries.json')))
    assert len(cases) >= 8
    for c in cases:
        assert ip.bucket_boundaries(c['histogram'], c['numbuckets']) == c['boundaries']


def _write_stream(root, kind, items, dim=None, alphabet=None):
    """a nabu data directory (pointers.scp, data/file<i>, max_length, sequence_length_histogram.npy, dim | alphabet)
    written by the package's own writers"""
    writer = tfwriters.factory('audio_feature' if kind == 'audio' else 'string_eos')(root)
    for name, value in items:
        writer.write(value, name)
    if kind == 'audio':
        writer.write_metadata(dim)
    else:
        writer.write_metadata(alphabet)


def test_writers_lay_out_a_nabu_data_directory(tmp_path):
    """byte-level layout against the format description, independent of the readers"""
    root = str(tmp_path / 'fbank')
    w = tfwriters.factory('audio_feature')(root)
    x = np.arange(6, dtype=np.float32).reshape(3, 2)
    w.write(x, 'utt a')
    w.write(np.zeros((5, 2), np.float32), 'uttb')
    w.write_metadata(2)
    assert open(os.path.join(root, 'pointers.scp')).read() == 'utt a\t%s/data/file0\nuttb\t%s/data/file1\n' % (root, root)
    assert open(os.path.join(root, 'max_length')).read() == '5' and open(os.path.join(root, 'dim')).read() == '2'
    assert np.load(os.path.join(root, 'sequence_length_histogram.npy')).tolist() == [0, 0, 0, 1, 0, 1]
    (rec,) = list(tfrecord.read_records(os.path.join(root, 'data', 'file0')))
    ex = tfrecord.parse_example(rec)
    assert ex['shape'] == [np.array([3, 2], np.int32).tobytes()] and ex['data'] == [x.tobytes()]
    root = str(tmp_path / 'text')
    w = tfwriters.factory('string_eos')(root)
    w.write('b a d', 'utt a')
    w.write_metadata(['a', 'b', 'c', 'd'])
    ex = tfrecord.parse_example(next(iter(tfrecord.read_records(os.path.join(root, 'data', 'file0')))))
    assert ex['data'] == [b'b a d'] and list(ex['length']) == [5]             # string_writer.py: len of the STRING
    assert open(os.path.join(root, 'alphabet')).read() == 'a b c d' and open(os.path.join(root, 'max_length')).read() == '3'
    with pytest.raises(Exception, match='hybrid'):
        tfwriters.factory('alignment')


def test_readers_and_bucketed_batches(tmp_path):
    rng = np.random.default_rng(0)
    alphabet = ['a', 'b', 'c', 'd']
    lens = [12, 30, 7, 25, 18, 9, 28, 14, 22, 11]
    feats = [('utt%d' % i, rng.standard_normal((L, 5)).astype(np.float32)) for i, L in enumerate(lens)]
    texts = [('utt%d' % i, ' '.join(rng.choice(alphabet, size=1 + L // 6))) for i, L in enumerate(lens)]
    fdir, tdir = str(tmp_path / 'fbank'), str(tmp_path / 'text')
    _write_stream(fdir, 'audio', feats, dim=5)
    _write_stream(tdir, 'text', texts[:-1], alphabet=alphabet)            # the last utterance has no transcription
    fconf, tconf = {'dir': fdir, 'type': 'audio_feature'}, {'dir': tdir, 'type': 'string_eos'}

    elements, names = ip.get_filenames([[fconf], [tconf]])
    assert names == ['utt%d-0' % i for i in range(9)] and all(len(e.split('\t')) == 2 for e in elements)

    ar = tfreaders.factory('audio_feature')([fdir])
    x, n = ar(elements[1].split('\t')[0])
    assert n == 30 and x.dtype == np.float32 and np.array_equal(x, feats[1][1])
    sr = tfreaders.factory('string_eos')([tdir])
    ids, n = sr(elements[1].split('\t')[1])
    want = [alphabet.index(s) for s in texts[1][1].split(' ')] + [len(alphabet)]       # EOS = alphabet size
    assert ids.dtype == np.int32 and ids.tolist() == want and n == len(want)
    assert sr.metadata['eos_label'] == 4 and sr.metadata['max_length'] == max(1 + L // 6 for L in lens[:-1]) + 1

    # the plain string reader of the CTC recipes (type = string): same ids, no EOS, lengths and histogram as written
    pr = tfreaders.factory('string')([tdir])
    ids2, n2 = pr(elements[1].split('\t')[1])
    assert ids2.tolist() == want[:-1] and n2 == len(want) - 1 and ids2.dtype == np.int32
    assert pr.metadata['max_length'] == sr.metadata['max_length'] - 1 and 'eos_label' not in pr.metadata
    assert pr.metadata['sequence_length_histogram'].sum() == 9

    # batch plan restated: boundaries from the histogram of the FIRST stream, variable batch sizes, num_steps
    hist = ar.metadata['sequence_length_histogram']
    bounds = ip.bucket_boundaries(hist, 3)
    sizes = [max(int(4 * bounds[0] / b), 1) for b in bounds + [hist.size]]
    src = ip.BatchSource([[fconf], [tconf]], ['features'], ['text'], batch_size=4, numbuckets=3,
                         variable_batch_size=True, allow_smaller_final_batch=True)
    assert src.boundaries == bounds and src.batch_sizes == sizes and src.input_dims == {'features': 5}
    seen = 0
    for inputs, ilen, targets, tlen in src:
        B = inputs['features'].shape[0]
        L = ilen['features'].numpy()
        b = np.searchsorted(bounds, L, side='right')
        assert len(set(b.tolist())) == 1 and B <= sizes[int(b[0])]                 # one bucket per batch
        assert inputs['features'].shape == (B, int(L.max()), 5)                    # dynamic_pad to the longest
        for i in range(B):
            assert np.all(inputs['features'][i, L[i]:].numpy() == 0)
            assert int(targets['text'][i, tlen['text'][i] - 1]) == 4               # ends with EOS
        seen += B
    assert seen == 9
    # fixed batch size, no buckets: num_steps = floor(#utterances / batch_size), the tail is dropped
    src = ip.BatchSource([[fconf], [tconf]], ['features'], ['text'], batch_size=4)
    assert len(src) == int(hist.sum() / 4) and sum(b[0]['features'].shape[0] for b in src) == 8


def test_batch_source_shards_every_batch_over_the_ranks(tmp_path):
    """Synchronous data parallelism from data directories: the ranks walk the same global batches and keep
    utterances rank::world of each; together they hold every global batch exactly once."""
    rng = np.random.default_rng(1)
    lens = [12, 30, 7, 25, 18, 9, 28, 14]
    feats = [('utt%d' % i, rng.standard_normal((L, 3)).astype(np.float32)) for i, L in enumerate(lens)]
    fdir = str(tmp_path / 'fbank')
    _write_stream(fdir, 'audio', feats, dim=3)
    fconf = {'dir': fdir, 'type': 'audio_feature'}
    whole = list(ip.BatchSource([[fconf]], ['features'], [], batch_size=4, shuffle_seed=5))
    parts = [list(ip.BatchSource([[fconf]], ['features'], [], batch_size=4, shuffle_seed=5, rank=r, world=2))
             for r in range(2)]
    assert len(whole) == len(parts[0]) == len(parts[1]) == 2
    for g, p0, p1 in zip(whole, parts[0], parts[1]):
        L = g[1]['features']
        assert p0[1]['features'].tolist() == L[0::2].tolist() and p1[1]['features'].tolist() == L[1::2].tolist()
        for r, p in enumerate((p0, p1)):
            T = p[0]['features'].shape[1]
            assert np.array_equal(p[0]['features'].numpy(), g[0]['features'][r::2, :T].numpy())
    # batch sizes that do not divide the ranks (the reference's variable batch sizes: 16, 14, 13, 11, ...) are rounded DOWN
    # to a multiple of the world size and the step count follows: stock recipes run under torchrun unchanged
    odd = [ip.BatchSource([[fconf]], ['features'], [], batch_size=3, shuffle_seed=5, rank=r, world=2) for r in range(2)]
    assert odd[0].batch_sizes == [2] and len(odd[0]) == 4
    got = [list(o) for o in odd]
    assert len(got[0]) == len(got[1]) == 4 and all(b[0]['features'].shape[0] == 1 for b in got[0] + got[1])
    # variable batch sizes over buckets, 3 ranks, smaller final batches: every rank sees the same number of batches with
    # the same number of utterances in each
    kw = dict(batch_size=5, numbuckets=2, variable_batch_size=True, allow_smaller_final_batch=True, shuffle_seed=1)
    tri = [list(ip.BatchSource([[fconf]], ['features'], [], rank=r, world=3, **kw)) for r in range(3)]
    assert len(tri[0]) == len(tri[1]) == len(tri[2]) >= 1
    for b0, b1, b2 in zip(*tri):
        assert b0[0]['features'].shape[0] == b1[0]['features'].shape[0] == b2[0]['features'].shape[0] >= 1


def test_prefetch_thread_yields_the_same_batches_in_the_same_order(tmp_path):
    rng = np.random.default_rng(2)
    lens = rng.integers(5, 40, size=23)
    fdir = str(tmp_path / 'fbank')
    _write_stream(fdir, 'audio', [('u%d' % i, rng.standard_normal((L, 4)).astype(np.float32))
                                  for i, L in enumerate(lens)], dim=4)
    fconf = {'dir': fdir, 'type': 'audio_feature'}
    kw = dict(batch_size=4, numbuckets=3, variable_batch_size=True, allow_smaller_final_batch=True, shuffle_seed=1)
    plain = ip.BatchSource([[fconf]], ['features'], [], **kw)
    ahead = ip.BatchSource([[fconf]], ['features'], [], prefetch=2, **kw)
    for epoch in range(2):                                   # the shuffle advances per epoch on both
        a, b = list(plain), list(ahead)
        assert len(a) == len(b) > 3
        for x, y in zip(a, b):
            assert np.array_equal(x[0]['features'].numpy(), y[0]['features'].numpy())
            assert x[1]['features'].tolist() == y[1]['features'].tolist()
    # a consumer that stops early does not leave the producer blocked; a reader error surfaces in the consumer
    it = iter(ahead)
    next(it)
    it.close()
    os.remove(os.path.join(fdir, 'data', 'file3'))
    with pytest.raises(Exception):
        list(ip.BatchSource([[fconf]], ['features'], [], prefetch=2, **kw))
