"""Per-utterance readers of nabu's prepared data.

reference: processing/tfreaders/{tfreader.py, audio_feature_reader.py:12-78, string_reader_eos.py:13-111,
tfreader_factory.py}.  A reader is built from the data directories of one input / target stream, loads their
metadata files and turns one TFRecord file (one serialized Example) into `(array, sequence_length)`.
"""
import os

import numpy as np

from . import tfrecord


class TfReader(object):
    def __init__(self, datadirs):
        self.datadirs = list(datadirs)
        self.metadata = self._read_metadata(self.datadirs)

    def __call__(self, filename):
        records = list(tfrecord.read_records(filename))
        if len(records) != 1:
            raise IOError('%s: expected one example per file, found %d' % (filename, len(records)))
        return self._process_features(tfrecord.parse_example(records[0]))


class AudioFeatureReader(TfReader):
    """`data` = raw little-endian float32 bytes of a [T, dim] matrix (tfwriters/array_writer.py:11-27)."""

    def _read_metadata(self, datadirs):
        md = {}
        md['max_length'] = max(int(open(os.path.join(d, 'max_length')).read()) for d in datadirs)
        md['sequence_length_histogram'] = np.zeros([md['max_length'] + 1])
        for d in datadirs:
            h = np.load(os.path.join(d, 'sequence_length_histogram.npy'))
            md['sequence_length_histogram'][:h.shape[0]] += h
        md['dim'] = int(open(os.path.join(datadirs[0], 'dim')).read())
        for d in datadirs:
            if md['dim'] != int(open(os.path.join(d, 'dim')).read()):
                raise Exception('all audio feature reader dimensions must be the same')
        return md

    def _process_features(self, features):
        data = np.frombuffer(features['data'][0], '<f4').reshape(-1, self.metadata['dim'])
        return data, data.shape[0]


class StringReader(TfReader):
    """`data` = space-joined symbols -> int32 ids, no EOS (string_reader.py:11-103): what the CTC recipes read their
    targets with (config/recipes/DBLSTM/TIMIT/database.cfg: `type = string`).  Same alphabet lookup as the EOS reader
    below: `nonesymbol` at index 0, minus 1, so symbols map to 0..len(alphabet)-1."""

    def _read_metadata(self, datadirs):
        md = {}
        md['max_length'] = max(int(open(os.path.join(d, 'max_length')).read()) for d in datadirs)
        md['sequence_length_histogram'] = np.zeros([md['max_length'] + 1])
        for d in datadirs:
            h = np.load(os.path.join(d, 'sequence_length_histogram.npy'))
            md['sequence_length_histogram'][:h.shape[0]] += h
        nonesymbol = open(os.path.join(datadirs[0], 'nonesymbol')).read()
        alphabet = open(os.path.join(datadirs[0], 'alphabet')).read().split()
        for d in datadirs:
            if alphabet != open(os.path.join(d, 'alphabet')).read().split():
                raise Exception('all string reader alphabets must be the same')
        md['alphabet'] = [nonesymbol] + alphabet
        return md

    def _symbols(self, features):
        symbols = features['data'][0].decode('utf-8').split(' ')
        symbols = [s for s in symbols if s != ''] if symbols != [''] else []
        index = {s: i for i, s in enumerate(self.metadata['alphabet'])}
        try:
            return [index[s] - 1 for s in symbols]
        except KeyError:
            raise Exception('not all string elements found in alphabet: %r' % features['data'][0])

    def _process_features(self, features):
        ids = self._symbols(features)
        return np.array(ids, np.int32), len(ids)


class StringReaderEOS(TfReader):
    """`data` = space-joined symbols (tfwriters/string_writer.py:10-27) -> int32 ids, EOS (= alphabet size) appended;
    the returned length counts the EOS (string_reader_eos.py:88-111).  The reference's alphabet lookup puts the
    `nonesymbol` at index 0 and subtracts 1, so symbols map to 0..len(alphabet)-1 and the nonesymbol to -1."""

    def _read_metadata(self, datadirs):
        md = {}
        md['max_length'] = max(int(open(os.path.join(d, 'max_length')).read()) for d in datadirs) + 1
        md['sequence_length_histogram'] = np.zeros([md['max_length'] + 1])
        for d in datadirs:
            h = np.load(os.path.join(d, 'sequence_length_histogram.npy'))
            h = np.concatenate([[0], h])                 # every sequence grows by the EOS
            md['sequence_length_histogram'][:h.shape[0]] += h
        nonesymbol = open(os.path.join(datadirs[0], 'nonesymbol')).read()
        alphabet = open(os.path.join(datadirs[0], 'alphabet')).read().split()
        for d in datadirs:
            if alphabet != open(os.path.join(d, 'alphabet')).read().split():
                raise Exception('all string reader alphabets must be the same')
        md['alphabet'] = [nonesymbol] + alphabet
        md['eos_label'] = len(md['alphabet']) - 1
        return md

    def _process_features(self, features):
        symbols = features['data'][0].decode('utf-8').split(' ')
        symbols = [s for s in symbols if s != ''] if symbols != [''] else []
        index = {s: i for i, s in enumerate(self.metadata['alphabet'])}
        try:
            ids = [index[s] - 1 for s in symbols]
        except KeyError:
            raise Exception('not all string elements found in alphabet: %r' % features['data'][0])
        data = np.array(ids + [self.metadata['eos_label']], np.int32)
        return data, len(ids) + 1


def factory(datatype):
    """reference: processing/tfreaders/tfreader_factory.py"""
    if datatype == 'audio_feature':
        return AudioFeatureReader
    if datatype == 'string':
        return StringReader
    if datatype == 'string_eos':
        return StringReaderEOS
    raise Exception('unknown or unsupported data type: %s (the hot path reads audio_feature, string and string_eos)'
                    % datatype)
