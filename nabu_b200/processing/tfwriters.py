"""Writers of nabu's prepared-data format without TensorFlow -- the inverse of
tfreaders.py, so data directories the Trainer / Evaluator / Recognizer consume can be produced here as well.

reference: processing/tfwriters/tfwriter.py:34-55 (`TfWriter.write`: one TFRecord file `data/file<n>` holding ONE
serialized tf.train.Example per utterance, plus a `<name>\\t<file>` line in pointers.scp), array_writer.py:11-27
(`shape` = int32 shape bytes, `data` = float32 bytes), string_writer.py:10-27 (`length` = len of the string, `data` =
the string).  The metadata files next to them are what the processors write at the end of `run data`
(processors/audio_processor.py:76-88: sequence_length_histogram.npy, max_length, dim; text_processor.py:71-87:
max_length, sequence_length_histogram.npy, alphabet, dim, nonesymbol)."""
import os
from abc import ABCMeta, abstractmethod

import numpy as np

from . import tfrecord


class TfWriter(object, metaclass=ABCMeta):
    """TfWriter(datadir).write(data, name)"""

    def __init__(self, datadir):
        if not os.path.exists(datadir):
            os.makedirs(datadir)
        self.datadir = datadir
        self.scp_file = os.path.join(datadir, 'pointers.scp')
        self.write_dir = os.path.join(datadir, 'data')
        os.makedirs(self.write_dir)
        self.filenum = 0
        self.lengths = []

    def write(self, data, name):
        filename = os.path.join(self.write_dir, 'file%d' % self.filenum)
        self.filenum += 1
        tfrecord.write_records(filename, [self._get_example(data)])
        with open(self.scp_file, 'a') as fid:
            fid.write('%s\t%s\n' % (name, filename))
        self.lengths.append(self._length(data))

    def _histogram(self):
        """max_length and the histogram of sequence lengths the bucketing plan is computed from"""
        max_length = max(self.lengths) if self.lengths else 0
        with open(os.path.join(self.datadir, 'max_length'), 'w') as fid:
            fid.write(str(max_length))
        np.save(os.path.join(self.datadir, 'sequence_length_histogram.npy'),
                np.bincount(np.asarray(self.lengths, np.int64), minlength=max_length + 1))

    @abstractmethod
    def _get_example(self, data):
        """the serialized tf.train.Example of one utterance"""

    @abstractmethod
    def _length(self, data):
        """the sequence length the processors count for this utterance"""


class ArrayWriter(TfWriter):
    """float32 feature matrices [T, dim] (audio_feature)"""

    def _get_example(self, data):
        data = np.asarray(data)
        return tfrecord.make_example({'shape': np.array(data.shape, np.int32).tobytes(),
                                      'data': data.reshape([-1]).astype(np.float32).tobytes()})

    def _length(self, data):
        return int(np.asarray(data).shape[0])

    def write_metadata(self, dim):
        self._histogram()
        with open(os.path.join(self.datadir, 'dim'), 'w') as fid:
            fid.write(str(dim))


class StringWriter(TfWriter):
    """space separated symbol strings (string / string_eos)"""

    def _get_example(self, data):
        return tfrecord.make_example({'length': [len(data)], 'data': data.encode('utf-8')})

    def _length(self, data):
        return len(data.split(' '))

    def write_metadata(self, alphabet, nonesymbol='<none>'):
        self._histogram()
        with open(os.path.join(self.datadir, 'alphabet'), 'w') as fid:
            fid.write(' '.join(alphabet))
        with open(os.path.join(self.datadir, 'dim'), 'w') as fid:
            fid.write(str(len(alphabet)))
        with open(os.path.join(self.datadir, 'nonesymbol'), 'w') as fid:
            fid.write(nonesymbol)


def factory(datatype):
    """processing/tfwriters/tfwriter_factory.py:4-30: the writer class for a data type of the database cfg"""
    if datatype == 'audio_feature':
        return ArrayWriter
    if datatype in ('string', 'string_eos'):
        return StringWriter
    if datatype in ('binary', 'alignment'):
        raise Exception('data type %s belongs to the hybrid (DNN) recipes, outside the hot path' % datatype)
    raise Exception('unknown data type: %s' % datatype)
