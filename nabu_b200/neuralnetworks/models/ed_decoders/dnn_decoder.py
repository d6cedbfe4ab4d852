"""Feed-forward decoder (reference: .../ed_decoders/dnn_decoder.py:10-62).

With num_layers = 0 (every CTC recipe, e.g. config/recipes/DBLSTM/TIMIT/model.cfg:21-25) this is the
CTC output projection <output>/outlayer = tf.contrib.layers.linear; hidden layers belong to the
Kaldi-hybrid DNN recipes, which are outside the hot path."""
from . import ed_decoder
from .... import engine


class DNNDecoder(ed_decoder.EDDecoder):

    def _check(self):
        if int(self.conf['num_layers']) != 0:
            raise Exception('dnn_decoder: only num_layers = 0 (output projection) is on the B200 hot path')

    def declare(self, encoded_dims):
        self._check()
        dim = list(encoded_dims.values())[0]
        for o in self.output_dims:
            base = '%s/%s/outlayer' % (self.scope, o)
            self.store.get(base + '/weights', (dim, self.output_dims[o]), 'glorot')
            self.store.get(base + '/biases', (self.output_dims[o],), 'zeros')

    def _decode(self, encoded, encoded_seq_length, targets, target_seq_length, is_training):
        self._check()
        x = list(encoded.values())[0]
        lens = list(encoded_seq_length.values())[0]
        outputs, output_seq_length = {}, {}
        for o in self.output_dims:
            base = '%s/%s/outlayer' % (self.scope, o)
            W = self.store.get(base + '/weights', (x.shape[-1], self.output_dims[o]))
            b = self.store.get(base + '/biases', (self.output_dims[o],))
            outputs[o] = engine.linear(x, W, b)
            output_seq_length[o] = lens
        return outputs, output_seq_length, ()

    def zero_state(self, encoded_dim, batch_size):
        return ()
