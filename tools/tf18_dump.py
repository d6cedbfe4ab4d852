#!/usr/bin/env python
"""Golden vectors from the REAL reference: run inside a nabu checkout with Python 2 + tensorflow==1.8.0.

Nothing in this repository can execute the reference (Python 2, TF 1.8, neither is in the
build image), so every "matches TF" statement here means "matches the restated oracle".  This script is the way out:
a maintainer who has a TF-1.8 environment runs it once per recipe,

    cd /path/to/nabu && python /path/to/nabu_b200/tools/tf18_dump.py \
        --recipe config/recipes/DBLSTM/TIMIT --out /path/to/nabu_b200/tests/golden/tf18/dblstm_timit

and commits the directory it writes.  tests/test_tf18_golden.py then pins the oracle (CPU) and the CUDA path (GPU)
against these files: it restores `network.ckpt` with nabu_b200's own checkpoint reader (which is thereby pinned
against a checkpoint written by TensorFlow itself), feeds `inputs.npz` and compares with `outputs.npz`.

What is dumped (stochastic parts off: input_noise = 0, dropout = 1, sample_prob = 0, as in every parity run):
  model.cfg / trainer.cfg / recognizer.cfg   the cfgs used (copies, with the three overrides above)
  network.ckpt.*                             tf.train.Saver(model.variables, sharded=True), as SaveAtEnd writes it
  inputs.npz    features [B,T,D] f32, features_len [B] i32, targets [B,L] i32, targets_len [B] i32
  outputs.npz   logits [B,U,V], logits_len [B], loss (scalar, the trainer's loss function),
                grad/<variable name> for every model variable (d loss / d variable, unclipped),
                decoded_indices / decoded_values / decoded_shape (the recognizer's decoder, ctc_decoder) or
                decoded_sequences / decoded_lengths / decoded_scores / decoded_alignments (beam_search_decoder)

The script only calls the reference's public classes (Model, loss_functions.factory, decoder_factory.factory); written
for TF 1.8 / Python 2 but kept free of py2-only syntax.  It has never been run in this repository's build image.
"""
from __future__ import print_function

import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--recipe', required=True, help='recipe directory holding model.cfg, trainer.cfg, recognizer.cfg')
    ap.add_argument('--out', required=True)
    ap.add_argument('--batch', type=int, default=6)
    ap.add_argument('--frames', type=int, default=60)
    ap.add_argument('--dim', type=int, default=40)
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--time', action='store_true', help='also time 5 train steps on the CPU (frames/s)')
    args = ap.parse_args()

    sys.path.append(os.getcwd())
    import tensorflow as tf
    from six.moves import configparser
    from nabu.neuralnetworks.models.model import Model
    from nabu.neuralnetworks.trainers import loss_functions
    from nabu.neuralnetworks.decoders import decoder_factory

    def read(name):
        conf = configparser.ConfigParser()
        conf.read(os.path.join(args.recipe, name))
        return conf

    model_cfg, trainer_cfg, recognizer_cfg = read('model.cfg'), read('trainer.cfg'), read('recognizer.cfg')
    model_cfg.set('encoder', 'input_noise', '0')
    model_cfg.set('encoder', 'dropout', '1')
    if model_cfg.get('decoder', 'decoder') == 'speller':
        model_cfg.set('decoder', 'dropout', '1')
        model_cfg.set('decoder', 'sample_prob', '0')
    if not os.path.isdir(args.out):
        os.makedirs(args.out)
    for name, conf in (('model.cfg', model_cfg), ('trainer.cfg', trainer_cfg), ('recognizer.cfg', recognizer_cfg)):
        with open(os.path.join(args.out, name), 'w') as fid:
            conf.write(fid)

    in_name = model_cfg.get('io', 'inputs').split(' ')[0]
    out_name = model_cfg.get('io', 'outputs').split(' ')[0]
    out_dim = int(model_cfg.get('io', 'output_dims').split(' ')[0])
    trainlabels = int(trainer_cfg.get('trainer', 'trainlabels'))
    loss_name = trainer_cfg.get('trainer', 'loss')

    rng = np.random.RandomState(args.seed)
    B, T, D = args.batch, args.frames, args.dim
    x = rng.randn(B, T, D).astype(np.float32)
    x_len = rng.randint(int(0.6 * T), T + 1, size=B).astype(np.int32)
    x_len[0] = T
    for b in range(B):
        x[b, x_len[b]:] = 0
    if loss_name == 'CTC':
        # labels 0..out_dim-1 (the blank is the extra trainlabel), no EOS
        y_len = np.maximum(x_len // 10, 1).astype(np.int32)
        L = int(y_len.max())
        y = rng.randint(0, out_dim, size=(B, L)).astype(np.int32)
    else:
        # EOS-terminated targets as string_reader_eos produces them: EOS = out_dim (the extra trainlabel)
        y_len = rng.randint(3, 9, size=B).astype(np.int32)
        L = int(y_len.max())
        y = rng.randint(0, out_dim, size=(B, L)).astype(np.int32)
        for b in range(B):
            y[b, y_len[b] - 1] = out_dim
            y[b, y_len[b]:] = 0
    for b in range(B):
        y[b, y_len[b]:] = 0
    np.savez(os.path.join(args.out, 'inputs.npz'), features=x, features_len=x_len, targets=y, targets_len=y_len)

    tf.set_random_seed(args.seed)
    model = Model(conf=model_cfg, trainlabels=trainlabels, constraint=None)
    p_x = tf.placeholder(tf.float32, [B, T, D])
    p_xl = tf.placeholder(tf.int32, [B])
    p_y = tf.placeholder(tf.int32, [B, L])
    p_yl = tf.placeholder(tf.int32, [B])
    logits, logit_len = model({in_name: p_x}, {in_name: p_xl}, {out_name: p_y}, {out_name: p_yl}, True)
    loss = loss_functions.factory(loss_name)({out_name: p_y}, logits, logit_len, {out_name: p_yl})
    variables = model.variables
    grads = tf.gradients(loss, variables)
    decoder = decoder_factory.factory(recognizer_cfg.get('decoder', 'decoder'))(recognizer_cfg, model)
    decoded = decoder({in_name: p_x}, {in_name: p_xl})[out_name]
    saver = tf.train.Saver(variables, sharded=True)
    feed = {p_x: x, p_xl: x_len, p_y: y, p_yl: y_len}

    out = {}
    with tf.Session(config=tf.ConfigProto(device_count={'GPU': 0})) as sess:
        sess.run(tf.global_variables_initializer())
        saver.save(sess, os.path.join(args.out, 'network.ckpt'))
        lg, ll, ls, gr = sess.run([logits[out_name], logit_len[out_name], loss, grads], feed)
        out['logits'], out['logits_len'], out['loss'] = lg, ll, np.float32(ls)
        for var, g in zip(variables, gr):
            if isinstance(g, tf.IndexedSlicesValue):
                dense = np.zeros(var.shape.as_list(), np.float32)
                np.add.at(dense, g.indices, g.values)
                g = dense
            out['grad/' + var.op.name] = g
        dec = sess.run(decoded, feed)
        if isinstance(dec, tf.SparseTensorValue):
            out['decoded_indices'], out['decoded_values'], out['decoded_shape'] = dec.indices, dec.values, dec.dense_shape
        else:
            # beam_search_decoder returns (sequences [B,W,L], lengths [B,W], scores [B,W], alignments [B,W,L,T'])
            out['decoded_sequences'], out['decoded_lengths'], out['decoded_scores'] = dec[0], dec[1], dec[2]
            out['decoded_alignments'] = dec[3]
        if args.time:
            import time
            opt = tf.train.AdamOptimizer(1e-3)
            clipped = [tf.clip_by_value(g, -1., 1.) for g in grads]
            update = opt.apply_gradients(list(zip(clipped, variables)))
            sess.run(tf.variables_initializer(opt.variables()))
            sess.run([update, loss], feed)
            t0 = time.time()
            for _ in range(5):
                sess.run([update, loss], feed)
            out['cpu_frames_per_s'] = np.float64(5 * x_len.sum() / (time.time() - t0))
            print('TF-%s CPU: %.1f frames/s' % (tf.__version__, out['cpu_frames_per_s']))
    np.savez(os.path.join(args.out, 'outputs.npz'), **out)
    print('wrote %s (loss %.6f, %d variables)' % (args.out, ls, len(variables)))


if __name__ == '__main__':
    main()
