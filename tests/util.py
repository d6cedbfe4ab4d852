"""Shared helpers for the parity tests: seeded synthetic inputs and the
configs the reference recipes would provide."""
import configparser

import numpy as np


def make_conf(text):
    c = configparser.ConfigParser()
    c.read_string(text)
    return c


def synthetic_ctc_batch(B, T, D, V, ragged, seed=1234):
    """features N(0,1) [B,T,D]; lengths = T or U{ceil(.6T)..T}; labels U{0..V-2}, L = len//10."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    if ragged:
        lens = np.random.default_rng(4321).integers(int(np.ceil(0.6 * T)), T + 1, size=B).astype(np.int32)
        lens[0] = T
    else:
        lens = np.full(B, T, np.int32)
    lab_len = np.maximum(lens // 10, 1).astype(np.int32)
    Lmax = int(lab_len.max())
    labels = np.random.default_rng(99).integers(0, V - 1, size=(B, Lmax)).astype(np.int32)
    for b in range(B):
        x[b, lens[b]:] = 0
    return x, lens, labels, lab_len


def synthetic_las_batch(B, T, D, V, U, ragged, seed=1234):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    if ragged:
        lens = np.random.default_rng(4321).integers(int(np.ceil(0.6 * T)), T + 1, size=B).astype(np.int32)
        lens[0] = T
        tl = np.random.default_rng(77).integers(max(1, U // 2), U + 1, size=B).astype(np.int32)
        tl[0] = U
    else:
        lens = np.full(B, T, np.int32)
        tl = np.full(B, U, np.int32)
    targets = np.random.default_rng(99).integers(0, V - 1, size=(B, U)).astype(np.int32)
    for b in range(B):
        targets[b, tl[b] - 1] = V - 1          # EOS terminated (string_reader_eos.py:57-60)
        targets[b, tl[b]:] = 0
        x[b, lens[b]:] = 0
    return x, lens, targets, tl


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
