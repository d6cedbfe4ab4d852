"""CPU oracle for the nabu hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nabu_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker
(or as the timed CPU arm), never as the product.

PARITY PIN (round 2): the reference (vrenkens/nabu @ 39deb62) holds no golden
vectors, known-answer tests or fixtures for this path
and its arithmetic lives in TensorFlow 1.8.0, which is neither vendored under
/root/reference nor installable here.  What pins the oracle:

1. THE REFERENCE'S OWN PYTHON, EXECUTED.  ``tests/golden/make_tf18shim_golden.py``
   imports the reference's unmodified modules from /root/reference (Model,
   Listener, DBLSTM, Speller / RNNDecoder, DNNDecoder, layer, ops, attention,
   rnn_cell, loss_functions, CTCDecoder, BeamSearchDecoder and the reference's
   own 600-line beam search) over ``tests/golden/tf18shim`` - an eager, torch-fp64
   restatement of the TensorFlow-1.8 calls they make - and commits inputs,
   variables (under the names the reference's scopes produce), logits, loss,
   every gradient and the decoders' outputs under
   ``tests/golden/tf18shim_cases/``.  ``tests/test_tf18_golden.py`` holds the
   oracle (CPU) and the CUDA path (GPU) to them at 1e-4 / bit-exact ids.  All
   code the reference authors itself is thereby pinned by running it.
2. TensorFlow's own kernels (LSTM cells, dynamic_rnn, AttentionWrapper,
   dynamic_decode, ctc_loss, top_k, ...) are NOT run: the shim restates them a
   second time, independently of this package, from the published r1.8
   sources; ``tf.nn.ctc_beam_search_decoder`` is not restated twice (the shim
   delegates it to this oracle).  For those ops the claim remains "matches the
   restated algorithm", backed by independent witnesses: brute-force CTC path
   enumeration, ``torch.nn.functional.ctc_loss`` (CPU), ``torch.nn.LSTM`` on
   packed sequences (CPU), torch-autograd twins of the attention decoder,
   exhaustive hypothesis enumeration for the beam searches, and the closed form
   of TF-Adam (``tests/golden/make_golden.py`` freezes those as fixtures).
3. A dump of a real TF-1.8 run (``tools/tf18_dump.py``, same layout) would be
   picked up by the same harness; none exists in this repository.
"""
from .nabu_oracle import *  # noqa: F401,F403
