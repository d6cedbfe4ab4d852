"""CPU oracle for the nabu hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nabu_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker
(or as the timed CPU arm), never as the product.

PARITY UNPINNED: the reference (vrenkens/nabu @ 39deb62) holds no golden
vectors, known-answer tests or fixtures for this path
and its arithmetic lives in TensorFlow 1.8.0, which is neither vendored under
/root/reference nor installable here.  The oracle therefore restates the
published TF-1.8 algorithms at the reference's own call sites (cited per
function) and is pinned instead against independent witnesses: brute-force CTC
path enumeration, ``torch.nn.functional.ctc_loss`` (CPU), ``torch.nn.LSTM`` on
packed sequences (CPU), torch-autograd twins of the attention decoder,
exhaustive hypothesis enumeration for the beam searches, and the closed form
of TF-Adam.  ``tests/golden/make_golden.py`` freezes those witnesses' outputs
as fixtures.
"""
from .nabu_oracle import *  # noqa: F401,F403
