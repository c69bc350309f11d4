"""oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain fp32 torch-CPU / numpy, the arithmetic of the
BERT4Rec and SASRec train step and evaluation path of
Furyton/Recommender-Baseline-Model (``NerualNetwork/bert4rec&sas4rec``, written
``NN/`` below).  It is the *checker* for the CUDA path in
``recommender-baseline-model_b200/``; it is never the thing measured or shipped.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package never does, and
fails loudly when its CUDA library is missing.

Parity status
-------------
The reference has **no tests, golden vectors or known-answer fixtures**
(SURVEY.md section 4), so there is nothing of the reference's own to pin against.
Instead the oracle is pinned against *outputs of the reference itself run in the
build container*: ``tests/golden/make_golden.py`` imports the unmodified
reference modules from ``/root/reference`` and stores small input/output vectors
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle
function against them (bit-exact for integer work, <=1e-6 for fp32).  The
reference's arithmetic that lives in its third-party dependency ``torch``
(pinned ``torch==1.8.1`` in ``NN/requirements.txt:2``; 2.11.0 is installed here)
-- ``nn.MultiheadAttention``, ``nn.LayerNorm``, ``nn.CrossEntropyLoss``,
``nn.BCEWithLogitsLoss``, ``optim.Adam``, ``Tensor.argsort`` -- is restated
explicitly and validated against the installed torch by the same script.
"""
from . import common, bert4rec, sasrec, metrics, optim  # noqa: F401
