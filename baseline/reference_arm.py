"""The reference's OWN classes on the host CPU -- the baseline arm of bench.py.

``install()`` (called by ``__graft_entry__.build()`` in the build container, where ``/root/reference`` exists) copies the
reference's python package to the git-ignored ``baseline/_ref/NN`` so that it travels to the GPU box with the snapshot;
nothing of it is ever committed.  The classes are used UNMODIFIED: ``models.model_factory`` builds ``BERTModel`` /
``SASModel``; the trainer objects are created without running ``AbstractTrainer.__init__`` (NN/trainers/base.py:18-52:
it opens TensorBoard writers, an experiment folder and -- BERT -- calls ``dataiter.next()``, which current torch no
longer has; all out of scope) and get exactly the attributes the hot-path methods read.  What is
timed is the reference's ``calculate_loss`` / ``loss.backward()`` / ``optimizer.step()`` sequence of
``train_one_epoch`` (NN/trainers/base.py:114-123) and its ``calculate_metrics`` (NN/trainers/sas.py:56-62).

When ``baseline/_ref`` is absent the callers fall back to the oracle port (``kind: "port"``)."""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/NerualNetwork/bert4rec&sas4rec"
REF_DST = os.path.join(HERE, "_ref", "NN")


def install() -> bool:
    """Copy the reference package (python files only) to baseline/_ref/NN.  No-op when /root/reference is absent."""
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(REF_DST)
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    shutil.copytree(REF_SRC, REF_DST, ignore=shutil.ignore_patterns("Data", "*.PNG", "*.png", "__pycache__", "*.pyc"))
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "models", "__init__.py"))


_mods = None


def modules():
    """(model_factory, BERTTrainer, SASTrainer, recalls_ndcgs_and_mrr_for_ks) of the reference."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("baseline/_ref/NN is missing (run __graft_entry__.build() where /root/reference exists)")
        argv, sys.argv = sys.argv, sys.argv[:1]  # NN/options.py parses sys.argv at import time
        sys.path.insert(0, REF_DST)
        try:
            from models import model_factory
            from trainers.bert import BERTTrainer
            from trainers.sas import SASTrainer
            from trainers.utils import recalls_ndcgs_and_mrr_for_ks
        finally:
            sys.argv = argv
        _mods = (model_factory, BERTTrainer, SASTrainer, recalls_ndcgs_and_mrr_for_ks)
    return _mods


def make_trainer(args):
    """The reference trainer for ``args.model_code`` around a freshly built reference model (CPU), constructor side effects
    skipped (see the module docstring); ``args`` needs the model fields plus optimizer / lr / weight_decay / metric_ks."""
    import torch.nn as nn
    model_factory, BERTTrainer, SASTrainer, _ = modules()
    model = model_factory(args)
    cls = BERTTrainer if args.model_code == "bert" else SASTrainer
    t = object.__new__(cls)
    t.args, t.device, t.model = args, args.device, model.to(args.device)
    t.optimizer = t._create_optimizer()       # NN/trainers/base.py:225-233
    t.metric_ks = args.metric_ks
    if args.model_code == "bert":
        t.ce = nn.CrossEntropyLoss(ignore_index=0)   # NN/trainers/bert.py:11
    else:
        t.bce_criterion = nn.BCEWithLogitsLoss()     # NN/trainers/sas.py:13
        t.l2_emb = args.l2_emb
    return t


def train_step(t, batch) -> float:
    """The body of the reference's epoch loop, NN/trainers/base.py:114-123."""
    t.optimizer.zero_grad()
    loss = t.calculate_loss(batch)
    value = loss.item()
    loss.backward()
    t.optimizer.step()
    return value
