"""rbm_b200 -- B200-native (sm_100a) BERT4Rec / SASRec train step and full-catalogue evaluation behind the model /
trainer / metric API of Furyton/Recommender-Baseline-Model (``NerualNetwork/bert4rec&sas4rec``).

Import as ``rbm_b200`` (the directory name carries the reference's name and is not a python identifier; the repo-root
``rbm_b200.py`` shim registers it).  All math runs in hand-written CUDA kernels reached through the C ABI of
``librbm_b200.so`` (include/rbm.h); there is no CPU or PyTorch fallback.
"""
from . import lib  # noqa: F401
from .models import MODELS, model_factory, BERTModel, SASModel  # noqa: F401
from .trainers import TRAINERS, trainer_factory, BERTTrainer, SASTrainer  # noqa: F401
from .trainers.utils import recalls_ndcgs_and_mrr_for_ks  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .dataloaders import DATALOADERS, dataloader_factory  # noqa: F401

__all__ = ["MODELS", "model_factory", "BERTModel", "SASModel", "TRAINERS", "trainer_factory", "BERTTrainer", "SASTrainer",
           "recalls_ndcgs_and_mrr_for_ks", "FusedAdam", "DATALOADERS", "dataloader_factory", "lib"]
