"""Model registry -- same surface as NN/models/__init__.py:4-12."""
from .bert import BERTModel
from .sas import SASModel

MODELS = {
    BERTModel.code(): BERTModel,
    SASModel.code(): SASModel,
}


def model_factory(args):
    return MODELS[args.model_code](args)
