"""Trainer registry -- same surface as NN/trainers/__init__.py:4-12."""
from .bert import BERTTrainer
from .sas import SASTrainer

TRAINERS = {
    BERTTrainer.code(): BERTTrainer,
    SASTrainer.code(): SASTrainer,
}


def trainer_factory(args, model, train_loader, val_loader, test_loader, export_root):
    return TRAINERS[args.model_code](args, model, train_loader, val_loader, test_loader, export_root)
