"""AbstractTrainer -- the train / validate / test loops of NN/trainers/base.py:17-262 around the hot path.

Kept: constructor signature, ``train()``, ``validate()``, ``test()``, ``calculate_loss`` / ``calculate_metrics`` hooks,
Adam/SGD + StepLR, ``metric_ks`` / ``best_metric``, checkpoint dict keys (``model_state_dict``,
``optimizer_state_dict``, ``epoch``; NN/config.py:3-4).  Out of scope: TensorBoard writers,
experiment-folder naming, DataParallel -- multi-GPU is one process per GPU (rbm_b200.dist)."""
from __future__ import annotations

import json
import os
from abc import ABCMeta, abstractmethod

import torch
import torch.optim as optim

from ..optim import FusedAdam

STATE_DICT_KEY = 'model_state_dict'
OPTIMIZER_STATE_DICT_KEY = 'optimizer_state_dict'


class AverageMeterSet(object):
    """Mean of per-batch values (NN/utils.py:100-155: ``update(name, value, n=1)``)."""

    def __init__(self):
        self.sums, self.counts = {}, {}

    def update(self, name, value, n=1):
        self.sums[name] = self.sums.get(name, 0.0) + value
        self.counts[name] = self.counts.get(name, 0) + n

    def averages(self):
        return {k: self.sums[k] / self.counts[k] for k in self.sums}


class AbstractTrainer(metaclass=ABCMeta):
    def __init__(self, args, model, train_loader, val_loader, test_loader, export_root):
        self.args = args
        self.device = args.device
        self.model = model.to(self.device)
        self.optimizer = self._create_optimizer()
        if getattr(args, 'resume_path', None) is not None:
            checkpoint = torch.load(args.resume_path, map_location=torch.device(args.device))
            self.model.load_state_dict(checkpoint[STATE_DICT_KEY])
        self.train_loader, self.val_loader, self.test_loader = train_loader, val_loader, test_loader
        self.lr_scheduler = optim.lr_scheduler.StepLR(self.optimizer, step_size=args.decay_step, gamma=args.gamma)
        self.num_epochs = args.num_epochs
        self.metric_ks = args.metric_ks
        self.best_metric = args.best_metric
        self.export_root = export_root
        self.batch_size = args.train_batch_size
        self.best_value = None
        self.dist_sync = None  # set to rbm_b200.dist.GradSync for data-parallel training

    @classmethod
    @abstractmethod
    def code(cls):
        pass

    @abstractmethod
    def calculate_loss(self, batch):
        pass

    @abstractmethod
    def calculate_metrics(self, batch):
        pass

    def close_training(self):
        pass

    # ---------------------------------------------------------------- one optimisation step (base.py:114-123)
    def train_step(self, batch):
        self.optimizer.zero_grad()
        loss = self.calculate_loss(batch)
        loss.backward()
        if self.dist_sync is not None:
            self.dist_sync.allreduce_grads()
        self.optimizer.step()
        return loss

    def train(self):
        accum_iter = 0
        self.validate(0, accum_iter)
        for epoch in range(self.num_epochs):
            accum_iter = self.train_one_epoch(epoch, accum_iter)
            self.validate(epoch, accum_iter)
            self.lr_scheduler.step()
        self.close_training()

    def get_lr(self):
        for param_group in self.optimizer.param_groups:
            return param_group['lr']

    def train_one_epoch(self, epoch, accum_iter):
        self.model.train()
        tot_loss = 0.
        for batch in self.train_loader:
            loss = self.train_step(batch)
            tot_loss += loss.item()
            accum_iter += self.batch_size
        print(tot_loss)
        return accum_iter

    def _evaluate(self, loader):
        self.model.eval()
        meters = AverageMeterSet()
        with torch.no_grad():
            for batch in loader:
                for k, v in self.calculate_metrics(batch).items():
                    meters.update(k, v)
        return meters.averages()

    def validate(self, epoch, accum_iter):
        averages = self._evaluate(self.val_loader)
        print(averages)
        if self.export_root is not None and averages:
            os.makedirs(os.path.join(self.export_root, 'models'), exist_ok=True)
            state = {**self._create_state_dict(), 'epoch': epoch}
            torch.save(state, os.path.join(self.export_root, 'models', 'checkpoint-recent.pth'))
            cur = averages.get(self.best_metric)
            if cur is not None and (self.best_value is None or cur > self.best_value):
                self.best_value = cur
                torch.save(state, os.path.join(self.export_root, 'models', 'best_acc_model.pth'))
        return averages

    def test(self):
        path = getattr(self.args, 'test_model_path', None) or os.path.join(self.export_root, 'models', 'best_acc_model.pth')
        self.model.load_state_dict(torch.load(path, map_location=torch.device(self.device)).get(STATE_DICT_KEY))
        averages = self._evaluate(self.test_loader)
        print(averages)
        if self.export_root is not None:
            os.makedirs(os.path.join(self.export_root, 'logs'), exist_ok=True)
            with open(os.path.join(self.export_root, 'logs', 'test_metrics.json'), 'w') as f:
                json.dump(averages, f)
        return averages

    def _create_optimizer(self):
        args = self.args
        if args.optimizer.lower() == 'adam':
            return FusedAdam(self.model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
        elif args.optimizer.lower() == 'sgd':
            return optim.SGD(self.model.parameters(), lr=args.lr, weight_decay=args.weight_decay, momentum=args.momentum)
        else:
            raise ValueError

    def _create_state_dict(self):
        return {STATE_DICT_KEY: self.model.state_dict(), OPTIMIZER_STATE_DICT_KEY: self.optimizer.state_dict()}
