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
        self._graph = None     # set by capture_train_step()
        self._graph_b = None
        self._graph_row_cap = 0  # SASRec live-row capacity the graph was captured with (0 = dense path)

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
        if self._graph is not None:
            return self._replay(batch)
        self.optimizer.zero_grad()
        loss = self.calculate_loss(batch)
        loss.backward()
        if self.dist_sync is not None:
            self.dist_sync.allreduce_grads()
        self.optimizer.step()
        return loss

    # ---------------------------------------------------------------- the same step as ONE CUDA graph
    def capture_train_step(self, example_batch, warmup: int = 3, collective: str = "split"):
        """Capture zero_grad + loss + backward + optimizer step into a CUDA graph; later ``train_step`` calls copy the batch
        into the graph's static input buffers and replay it (one launch instead of ~200: at the reference's batch sizes of
        64-128 the step is host-bound otherwise).  Dropout sites and Adam's bias correction advance through the library's
        device-side step counter (``rbm_set_step_counter``), so replay number j reproduces eager step s0 + j bit for bit.
        Fused Adam, fixed batch shape and learning rate (call again after the scheduler changes ``lr``).

        Data-parallel (``dist_sync`` set; every rank must call this at the same point): ``collective="split"`` captures TWO
        graphs -- [zero_grad, loss, backward, bucket pack] and [bucket unpack, Adam] -- and issues the one NCCL all-reduce of
        the flat bucket eagerly between their replays (three host calls per step, so the ranks do not drift apart on host
        jitter); ``collective="graph"`` captures the all-reduce too (one graph)."""
        import copy
        from .. import lib as L
        if collective not in ("split", "graph"):
            raise ValueError("collective must be 'split' or 'graph'")
        if not isinstance(self.optimizer, FusedAdam):
            raise RuntimeError("capture_train_step needs the fused Adam optimizer")
        if getattr(self.model, "_shard", None) is not None:
            raise RuntimeError("capture_train_step: a row-sharded model exchanges rows inside the loss and its backward; those "
                               "collectives are not captured -- use the eager step")
        sync = self.dist_sync if (self.dist_sync is not None and self.dist_sync.world > 1) else None
        dev = torch.device(self.device)
        static = tuple(torch.as_tensor(x).to(dev).clone() for x in example_batch)
        # SASRec runs its token-wise layers on the non-padding rows only (models/sas.py, csrc/rows.cu): a graph needs a fixed row
        # capacity, chosen from the example batch (0 = dense path); _replay sends a batch that does not fit through an eager step
        self._graph_row_cap = 0
        if hasattr(self.model, "row_capacity_for"):
            self._graph_row_cap = self.model.row_capacity_for(*static)
            self.model._row_cap = self._graph_row_cap
        # warm-up on a side stream (lazy one-time work: function attributes, context binding, workspaces, Adam state, the
        # NCCL communicator), then put model / optimizer / step counters back so that the captured step is the next real one
        model_sd = copy.deepcopy(self.model.state_dict())
        opt_sd = copy.deepcopy(self.optimizer.state_dict())
        step0 = self.model._step
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        try:
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self.optimizer.zero_grad()
                    loss = self.calculate_loss(static)
                    loss.backward()
                    if sync is not None:
                        sync.allreduce_grads()
                    self.optimizer.step()
        except BaseException:
            self.model._row_cap = None
            raise
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.model.load_state_dict(model_sd)
        self.optimizer.load_state_dict(opt_sd)
        self.model._step = step0
        self.optimizer.init_state()  # state buffers must exist before capture (a lazily created zero buffer would be re-zeroed by every replay)
        # the device-side step counter: 0 now, +1 at the end of every replay
        self._graph_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        L.check(L.load().rbm_set_step_counter(self._graph_counter.data_ptr()), "set_step_counter")
        # gradients are (re)allocated INSIDE the capture, as in an eager step (autograd hands each produced gradient over to
        # .grad: no zero fill and no accumulation add per parameter); their addresses are final once allocated, and the
        # descriptor tables that hold them reach the device through pre-allocated pinned buffers (a copy node per replay)
        self.optimizer.prepare_capture()
        if sync is not None:
            sync.prepare_capture()
        # NCCL's watchdog thread polls CUDA events while we capture: only this thread's calls belong to the capture
        mode = {"capture_error_mode": "thread_local"} if sync is not None else {}
        launches0 = L.launch_count
        graph = torch.cuda.CUDAGraph()
        graph_b = None
        try:
            loss = self._capture(graph, static, sync, collective, mode)
            if sync is not None and collective == "split":
                graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph_b, pool=graph.pool(), **mode):
                    sync.unpack()
                    self.optimizer.step()
                    self._graph_counter.add_(1)
        except BaseException:
            L.load().rbm_set_step_counter(None)
            self.model._row_cap = None
            raise
        self.model._row_cap = None  # eager steps (odd-shaped or over-capacity batches) size themselves
        self._graph_launches = L.launch_count - launches0
        self._graph, self._graph_b, self._graph_static, self._graph_loss, self._graph_lr = graph, graph_b, static, loss, self.get_lr()
        L.workspaces.freeze()  # the graph has the workspace addresses baked in: they must outlive every replay
        return self

    def _capture(self, graph, static, sync, collective, mode):
        with torch.cuda.graph(graph, **mode):
            self.optimizer.zero_grad(set_to_none=True)
            loss = self.calculate_loss(static)
            loss.backward()
            if sync is not None:
                sync.pack()
                if collective == "graph":
                    sync.allreduce_bucket()
                    sync.unpack()
            if sync is None or collective == "graph":
                self.optimizer.step()
                self._graph_counter.add_(1)
        return loss

    def release_train_graph(self):
        """Back to eager steps (the python-side step counters are advanced by the number of replays)."""
        if self._graph is None:
            return
        from .. import lib as L
        done = int(self._graph_counter.item())
        L.check(L.load().rbm_set_step_counter(None), "set_step_counter")
        # the capture itself advanced the python counters by one step; the replays did the rest on the device
        self.model._step += done - 1
        for st in self.optimizer.state.values():
            if "step" in st:
                st["step"] += done - 1
        self._graph = self._graph_b = None
        L.workspaces.unfreeze()

    def _graph_steps_done(self) -> int:
        """Optimisation steps taken since capture that the python-side counters (``optimizer.state[*]['step']``,
        ``model._step``) have not seen yet: they advance on the device while a graph is active."""
        return 0 if self._graph is None else int(self._graph_counter.item()) - 1

    def _odd_shaped_step(self, batch):
        """A batch whose shape differs from the captured one (typically the last batch of an epoch: like the reference's
        ``DataLoader``, the loaders have no ``drop_last``).  One eager step at exactly the position the next replay would have
        taken: the device-side step counter stays installed (kernels add it to their dropout sites and to Adam's step), so
        the python-side counters are first moved back by the one step the capture advanced them, and the device counter is
        advanced afterwards."""
        self.model._step -= 1
        for st in self.optimizer.state.values():
            if "step" in st:
                st["step"] -= 1
        self.optimizer.zero_grad()
        loss = self.calculate_loss(batch)
        loss.backward()
        if self.dist_sync is not None:
            self.dist_sync.allreduce_grads()
        self.optimizer.step()
        self._graph_counter.add_(1)
        return loss

    def _replay(self, batch):
        from .. import lib as L
        if self.get_lr() != self._graph_lr:
            raise RuntimeError("the learning rate changed since capture_train_step(): release_train_graph() and capture again")
        batch = tuple(torch.as_tensor(x) for x in batch)
        if len(batch) != len(self._graph_static) or any(tuple(s.shape) != tuple(d.shape) for s, d in zip(batch, self._graph_static)):
            return self._odd_shaped_step(batch)  # never copy_ a mismatching batch: a 1-row remainder would broadcast silently
        for dst, src in zip(self._graph_static, batch):
            dst.copy_(src, non_blocking=True)
        # a batch with more live / labelled rows than the captured capacity takes an eager step instead (the static buffers then just
        # hold an unused copy).  The rows are counted on the device copy: the graph cannot start before the copies have landed
        # anyway, and a few small reductions + one read-back cost less than counting a [B, L] int64 batch on the host.
        if self._graph_row_cap and self.model.live_row_count(*self._graph_static) > self._graph_row_cap:
            return self._odd_shaped_step(batch)
        self._graph.replay()
        if self._graph_b is not None:
            self.dist_sync.allreduce_bucket()
            self._graph_b.replay()
        L.count_launches(self._graph_launches)
        return self._graph_loss

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
        self._check_shard_overflow()
        print(tot_loss)
        return accum_iter

    def _check_shard_overflow(self):
        """Row-sharded model (rbm_b200.dist.shard_bert_model) with a slot ``capacity``: a rank that had more labelled rows than
        slots dropped the excess from loss and gradients -- the '1 rank == N ranks' contract is broken, so fail loudly (checked
        once per epoch: the flag is accumulated on the device, no per-step sync)."""
        sh = getattr(self.model, "_shard", None)
        if sh is None or getattr(sh, "overflow_any", None) is None:
            return
        flag = sh.overflow_any.to(torch.int32).reshape(1)
        if sh.world > 1:  # collective: every rank reaches the end of the epoch; a rank must also learn of the OTHERS' overflow
            import torch.distributed as dist
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=sh.group)
        sh.overflow_any = None
        if bool(flag.item()):
            raise RuntimeError("hybrid_vocab_parallel_loss: some rank had more labelled rows than its slot capacity (%s) during this "
                               "epoch; rows were dropped from the loss. Raise `capacity` in shard_bert_model() or pass None." % sh.capacity)

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
        """NN/trainers/base.py:255-259, plus ``dropout_step`` (the position of the dropout stream; restored by
        ``rbm_b200.checkpoint.load_checkpoint``).  While a CUDA graph is active the real step lives in the device-side
        counter: it is folded into the saved Adam steps and the dropout step here."""
        extra = self._graph_steps_done()
        osd = self.optimizer.state_dict()
        if extra:
            osd = {"state": {k: {f: (v + extra if f == "step" else v) for f, v in st.items()} for k, st in osd["state"].items()},
                   "param_groups": osd["param_groups"]}
        return {STATE_DICT_KEY: self.model.state_dict(), OPTIMIZER_STATE_DICT_KEY: osd, "dropout_step": self.model._step + extra}
