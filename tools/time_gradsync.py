"""torchrun micro-benchmark: GradSync.allreduce_grads (pack -> one NCCL all-reduce -> unpack) on the cfg2 parameter set."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import rbm_b200
from types import SimpleNamespace
from rbm_b200.dist import GradSync

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
# cfg2 BERT4Rec parameter set: token table [V+2, d], positions, 2 blocks of (4 attention linears, 2 FFN linears, 2 LN), out layer
V, d, Ln = 3416, 64, 200
shapes = [(V + 2, d), (Ln, d)]
for _ in range(2):
    shapes += [(d, d), (d,)] * 4 + [(4 * d, d), (4 * d,), (d, 4 * d), (d,)] + [(d,), (d,)] * 2
shapes += [(V + 1, d), (V + 1,)]
model = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(*s_, device="cuda:%d" % lr)) for s_ in shapes])
for p in model.parameters():
    p.grad = torch.randn_like(p)
gs = GradSync(model.parameters())
for _ in range(5):
    gs.allreduce_grads()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record()
t0 = time.perf_counter()
for _ in range(50):
    gs.allreduce_grads()
t1 = time.perf_counter()
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print("GradSync: %.1f us per call on the device, %.1f us per call of host time, bucket %.2f MB, %d tensors" % (
        e0.elapsed_time(e1) * 1000 / 50, (t1 - t0) * 1e6 / 50, gs.total * 4 / 1e6, len(gs.params)))
dist.destroy_process_group()
