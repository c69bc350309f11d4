"""World-size-2 gloo tests (CPU) of the multi-rank host logic: flat-bucket gradient averaging, vocab shard ranges and
the top-k list exchange.  The CUDA kernels on either side of the collectives are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import metrics as om


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rbm_b200
        from rbm_b200 import dist as rd
        # ---- flat-bucket gradient averaging (layout + one all-reduce); packing itself is a CUDA kernel (GPU test)
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 4097, 12, 8192)]
        gs = rd.GradSync(params)
        assert gs.world == world and gs.offsets == [0, 8, 4108, 4120] and gs.total == 4120 + 8192
        assert gs.cmap.tolist() == [[0, 0], [1, 0], [1, 1], [2, 0], [3, 0], [3, 1]]
        grads = [torch.full((p.numel(),), float(rank + 1) * (i + 1)) for i, p in enumerate(params)]
        for g, o in zip(grads, gs.offsets):
            gs.bucket[o:o + g.numel()] = g / world
        gs.allreduce_bucket()
        for i, (p, o) in enumerate(zip(params, gs.offsets)):
            expect = (i + 1) * sum(r + 1 for r in range(world)) / world
            assert torch.allclose(gs.bucket[o:o + p.numel()], torch.full((p.numel(),), expect))
        # ---- vocab shards tile the table exactly
        V = 1001
        ranges = [rd.shard_range(V, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == V and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        # ---- top-k exchange: every rank ends up with all shards' lists; merged result is shard-count invariant
        rng = np.random.RandomState(5)
        U, k = 9, 10
        scores = np.round(rng.randn(U, V).astype(np.float32), 2)
        b, e = ranges[rank]
        v, i = om.topk_canonical(scores[:, b:e], k, id_offset=1 + b)
        gv, gi = rd.gather_topk(torch.from_numpy(v), torch.from_numpy(i))
        assert gv.shape == (world, U, k)
        mv, mi = om.topk_merge(gv.numpy(), gi.numpy(), k)
        rv, ri = om.topk_canonical(scores, k, id_offset=1)
        np.testing.assert_array_equal(mi, ri)
        np.testing.assert_array_equal(mv, rv)
        # ---- vocab-parallel cross-entropy: per-shard (log-sum-exp, loss) -> all-gather -> global LSE / loss == full CE
        torch.manual_seed(3)
        P, V1, d, cap = 29, 203, 16, 40
        hrows, wfull, bfull = torch.randn(P, d), torch.randn(V1, d), torch.randn(V1)
        tgt = torch.randint(1, V1, (P,))
        logits = hrows @ wfull.t() + bfull
        lo, hi = rd.shard_range(V1, rank, world)
        ls = torch.logsumexp(logits[:, lo:hi], 1)
        tl = torch.where((tgt >= lo) & (tgt < hi), logits[torch.arange(P), tgt], torch.zeros(P))
        lse_s = torch.full((cap,), float("inf") if rank == 0 else float("nan"))  # entries beyond `count` are garbage by contract
        lse_s[:P] = ls
        loss_s = (ls - tl).mean().reshape(1)
        lse_list = [torch.empty_like(lse_s) for _ in range(world)]
        loss_list = [torch.empty_like(loss_s) for _ in range(world)]
        dist.all_gather(lse_list, lse_s)
        dist.all_gather(loss_list, loss_s)
        LSE, loss = rd.combine_shard_lse(torch.stack(lse_list), torch.cat(loss_list), torch.tensor([P], dtype=torch.int32))
        assert abs(float(loss) - float(torch.nn.functional.cross_entropy(logits, tgt))) < 1e-5
        assert torch.allclose(LSE[:P], torch.logsumexp(logits, 1), atol=1e-5)
        q.put((rank, "ok"))
    except Exception as ex:  # noqa
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    import __graft_entry__ as entry
    entry.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
