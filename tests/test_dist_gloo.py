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
        # ---- data-parallel batch x vocab-parallel output layer: fixed-capacity slots of the labelled rows, all-gather,
        #      gradient slices back to the owners (the compaction and the scoring are CUDA kernels: GPU test)
        torch.manual_seed(7)
        n_loc, dd, Vh, cap = 24, 8, 50, 10
        H_glob = torch.randn(world * n_loc, dd)
        lab_glob = torch.where(torch.rand(world * n_loc) < 0.25, torch.randint(1, Vh, (world * n_loc,)), torch.zeros((), dtype=torch.int64))
        lab_glob[5] = 7  # at least one label
        Wf = torch.randn(Vh, dd)
        Hl = H_glob[rank * n_loc:(rank + 1) * n_loc].clone()
        ll = lab_glob[rank * n_loc:(rank + 1) * n_loc]
        nz = torch.nonzero(ll).flatten()
        rows = torch.full((n_loc,), 12345, dtype=torch.int32)  # garbage beyond count, as the kernel leaves it
        tg = torch.full((n_loc,), -99, dtype=torch.int64)
        rows[:nz.numel()], tg[:nz.numel()] = nz.to(torch.int32), ll[nz]
        rows[nz.numel():] = 3
        live, rows_c, labels_c, overflow = rd.masked_row_slots(torch.tensor([nz.numel()], dtype=torch.int32), rows, tg, cap)
        assert not bool(overflow) and int(live.sum()) == nz.numel() and (labels_c[~live] == 0).all() and (rows_c[~live] == 0).all()
        h_all, l_all = rd.gather_slots(Hl, live, rows_c, labels_c)
        assert h_all.shape == (world * cap, dd) and torch.equal(l_all[l_all != 0].sort().values, lab_glob[lab_glob != 0].sort().values)
        h_all = h_all.clone().requires_grad_(True)
        loss = torch.nn.functional.cross_entropy(h_all @ Wf.t(), l_all, ignore_index=0)
        Hg = H_glob.clone().requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(Hg @ Wf.t(), lab_glob, ignore_index=0)
        assert abs(float(loss) - float(ref)) < 1e-6
        loss.backward(); ref.backward()
        dh = rd.scatter_slot_grads(h_all.grad, rows_c, live, n_loc, cap, rank, float(world))
        assert torch.allclose(dh, Hg.grad[rank * n_loc:(rank + 1) * n_loc] * world, atol=1e-7)
        _, _, _, ovf = rd.masked_row_slots(torch.tensor([nz.numel()], dtype=torch.int32), rows, tg, max(1, nz.numel() - 1))
        assert bool(ovf)
        # ---- row-sharded item table: scatter keys of a shard (owned ids -> local rows, everything else -> the skipped key)
        Vt = 23
        lo_t, hi_t = rd.shard_range(Vt, rank, world)
        tok_all = torch.tensor([0, 1, 5, 11, 12, 22, 0, 12, 3], dtype=torch.int64)
        keys = rd.shard_keys(tok_all, lo_t, hi_t)
        for t_, k_ in zip(tok_all.tolist(), keys.tolist()):
            assert k_ == (t_ - lo_t if (lo_t <= t_ < hi_t and t_ != 0) else hi_t - lo_t)
        owned = torch.zeros(Vt, dtype=torch.int64)
        owned[tok_all[keys < hi_t - lo_t]] = 1
        dist.all_reduce(owned)
        assert owned[0] == 0 and set(torch.nonzero(owned).flatten().tolist()) == {1, 3, 5, 11, 12, 22} and int(owned.max()) == 1
        # ---- sampled-candidate scoring / SASRec pos-neg scoring against a row-sharded table (CPU twin of
        #      dist.sharded_candidate_scores / ShardedSasScoreFn: the dot products are CUDA kernels, the exchange is this):
        #      owned ids -> local rows, the rest -> row 0 with a zero mask; the sum over ranks has ONE non-zero addend per element
        torch.manual_seed(13)
        Vc, dc2, Bc, Cc = 41, 6, 5, 7
        table = torch.randn(Vc, dc2)
        bias_c = torch.randn(Vc)
        h_glob = torch.randn(world * Bc, dc2)
        cand_glob = torch.randint(0, Vc, (world * Bc, Cc))
        lo_c, hi_c = rd.shard_range(Vc, rank, world)
        loc, own = rd.owned_local_ids(cand_glob, lo_c, hi_c)
        assert bool(((loc >= 0) & (loc < hi_c - lo_c)).all()) and bool((loc[~own] == 0).all())
        assert torch.equal(loc[own] + lo_c, cand_glob[own])
        shard_t, shard_b = table[lo_c:hi_c], bias_c[lo_c:hi_c]
        part = ((shard_t[loc] * h_glob.unsqueeze(1)).sum(-1) + shard_b[loc]) * own.float()
        cnt_own = own.long().clone()
        dist.all_reduce(cnt_own)
        assert bool((cnt_own == 1).all())  # every candidate is owned by exactly one rank
        mine = torch.empty(Bc, Cc)
        dist.reduce_scatter_tensor(mine, part.contiguous())
        ref_c = (table[cand_glob] * h_glob.unsqueeze(1)).sum(-1) + bias_c[cand_glob]
        assert torch.equal(mine, ref_c[rank * Bc:(rank + 1) * Bc])  # one non-zero addend: bit for bit
        # ---- top-k list exchange by user range (dist.sharded_model_topk): rank s holds the lists of ALL users against shard s;
        #      all_to_all hands rank r the lists of ITS users from every shard, in shard order
        Ut, kt = 4, 3
        vals_s = torch.arange(world * Ut * kt, dtype=torch.float32).reshape(world * Ut, kt) + 1000 * rank
        recv = torch.empty_like(vals_s)
        dist.all_to_all_single(recv, vals_s.contiguous())
        recv = recv.view(world, Ut, kt)
        for s_ in range(world):
            expect = torch.arange(world * Ut * kt, dtype=torch.float32).reshape(world * Ut, kt)[rank * Ut:(rank + 1) * Ut] + 1000 * s_
            assert torch.equal(recv[s_], expect)
        # ---- sharded SASRec loss (models/sas.py): value = count-weighted mean of the ranks' means = the global mean;
        #      gradient = this rank's share x world (GradSync divides by world)
        z = torch.randn(world * 9)
        live_g = torch.rand(world * 9) < 0.6
        live_g[0] = True
        zl = z[rank * 9:(rank + 1) * 9].clone().requires_grad_(True)
        ll_ = live_g[rank * 9:(rank + 1) * 9]
        l_loc = torch.nn.functional.softplus(-zl[ll_]).mean() if bool(ll_.any()) else zl.sum() * 0
        cnt_l = ll_.sum().float().reshape(1)
        both = torch.cat([l_loc.detach().reshape(1) * cnt_l, cnt_l])
        dist.all_reduce(both)
        share = l_loc * (cnt_l / both[1] * world).reshape(())
        l_glob = (both[0] / both[1]).reshape(()) + (share - share.detach())
        zg = z.clone().requires_grad_(True)
        ref_l = torch.nn.functional.softplus(-zg[live_g]).mean()
        ref_l.backward()
        l_glob.backward()
        assert abs(float(l_glob) - float(ref_l)) < 1e-6
        assert torch.allclose(zl.grad / world, zg.grad[rank * 9:(rank + 1) * 9], atol=1e-7)
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


# ------------------------------------------------------------------------------------ sharded checkpoints
class _Toy(torch.nn.Module):
    """Parameter names of the reference's SASRec tree, small: a row-sharded item table and replicated dense tensors."""

    def __init__(self, rows, d, V1):
        super().__init__()
        self.sas = torch.nn.Module()
        self.sas.item_emb = torch.nn.Embedding(rows, d)
        self.sas.pos_emb = torch.nn.Embedding(5, d)
        self.sas.last_layernorm = torch.nn.LayerNorm(d)


def _adam_steps(model, opt, seed, n=3, rows=None):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n):
        for name, p in model.named_parameters():
            full = torch.randn(*((rows[name],) + tuple(p.shape[1:]) if rows and name in rows else p.shape), generator=g)
            p.grad = full if not (rows and name in rows) else full[rows["_begin"]:rows["_end"]].clone()
        opt.step()


def _ckpt_worker(rank, world, port, q, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rbm_b200 import checkpoint as ck
        from rbm_b200.dist import shard_range
        V1, d = 11, 4  # 11 rows over 2 ranks: blocks of 6 and 5
        key = "sas.item_emb.weight"
        assert key in ck.ROW_SHARDED["sas"]
        torch.manual_seed(0)
        full = _Toy(V1, d, V1)
        opt_full = torch.optim.Adam(full.parameters(), lr=1e-2)
        _adam_steps(full, opt_full, seed=1)
        ref_file = os.path.join(tmpdir, "reference_format.pth")
        if rank == 0:  # what the reference's loggers write (NN/loggers.py:8-9, NN/trainers/base.py:255-259)
            torch.save({"model_state_dict": full.state_dict(), "optimizer_state_dict": opt_full.state_dict(), "epoch": 3}, ref_file)
        dist.barrier()
        b, e = shard_range(V1, rank, world)
        # ---- scatter on load: a reference-format file into a model that holds its row block only
        torch.manual_seed(100 + rank)
        mine = _Toy(e - b, d, V1)
        opt = torch.optim.Adam(mine.parameters(), lr=1e-2)
        epoch = ck.load_checkpoint(ref_file, mine, opt, sharded_keys=[key])
        assert epoch == 3
        assert torch.equal(mine.sas.item_emb.weight, full.sas.item_emb.weight[b:e])
        assert torch.equal(mine.sas.pos_emb.weight, full.sas.pos_emb.weight)
        st, stf = opt.state[mine.sas.item_emb.weight], opt_full.state[full.sas.item_emb.weight]
        assert torch.equal(st["exp_avg"], stf["exp_avg"][b:e]) and torch.equal(st["exp_avg_sq"], stf["exp_avg_sq"][b:e])
        assert float(st["step"]) == float(stf["step"]) == 3.0
        # ---- training continues identically on the shard (Adam is element-wise) ...
        _adam_steps(full, opt_full, seed=2, n=2)
        _adam_steps(mine, opt, seed=2, n=2, rows={key: V1, "_begin": b, "_end": e})
        assert torch.equal(mine.sas.item_emb.weight, full.sas.item_emb.weight[b:e])
        # ---- ... and gather on save writes ONE file with the reference's full shapes
        out_file = os.path.join(tmpdir, "gathered.pth")
        ck.save_checkpoint(out_file, mine, opt, epoch=5, sharded_keys=[key], total_rows={key: V1})
        got = torch.load(out_file, map_location="cpu", weights_only=False)
        # the reference's three keys (NN/config.py:3-4, NN/loggers.py:54) + the dropout-stream position (an extra key the reference ignores)
        assert set(got) == {"model_state_dict", "optimizer_state_dict", "epoch", "dropout_step"} and got["epoch"] == 5
        assert list(got["model_state_dict"]) == list(full.state_dict())
        for k, v in full.state_dict().items():
            assert torch.equal(got["model_state_dict"][k], v), k
        fo = opt_full.state_dict()
        assert got["optimizer_state_dict"]["param_groups"] == fo["param_groups"]
        for idx, stt in fo["state"].items():
            for f, v in stt.items():
                assert torch.equal(torch.as_tensor(got["optimizer_state_dict"]["state"][idx][f]), torch.as_tensor(v)), (idx, f)
        fresh = _Toy(V1, d, V1)
        fresh.load_state_dict(got["model_state_dict"])  # the reference's resume path (NN/trainers/base.py:29)
        fresh_opt = torch.optim.Adam(fresh.parameters(), lr=1e-2)
        fresh_opt.load_state_dict(got["optimizer_state_dict"])  # the line the reference leaves commented out (:30)
        # ---- the real BERT4Rec parameter tree: shard_bert_model (row blocks of the token table / output layer, in place) and
        #      gather-on-save give back exactly the unsharded state_dict; replicated_parameters leaves the shards out
        from types import SimpleNamespace
        import rbm_b200
        from rbm_b200.dist import shard_bert_model, replicated_parameters
        a = SimpleNamespace(model_code="bert", num_items=37, max_len=8, device="cpu", model_init_seed=4, bert_num_blocks=1,
                            bert_num_heads=2, bert_hidden_units=16, bert_dropout=0.0, bert_hidden_dropout=0.0)
        whole = rbm_b200.model_factory(a)
        part = shard_bert_model(rbm_b200.model_factory(a))
        keys = ck.ROW_SHARDED["bert"]
        sd_w, sd_p = whole.state_dict(), part.state_dict()
        assert list(sd_w) == list(sd_p)
        for k_, v_ in sd_w.items():
            if k_ in keys:
                b_, e_ = shard_range(v_.shape[0], rank, world)
                assert torch.equal(sd_p[k_], v_[b_:e_]), k_
            else:
                assert torch.equal(sd_p[k_], v_), k_
        rep = replicated_parameters(part)
        assert len(rep) == len(list(part.parameters())) - 3 and all(not getattr(p_, "_rbm_sharded", False) for p_ in rep)
        assert part._shard.tok_rows == 39 and part._shard.out_rows == 38 and part._shard.world == world
        f2 = os.path.join(tmpdir, "bert_gathered.pth")
        ck.save_checkpoint(f2, part, None, epoch=1, sharded_keys=keys, total_rows={k_: sd_w[k_].shape[0] for k_ in keys})
        got2 = torch.load(f2, map_location="cpu", weights_only=False)["model_state_dict"]
        for k_, v_ in sd_w.items():
            assert torch.equal(got2[k_], v_), k_
        again = shard_bert_model(rbm_b200.model_factory(SimpleNamespace(**{**vars(a), "model_init_seed": 9})))
        ck.load_checkpoint(f2, again, None, sharded_keys=keys)
        for k_, v_ in again.state_dict().items():
            assert torch.equal(v_, sd_p[k_]), k_
        q.put((rank, "ok"))
    except Exception:  # noqa
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_checkpoint_roundtrip_gloo(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ckpt_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res


def test_checkpoint_single_process(tmp_path):
    """world == 1: save_checkpoint / load_checkpoint are the plain reference-format round trip."""
    from rbm_b200 import checkpoint as ck
    torch.manual_seed(0)
    m = _Toy(7, 4, 7)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    _adam_steps(m, opt, seed=4)
    f = str(tmp_path / "one.pth")
    ck.save_checkpoint(f, m, opt, epoch=2)
    m2 = _Toy(7, 4, 7)
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-2)
    assert ck.load_checkpoint(f, m2, opt2) == 2
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert torch.equal(opt2.state[m2.sas.item_emb.weight]["exp_avg_sq"], opt.state[m.sas.item_emb.weight]["exp_avg_sq"])
