"""2-GPU NCCL check of the data-parallel step: averaged gradients / identical parameters on both ranks, and the sharded
full-catalogue top-k equal to the single-GPU result.  Skipped when fewer than 2 GPUs are visible."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import rbm_b200
        from rbm_b200.dist import GradSync, sharded_full_catalogue_topk
        dev = "cuda:%d" % rank
        V, Ln, d, nb, h, B = 500, 32, 64, 2, 2, 8
        args = SimpleNamespace(model_code="bert", num_items=V, max_len=Ln, device=dev, model_init_seed=0, bert_num_blocks=nb,
                               bert_num_heads=h, bert_hidden_units=d, bert_dropout=0.0, bert_hidden_dropout=0.0, optimizer="Adam",
                               lr=1e-3, weight_decay=0, momentum=None, decay_step=10, gamma=1.0, num_epochs=1, metric_ks=[10],
                               best_metric="NDCG@10", train_batch_size=B, resume_path=None)
        rng = np.random.RandomState(0)
        tok = rng.randint(1, V + 1, size=(B * world, Ln)).astype(np.int64)
        lab = np.where(rng.rand(B * world, Ln) < 0.2, tok, 0)
        tokm = np.where(lab != 0, V + 1, tok)
        # reference: the whole global batch on one GPU (mean over labels != 0 differs per shard -> compare with the shard-mean average)
        model = rbm_b200.model_factory(args)
        trainer = rbm_b200.trainer_factory(args, model, None, None, None, None)
        trainer.dist_sync = GradSync(model.parameters())
        model.train()
        sl = slice(rank * B, (rank + 1) * B)
        loss = trainer.train_step((torch.from_numpy(tokm[sl]), torch.from_numpy(lab[sl])))
        # after the step every rank holds identical parameters
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        assert all(torch.equal(o, other[0]) for o in other), "parameters diverged across ranks"
        # gradients were averaged: compare rank 0's .grad with the mean of per-rank local grads recomputed without sync
        model2 = rbm_b200.model_factory(args).to(dev).train()
        l2 = model2.loss(torch.from_numpy(tokm[sl]), torch.from_numpy(lab[sl]))
        l2.backward()
        g_local = torch.cat([p.grad.flatten() for p in model2.parameters()])
        dist.all_reduce(g_local)
        g_local /= world
        g_sync = torch.cat([p.grad.flatten() for p in model.parameters()])
        assert torch.allclose(g_sync, g_local, rtol=1e-5, atol=1e-7)
        # sharded full-catalogue top-k == single-GPU top-k (ids exactly, incl. global ids)
        model.eval()
        with torch.no_grad():
            hl = model.last_hidden(torch.from_numpy(tokm[:B]))
            v1, i1 = rbm_b200.ops.score_topk(hl, model.out.weight, model.out.bias, 1, V + 1, 10)
            v2, i2 = sharded_full_catalogue_topk(hl, model.out.weight, model.out.bias, V, 10)
        assert torch.equal(i1, i2) and torch.equal(v1, v2)
        # vocab-parallel cross-entropy (output layer row-sharded over the ranks, replicated rows) == single-GPU fused CE
        from rbm_b200.dist import vocab_parallel_cross_entropy, shard_range
        g = torch.Generator().manual_seed(11)
        n, V1c, dc = 4096, 3417, 64
        hid = torch.randn(n, dc, generator=g).to(dev)
        wfull = (torch.randn(V1c, dc, generator=g) * 0.1).to(dev)
        bfull = (torch.randn(V1c, generator=g) * 0.1).to(dev)
        labc = torch.randint(1, V1c, (n,), generator=g)
        labc[torch.rand(n, generator=g) > 0.15] = 0
        labc = labc.to(dev)
        h1, w1, b1 = hid.clone().requires_grad_(True), wfull.clone().requires_grad_(True), bfull.clone().requires_grad_(True)
        ref = rbm_b200.ops.score_cross_entropy(h1, labc, w1, b1)
        ref.backward()
        lo, hi = shard_range(V1c, rank, world)
        h2, w2, b2 = hid.clone().requires_grad_(True), wfull[lo:hi].clone().requires_grad_(True), bfull[lo:hi].clone().requires_grad_(True)
        out = vocab_parallel_cross_entropy(h2, labc, w2, b2, lo)
        out.backward()
        assert abs(out.item() - ref.item()) < 1e-5 * abs(ref.item()), (out.item(), ref.item())
        assert torch.allclose(h2.grad, h1.grad, rtol=1e-4, atol=1e-7), (h2.grad - h1.grad).abs().max()
        assert torch.allclose(w2.grad, w1.grad[lo:hi], rtol=1e-4, atol=1e-7)
        assert torch.allclose(b2.grad, b1.grad[lo:hi], rtol=1e-4, atol=1e-7)
        # data-parallel rows x vocab-parallel output layer (cfg4 layout): == fused CE of the concatenated batch on one GPU
        from rbm_b200.dist import hybrid_vocab_parallel_loss
        g2 = torch.Generator().manual_seed(23)
        Bl, Lh = 6, 40
        hid_g = torch.randn(world * Bl, Lh, dc, generator=g2).to(dev)
        lab_g = torch.randint(1, V1c, (world * Bl, Lh), generator=g2)
        lab_g[torch.rand(world * Bl, Lh, generator=g2) > (0.1 if rank == 0 else 0.1)] = 0
        lab_g[Bl:][torch.rand((world - 1) * Bl, Lh, generator=g2) > 0.5] = 0  # unequal counts per rank
        lab_g = lab_g.to(dev)
        hg, wg, bg = hid_g.clone().requires_grad_(True), wfull.clone().requires_grad_(True), bfull.clone().requires_grad_(True)
        ref_h = rbm_b200.ops.score_cross_entropy(hg.reshape(-1, dc), lab_g.reshape(-1), wg, bg)
        ref_h.backward()
        slr = slice(rank * Bl, (rank + 1) * Bl)
        for cap_h in (None, 64):
            hl = hid_g[slr].clone().requires_grad_(True)
            ws_, bs_ = wfull[lo:hi].clone().requires_grad_(True), bfull[lo:hi].clone().requires_grad_(True)
            out_h, ovf = hybrid_vocab_parallel_loss(hl, lab_g[slr], ws_, bs_, lo, capacity=cap_h)
            out_h.backward()
            assert not bool(ovf)
            assert abs(out_h.item() - ref_h.item()) < 1e-5 * abs(ref_h.item()), (out_h.item(), ref_h.item())
            assert torch.allclose(hl.grad, hg.grad[slr] * world, rtol=1e-4, atol=1e-7), (hl.grad - hg.grad[slr] * world).abs().max()
            assert torch.allclose(ws_.grad, wg.grad[lo:hi], rtol=1e-4, atol=1e-7)
            assert torch.allclose(bs_.grad, bg.grad[lo:hi], rtol=1e-4, atol=1e-7)
        _, ovf = hybrid_vocab_parallel_loss(hid_g[slr], lab_g[slr], wfull[lo:hi], bfull[lo:hi], lo, capacity=2)
        assert bool(ovf)
        # row-sharded item table, data-parallel sequences (SURVEY 8e input lookup): == the unsharded embedding stage on the
        # concatenated batch -- forward bit for bit, table-shard gradient and (rank-summed) positional gradient to 1e-6
        from rbm_b200.dist import sharded_embedding
        ge = torch.Generator().manual_seed(31)
        Ve, de, Le, Be = 1003, 64, 24, 5
        table_e = torch.randn(Ve, de, generator=ge).to(dev)
        pos_e = torch.randn(Le, de, generator=ge).to(dev)
        tok_e = torch.randint(0, Ve, (world * Be, Le), generator=ge)
        tok_e[:, :3] = 0
        tok_e[1, 5:9] = 7  # repeated item, several contributions to one row
        tok_e = tok_e.to(dev)
        dout_e = torch.randn(world * Be, Le, de, generator=ge).to(dev)
        for scale_e, zp_e, p_e in ((8.0, 1, 0.25), (1.0, 0, 0.0)):
            tf, pf = table_e.clone().requires_grad_(True), pos_e.clone().requires_grad_(True)
            full = rbm_b200.ops.EmbedFn.apply(tok_e, tf, pf, scale_e, zp_e, p_e, 77, 5)
            full.backward(dout_e)
            lo_e, hi_e = shard_range(Ve, rank, world)
            ts, ps = table_e[lo_e:hi_e].clone().requires_grad_(True), pos_e.clone().requires_grad_(True)
            sle = slice(rank * Be, (rank + 1) * Be)
            out_e = sharded_embedding(tok_e[sle], ts, ps, lo_e, Ve, scale_e, zp_e, p_e, 77, 5)
            assert torch.equal(out_e, full[sle]), (out_e - full[sle]).abs().max()
            out_e.backward(dout_e[sle])
            assert torch.allclose(ts.grad, tf.grad[lo_e:hi_e], rtol=1e-6, atol=1e-6), (ts.grad - tf.grad[lo_e:hi_e]).abs().max()
            dp = ps.grad.clone()
            dist.all_reduce(dp)
            assert torch.allclose(dp, pf.grad, rtol=1e-5, atol=1e-5), (dp - pf.grad).abs().max()
        # BASELINE configs[3] layout end to end: data-parallel body, row-sharded token table + output layer.  One optimisation
        # step of the sharded model on the ranks' own batches == the single-GPU model on the concatenated batch: loss, and every
        # gradient (replicated ones after the bucket all-reduce, sharded ones against the rows of the full gradient)
        from rbm_b200.dist import shard_bert_model, replicated_parameters
        args_s = SimpleNamespace(**{**vars(args), "bert_hidden_units": 64, "num_items": 777, "max_len": 24})
        Vs, Ls, Bs = 777, 24, 6
        rs_ = np.random.RandomState(3)
        tok_s = rs_.randint(1, Vs + 1, size=(world * Bs, Ls)).astype(np.int64)
        tok_s[:, :2] = 0
        lab_s = np.where((rs_.rand(world * Bs, Ls) < 0.25) & (tok_s != 0), tok_s, 0)
        lab_s[Bs:][rs_.rand((world - 1) * Bs, Ls) > 0.5] = 0
        tokm_s = np.where(lab_s != 0, Vs + 1, tok_s)
        full_m = rbm_b200.model_factory(args_s)
        full_t = rbm_b200.trainer_factory(args_s, full_m, None, None, None, None)
        full_m.train()
        loss_f = full_t.train_step((torch.from_numpy(tokm_s), torch.from_numpy(lab_s)))
        gfull = {k: p.grad.clone() for k, p in full_m.named_parameters()}
        sh_m = shard_bert_model(rbm_b200.model_factory(args_s).to(dev), capacity=64)
        sh_t = rbm_b200.trainer_factory(args_s, sh_m, None, None, None, None)
        sh_t.dist_sync = GradSync(replicated_parameters(sh_m))
        sh_m.train()
        sls = slice(rank * Bs, (rank + 1) * Bs)
        loss_s = sh_t.train_step((torch.from_numpy(tokm_s[sls]), torch.from_numpy(lab_s[sls])))
        assert not bool(sh_m._shard.overflow)
        assert abs(loss_s.item() - loss_f.item()) < 1e-5 * abs(loss_f.item()), (loss_s.item(), loss_f.item())
        gscale = max(float(g_.abs().max()) for g_ in gfull.values())
        for k, p in sh_m.named_parameters():
            ref_g = gfull[k]
            if getattr(p, "_rbm_sharded", False):
                b0, e0 = shard_range(ref_g.shape[0], rank, world)
                ref_g = ref_g[b0:e0]
            err = float((p.grad - ref_g).abs().max())
            assert err <= 1e-3 * max(float(ref_g.abs().max()), 1e-3 * gscale), (k, err, float(ref_g.abs().max()))
        # full-catalogue top-10 of the sharded model (all-gather of the last hidden rows, shard-local fused scoring with global ids,
        # all-to-all of the lists by user range, merge) == the unsharded model on the same users: ids exactly
        full_e = rbm_b200.model_factory(args_s).to(dev).eval()
        sh_e = shard_bert_model(rbm_b200.model_factory(args_s).to(dev)).eval()
        ev_tok = np.concatenate([tok_s[:, 1:], np.full((world * Bs, 1), Vs + 1, np.int64)], axis=1)  # history + [MASK]
        with torch.no_grad():
            fv, fi = full_e.full_catalogue_topk(torch.from_numpy(ev_tok), 10)
            sv, si = sh_e.full_catalogue_topk(torch.from_numpy(ev_tok[sls]), 10)
        assert torch.equal(si, fi[sls]), (si, fi[sls])
        assert torch.allclose(sv, fv[sls], rtol=1e-6, atol=1e-6)
        # sampled-candidate scores of the sharded BERT4Rec model == the unsharded model's, bit for bit
        cand_b = torch.from_numpy(rs_.randint(1, Vs + 1, size=(world * Bs, 21)).astype(np.int64))
        with torch.no_grad():
            cs_full = full_e.candidate_scores(torch.from_numpy(ev_tok), cand_b)
            cs_sh = sh_e.candidate_scores(torch.from_numpy(ev_tok[sls]), cand_b[sls])
        assert torch.equal(cs_sh, cs_full[sls]), (cs_sh - cs_full[sls]).abs().max()
        # SASRec with the item table row-sharded (shard_sas_model): one optimisation step on the ranks' own batches == the
        # single-GPU model on the concatenated batch (loss 1e-5; every gradient 1e-3 of scale), predict bit for bit, top-10 ids exactly
        from rbm_b200.dist import shard_sas_model
        Vq, Lq, Bq, dq = 611, 20, 5, 64
        a_q = SimpleNamespace(model_code="sas", num_items=Vq, max_len=Lq, device=dev, sas_hidden_units=dq, sas_num_blocks=2, sas_heads=2,
                              sas_dropout=0.0, l2_emb=0.0, optimizer="Adam", lr=1e-3, weight_decay=0, momentum=None, decay_step=10, gamma=1.0,
                              num_epochs=1, metric_ks=[10], best_metric="NDCG@10", train_batch_size=Bq, resume_path=None)
        rq = np.random.RandomState(9)
        seq_q = rq.randint(1, Vq + 1, size=(world * Bq, Lq)).astype(np.int64)
        seq_q[:, :3] = 0
        seq_q[Bq:, :9] = 0  # unequal numbers of live positions per rank
        pos_q = np.where(seq_q != 0, rq.randint(1, Vq + 1, size=seq_q.shape), 0)
        neg_q = np.where(seq_q != 0, rq.randint(0, Vq + 1, size=seq_q.shape), 0)
        torch.manual_seed(21)
        full_q = rbm_b200.model_factory(a_q).to(dev)
        torch.manual_seed(21)
        sh_q = shard_sas_model(rbm_b200.model_factory(a_q).to(dev))
        full_q.train(); sh_q.train()
        lq_full = full_q.loss(seq_q, pos_q, neg_q)
        lq_full.backward()
        slq = slice(rank * Bq, (rank + 1) * Bq)
        lq_sh = sh_q.loss(seq_q[slq], pos_q[slq], neg_q[slq])  # value: the global mean; gradient: this rank's share x world
        assert abs(lq_sh.item() - lq_full.item()) < 1e-5 * abs(lq_full.item()), (lq_sh.item(), lq_full.item())
        lq_sh.backward()
        gs_q = GradSync(replicated_parameters(sh_q))
        gs_q.allreduce_grads()
        gq = {k: p.grad for k, p in full_q.named_parameters()}
        gscale_q = max(float(g_.abs().max()) for g_ in gq.values())
        for k, p in sh_q.named_parameters():
            ref_g = gq[k]
            got_g = p.grad
            if getattr(p, "_rbm_sharded", False):
                b0, e0 = shard_range(ref_g.shape[0], rank, world)
                ref_g = ref_g[b0:e0]
            err = float((got_g - ref_g).abs().max())
            assert err <= 1e-3 * max(float(ref_g.abs().max()), 1e-3 * gscale_q), (k, err, float(ref_g.abs().max()))
        full_q.eval(); sh_q.eval()
        cand_q = rq.randint(1, Vq + 1, size=(world * Bq, 17)).astype(np.int64)
        with torch.no_grad():
            pq_full, pq_sh = full_q.predict(seq_q, cand_q), sh_q.predict(seq_q[slq], cand_q[slq])
            tq_full, tq_sh = full_q.full_catalogue_topk(seq_q, 10), sh_q.full_catalogue_topk(seq_q[slq], 10)
        assert torch.equal(pq_sh, pq_full[slq]), (pq_sh - pq_full[slq]).abs().max()
        assert torch.equal(tq_sh[1], tq_full[1][slq]) and torch.equal(tq_sh[0], tq_full[0][slq])
        # data-parallel step replayed from CUDA graphs (two graphs + one eager all-reduce) == the eager data-parallel step,
        # bit for bit, dropout on
        import copy
        args.bert_dropout = args.bert_hidden_dropout = 0.1
        batches = [(torch.from_numpy(tokm[sl]).to(dev), torch.from_numpy(lab[sl]).to(dev)),
                   (torch.from_numpy(np.ascontiguousarray(tokm[sl][::-1])).to(dev), torch.from_numpy(np.ascontiguousarray(lab[sl][::-1])).to(dev))]
        results = []
        for mode in ("eager", "split"):
            torch.manual_seed(0)  # the models' dropout seed defaults to torch.initial_seed() at construction time
            m = rbm_b200.model_factory(args)
            t = rbm_b200.trainer_factory(args, m, None, None, None, None)
            t.dist_sync = GradSync(m.parameters())
            m.train()
            ls = [t.train_step(batches[0]).item()]
            if mode != "eager":
                t.capture_train_step(batches[1], collective=mode)
            for i in range(1, 5):
                ls.append(t.train_step(batches[i % 2]).item())
            t.release_train_graph()
            ls.append(t.train_step(batches[1]).item())
            results.append((ls, torch.cat([p.detach().flatten() for p in m.parameters()]).clone()))
        assert results[0][0] == results[1][0], (results[0][0], results[1][0])
        assert torch.equal(results[0][1], results[1][1]), "graph-replayed data-parallel step differs from the eager one"
        q.put((rank, "ok"))
    except Exception:  # noqa
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_nccl_data_parallel_and_sharded_topk():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=280) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for r in res:
        print("rank", r[0], r[1])  # the whole traceback (pytest shortens the assertion message)
    assert all(r[1] == "ok" for r in res), res
