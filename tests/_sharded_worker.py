"""Worker of tests/test_gpu_sharded.py: one process per rank.  Default: gloo rendezvous, every rank on cuda:0
(collectives are host-side, no kernel waits on another process) -- runs on a one-GPU box.  SGS_TEST_BACKEND=nccl: one
GPU per rank over NCCL, which also exercises the peer-memory slab exchange (csrc/peer.cu, the SpMM epilogue's peer
stores) -- needs as many GPUs as ranks.  Checks that the destination-sharded learned step reproduces the single-GPU
step on the same graph."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_args(dev, conditional, pipeline):
    return SimpleNamespace(device=dev, mode="learned", hybrid_checkpoint=False, conditional=conditional,
                           sparse_edge_mlp=True, t_init=0.7, t_min=0.5, degree_bias_coef=0.3, reg1=True, reg2=True,
                           regularizer1_coef=1.0, consist_reg_coef=0.5, pipeline=pipeline, drop_rate=0.0)


def run(rank, world, port, pipeline, conditional, edge_mlp, n, e, f, c, hdim):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    backend = os.environ.get("SGS_TEST_BACKEND", "gloo")
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import extended as ox
        from sgs_gnn_b200 import _train_core, ops, sampling, sharded, synth
        from sgs_gnn_b200 import dist as sdist
        from sgs_gnn_b200.model import GNNModel
        dev = torch.device("cuda", rank) if backend == "nccl" else torch.device("cuda:0")
        ops.set_precision(gemm="fp32", scorer="fp32")
        b = synth.make_graph(None, seed=11, n=n, e=e, f=f, c=c).to(dev)
        q = int(b.num_edges * 0.25)
        g = torch.Generator().manual_seed(5)
        noise = [ox.exponential_noise(b.num_edges, g).to(dev) for _ in range(2)]
        crit = nn.CrossEntropyLoss()

        def fresh():
            torch.manual_seed(3)
            m = GNNModel(f, hdim, c, 0.0, edge_mlp).to(dev)
            m.train()
            return m

        # ---- the distributed select alone: identical global selection, ties included ----
        comm = sharded.Comm()
        sb = sharded.ShardedBatch(b, comm)
        assert sb.bounds[0] == 0 and sb.bounds[-1] == n

        # ---- the prefetching loader's compact upload (int32 gid, source row as row pointer) is bit-exact ----
        from sgs_gnn_b200 import loader
        plain = sb.to("cpu")
        host = plain.compact().pin_memory()
        assert host._gid32 is not None and host._src_rowptr is not None and host.edge_index.dtype == torch.int32
        e_loc = sb.edge_index.size(1)
        if comm.staged:     # (over NCCL only this rank's slice of x crosses the host link)
            assert host.upload_nbytes() == plain.nbytes() - 16 * e_loc + 4 * (n + 1)
        for up in loader.prefetch([host, host], dev):
            torch.cuda.synchronize()
            for k in sb._TENSORS:
                want_t, got_t = getattr(sb, k), getattr(up, k)
                if k == "edge_index":
                    assert got_t.dtype == torch.int32 and torch.equal(got_t.long(), want_t.long()), k
                else:
                    assert got_t.dtype == want_t.dtype and torch.equal(got_t, want_t), k
        for trial, (pv, nz) in enumerate([(torch.rand(b.num_edges, generator=g).to(dev), noise[0]),
                                          ((torch.round(torch.rand(b.num_edges, generator=g) * 3) / 3 + 0.1).to(dev),
                                           torch.ones(b.num_edges, device=dev))]):
            S = ops.sum_f32(pv)
            want = ops.sample_topq(pv, b.prob, q, ops.SAMPLE_TRAIN, 0.3, noise=nz, S=S)
            r = sdist.DistributedTopQ().select_ex(pv[sb.gid].contiguous(), sb.prob, nz[sb.gid].contiguous(), q,
                                                  ops.SAMPLE_TRAIN, 0.3, S=S, gid=sb.gid)
            mine = sb.gid[r.sel.long()].cpu()
            parts = [None] * world
            dist.all_gather_object(parts, mine.tolist())
            got = sorted(i for p_ in parts for i in p_)
            assert got == want.sel.cpu().tolist(), f"trial {trial}: sharded selection differs"
        if rank == 0 and backend == "nccl":
            print(f"peer-memory slab exchange: {'on' if comm.peer is not None else 'OFF'}", flush=True)
            assert r.n_global == q

        # ---- one learned step: single GPU vs sharded ----
        args = make_args(dev, conditional, pipeline)
        m_ref = fresh()
        sampling.clear_injected()
        sampling.inject_noise([t.clone() for t in noise])
        ops.reset_seed_counter()
        loss_ref, upd_ref = _train_core.learned_step(pipeline, args, 1, 10, m_ref, b, crit, q, lambda l: l.backward())
        m_sh = fresh()
        sampling.clear_injected()
        sampling.inject_noise([t.clone() for t in noise])
        ops.reset_seed_counter()
        loss_sh, upd_sh = sharded.learned_step(pipeline, args, 1, 10, m_sh, sb, crit, q, lambda l: l.backward())
        sharded.allreduce_partial_grads(list(m_sh.parameters()), comm)
        assert upd_ref == upd_sh
        assert abs(float(loss_ref) - float(loss_sh)) <= 1e-5 * max(1.0, abs(float(loss_ref))), (float(loss_ref),
                                                                                                float(loss_sh))
        worst = 0.0
        for (k, pr), (_, ps) in zip(m_ref.named_parameters(), m_sh.named_parameters()):
            if pr.grad is None:
                assert ps.grad is None or float(ps.grad.abs().max()) == 0.0, k
                continue
            assert ps.grad is not None, k
            err = float((pr.grad - ps.grad).abs().max() / (pr.grad.abs().max() + 1e-20))
            worst = max(worst, err)
            assert err <= 1e-4, (k, err)
        # ---- several optimiser steps through train(): learned-wins, then RANDOM-wins, then learned-wins again.  On a
        # random-wins step the scorer's gcn* parameters (members of optimizer_gnn, main.py:100) have grad None; the
        # sharded gradient all-reduce must keep it None or Adam would move them by their stale momentum ----
        if conditional and pipeline == "hybrid":
            from sgs_gnn_b200 import training_hybrid

            def opts(m):
                return (torch.optim.Adam([p_ for n_, p_ in m.named_parameters() if "gcn" in n_], lr=1e-2),
                        torch.optim.Adam([p_ for n_, p_ in m.named_parameters() if "edge_prob_mlp" in n_], lr=1e-2),
                        torch.optim.Adam(m.parameters(), lr=1e-2))
            m_a, m_b = fresh(), fresh()
            o_a, o_b = opts(m_a), opts(m_b)
            g2 = torch.Generator().manual_seed(17)
            for step, branch in enumerate(["learned", "random", "learned"]):
                args.force_branch = branch
                nz = [ox.exponential_noise(b.num_edges, g2).to(dev) for _ in range(2)]
                for m_, o_, batch_ in ((m_a, o_a, b), (m_b, o_b, sb)):
                    sampling.clear_injected()
                    sampling.inject_noise([t.clone() for t in nz])
                    ops.reset_seed_counter()
                    _, _, n_cond, _ = training_hybrid.train(args, 1 + step, 10, m_, o_[0], o_[1], o_[2], crit,
                                                           [batch_], q=q, alternate_frequency=0)
                    assert n_cond == (1 if branch == "learned" else 0)
            args.force_branch = None
            worst_p = 0.0
            for (k, pa), (_, pb) in zip(m_a.named_parameters(), m_b.named_parameters()):
                err = float((pa - pb).abs().max() / (1.0 + pa.abs().max()))
                worst_p = max(worst_p, err)
                assert err <= 2e-4, (k, err)
            if rank == 0:
                print(f"sharded[{world}] 3 steps (learned, random, learned): max param err {worst_p:.2e}", flush=True)
        if rank == 0:
            print(f"sharded[{world}] {pipeline} cond={conditional} {edge_mlp}: loss {float(loss_sh):.6f} "
                  f"(single {float(loss_ref):.6f}), max rel grad err {worst:.2e}, learned={upd_sh}", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = sys.argv[1:]
    run(int(a[0]), int(a[1]), int(a[2]), a[3], a[4] == "1", a[5], *(int(v) for v in a[6:11]))
