"""CPU, world_size 2 over gloo: the multi-GPU host logic (destination-range sharding, the distributed
radix-select protocol that only moves digit histograms, gate / gradient all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import extended as ox


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def test_destination_ranges_partition_all_edges():
    from sgs_gnn_b200 import dist as sdist, synth
    b = synth.make_graph("smallcora", seed=2)
    n, e = b.num_nodes, b.num_edges
    for world in (1, 2, 3, 8):
        parts = [sdist.shard_by_destination(b.edge_index, n, world, r) for r in range(world)]
        bounds = parts[0][1]
        assert bounds[0] == 0 and bounds[-1] == n and all(bounds[i] <= bounds[i + 1] for i in range(world))
        ids = torch.cat([p[0] for p in parts])
        assert ids.numel() == e and torch.equal(torch.sort(ids).values, torch.arange(e))
        for r, (idr, _) in enumerate(parts):
            d = b.edge_index[1, idr]
            assert bool(((d >= bounds[r]) & (d < bounds[r + 1])).all())
            assert bool((idr[1:] > idr[:-1]).all())
        sizes = [p[0].numel() for p in parts]
        assert max(sizes) - min(sizes) <= e // world * 0.25 + 200      # balanced by in-degree


def _topq_worker(rank, world, e, q, ties, interleaved, gather_max=None):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _topq_numpy import NumpyTopQOps
    from sgs_gnn_b200 import dist as sdist
    if gather_max is not None:      # force the binary-search fallback of the tie resolution
        sdist.DistributedTopQ.TIE_GATHER_MAX = gather_max
    g = torch.Generator().manual_seed(123)
    p = torch.rand(e, generator=g)
    if ties:
        p = torch.round(p * 4) / 4 + 0.25          # few distinct values -> many threshold ties
        noise = torch.ones(e)
    else:
        noise = ox.exponential_noise(e, g)
    prob = torch.softmax(torch.rand(e, generator=g), 0)
    if interleaved:
        # destination-range shards of a (src,dst)-sorted edge list interleave in edge-id order:
        # ties at the threshold must still go to the lowest GLOBAL ids
        owner = torch.randint(0, world, (e,), generator=g)
        gid = torch.nonzero(owner == rank).flatten()
    else:
        # contiguous id shards (rank order == edge id order)
        gid = torch.arange(rank * e // world, (rank + 1) * e // world)
    r = sdist.DistributedTopQ(NumpyTopQOps()).select_ex(p[gid], prob[gid], noise[gid], q, 0, 0.3,
                                                        gid=gid if interleaved else None)
    assert r.n_global == q and not r.invalid
    gathered = [None] * world
    dist.all_gather_object(gathered, gid[r.sel.long()].tolist())
    if rank == 0:
        got = sorted(i for part in gathered for i in part)
        S = p.sum(dtype=torch.float64).to(torch.float32)
        want = ox.sample_topq(p, prob, q, noise, 0.3, False, S=S)
        assert got == want.sel.tolist()
        assert np.array([r.tau_bits], dtype=np.uint32).view(np.float32)[0] == np.float32(want.tau)


@pytest.mark.parametrize("ties", [False, True])
@pytest.mark.parametrize("interleaved", [False, True])
def test_distributed_radix_select_matches_global_topq(ties, interleaved):
    _run(_topq_worker, 2, 5000, 1200, ties, interleaved)


@pytest.mark.parametrize("gather_max", [None, 4])
def test_distributed_radix_select_three_ranks_interleaved_ties(gather_max):
    _run(_topq_worker, 3, 3000, 700, True, True, gather_max)


def _dp_worker(rank, world):
    from sgs_gnn_b200 import dist as sdist
    assert sdist.is_dist()
    w = torch.nn.Parameter(torch.zeros(3))
    v = torch.nn.Parameter(torch.zeros(2, 2))
    u = torch.nn.Parameter(torch.zeros(1))          # no grad on any rank: must be skipped
    w.grad = torch.full((3,), float(rank + 1))
    v.grad = torch.full((2, 2), float(10 * (rank + 1)))
    sdist.allreduce_grads([w, v, u])
    assert torch.allclose(w.grad, torch.full((3,), 1.5)) and torch.allclose(v.grad, torch.full((2, 2), 15.0))
    assert u.grad is None
    a, b = sdist.allreduce_gate(torch.tensor(3.0 + rank), torch.tensor(5.0))
    assert float(a) == 7.0 and float(b) == 10.0


def test_gradient_and_gate_allreduce():
    _run(_dp_worker, 2)


def _dp_hetero_worker(rank, world, kind):
    """data-parallel ranks whose batches would take different paths (no train nodes / E <= q on one rank
    only) must not issue mismatched collectives: every rank raises the same RuntimeError before any of them."""
    from types import SimpleNamespace
    from sgs_gnn_b200 import _train_core, synth
    b = synth.make_graph(None, seed=3 + rank, n=60, e=400 if (kind != "small" or rank == 0) else 40, f=4, c=2)
    if kind == "notrain" and rank == 1:
        b.train_mask = torch.zeros_like(b.train_mask)
    args = SimpleNamespace(device="cpu", mode="learned", data_parallel=True, conditional=True, sparse_edge_mlp=True,
                           t_init=0.7, t_min=0.5, degree_bias_coef=0.3, reg1=True, reg2=True,
                           regularizer1_coef=1.0, consist_reg_coef=0.5)
    model = torch.nn.Linear(2, 2)
    opt = torch.optim.Adam(model.parameters())
    with pytest.raises(RuntimeError, match="data-parallel ranks disagree"):
        _train_core.train_epoch("hybrid", args, 1, 10, model, opt, opt, opt, torch.nn.CrossEntropyLoss(), [b], q=100)
    dist.barrier()      # both ranks got here: nobody is stuck in a collective


@pytest.mark.parametrize("kind", ["notrain", "small"])
def test_data_parallel_ranks_with_different_paths_fail_together(kind):
    _run(_dp_hetero_worker, 2, kind)


def _dp_none_grad_worker(rank, world):
    """A parameter with a gradient on ONE rank only still all-reduces with the same buffer layout everywhere."""
    from sgs_gnn_b200 import dist as sdist
    w = torch.nn.Parameter(torch.zeros(3))
    v = torch.nn.Parameter(torch.zeros(2))
    w.grad = torch.full((3,), 2.0)
    if rank == 0:
        v.grad = torch.full((2,), 4.0)
    sdist.allreduce_grads([w, v])
    assert torch.allclose(w.grad, torch.full((3,), 2.0))
    assert v.grad is not None and torch.allclose(v.grad, torch.full((2,), 2.0))


def test_gradient_allreduce_with_rank_local_none_grads():
    _run(_dp_none_grad_worker, 2)
