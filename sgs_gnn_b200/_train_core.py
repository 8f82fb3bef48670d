"""One training epoch of the learned-sparsifier pipelines, shared by training_hybrid.py and
training_straight_through.py.  Control flow, branch conditions, optimiser stepping and the
returned tuple mirror the reference line by line (training_hybrid.py:7-189,
training_straight_through.py:7-176); the arithmetic runs in libsgs_b200 kernels.

Differences that are deliberate:
  * host syncs per learned step: ONE 8-scalar D2H for the conditional gate (+ sampler validity
    flags) and the final loss read, instead of two logits D2H copies + sklearn + two .item();
  * the hybrid scorer backward touches only the q sampled edges (the only ones with non-zero
    upstream gradient) instead of all E.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import dist as sdist
from . import loader as sloader
from . import ops, sampling, sharded
from .ops import SAMPLE_TRAIN

_softmax_cache = {}


def _softmax_prob(prob):
    """F.softmax(batch.prob) of training_hybrid.py:46, cached per tensor (prob is per-graph)."""
    key = id(prob)
    hit = _softmax_cache.get(key)
    if hit is not None and hit[0]() is prob and hit[1] == prob._version:
        return hit[2]
    s = ops.softmax_f32(prob)
    try:
        _softmax_cache[key] = (weakref.ref(prob, lambda _r, k=key, c=_softmax_cache: c.pop(k, None)), prob._version, s)
    except TypeError:
        pass
    return s


def _has_train(batch):
    flag = getattr(batch, "_sgs_has_train", None)
    if flag is None:
        flag = bool(batch.train_mask.any())
        try:
            batch._sgs_has_train = flag
        except Exception:
            pass
    return flag


def _is_plain_ce(criterion):
    return (isinstance(criterion, nn.CrossEntropyLoss) and criterion.weight is None
            and criterion.reduction == "mean" and criterion.label_smoothing == 0.0
            and criterion.ignore_index == -100)


def _ce(criterion, out, batch, tm_u8, acc=None):
    if _is_plain_ce(criterion):
        return ops.fused_loss(out, batch.y, tm_u8, acc=acc, reg1=False, reg2=False)
    return criterion(out[batch.train_mask], batch.y[batch.train_mask])


def learned_step(pipeline, args, epoch, max_epoch, model, batch, criterion, q, backward_fn, checks=None):
    """One learned step on a device-resident batch.  Returns (loss tensor, update_edge_mlp).
    `checks`: a list the step appends device-side validity flags to (sampler state vectors, the edge list's
    out-of-range flag) that it has NOT read itself; the caller reads them together with the loss value."""
    n = batch.x.size(0)
    g_full = ops.graph_of(batch.edge_index, n)
    oob = g_full.take_oob_flag()
    if oob is not None:
        if checks is not None:
            checks.append(("oob", oob))
        elif int(oob.item()) != 0:
            raise RuntimeError("edge_index contains node ids outside [0, num_nodes)")
    tm_u8 = batch.train_mask.view(torch.uint8)
    scorer = model.edge_prob_mlp
    dp = sdist.is_dist() and bool(getattr(args, "data_parallel", False))
    coef = args.degree_bias_coef

    g_rand = None
    r = None
    if args.conditional or args.sparse_edge_mlp:
        # training_hybrid.py:45-48
        r = sampling.sample_random(_softmax_prob(batch.prob), q, validate=False)
        g_rand = g_full.subgraph(r.sel, ascending=True)

    # pass 1: probabilities of ALL edges (training_hybrid.py:51-64), no autograd tape
    profiler = getattr(model, "gpu_profiler", None)
    ops.seg_begin(profiler, "edge_mlp_pre")
    h_rand = None     # first layer of the random-baseline forward, when it shares the scorer's sweep (see below)
    # (opt-in, SGS_PAIR=1: measured neutral on the Reddit shape once the D = 256 gather was software-pipelined --
    # 3.66 ms for the D = 512 pair sweep vs 2 x 1.5 ms -- see profiles/r02_notes.md)
    pair = (bool(__import__("os").environ.get("SGS_PAIR")) and args.conditional and g_rand is not None
            and hasattr(scorer, "gcn2") and isinstance(scorer.gcn1, type(model.gcn1))
            and float(model.dropout.p) == float(scorer.dropout.p)
            and ops.gcn_conv_pair_available(batch.x, scorer.gcn1.lin.weight, model.gcn1.lin.weight))
    if pair:
        # scorer.gcn1 (model.py:107) and the random baseline's model.gcn1 (model.py:159 via training_hybrid.py:93) are
        # both relu / dropout GCN layers over the SAME random subgraph and the same features: one gather sweep
        h_e, h_rand = ops.gcn_conv_pair(batch.x, scorer.gcn1.lin.weight, scorer.gcn1.bias, model.gcn1.lin.weight,
                                        model.gcn1.bias, g_rand, True, scorer._drop(), ops.next_seed())
        out = scorer.gcn2(h_e, g_rand, None, relu=True)
    else:
        out = scorer.embed(batch.x, g_rand if g_rand is not None else g_full)
    ops.seg_end(profiler, "edge_mlp_pre")
    ops.seg_begin(profiler, "edge_score")
    seed_sc = ops.next_seed()
    p_drop = scorer._drop()
    fc1, fc2 = scorer.fc1, scorer.fc2
    with torch.no_grad():
        p_full = ops.edge_score_forward(out.detach(), g_full, fc1.weight, fc1.bias, fc2.weight.reshape(-1),
                                        fc2.bias.reshape(-1), None, p_drop, seed_sc)
    ops.seg_end(profiler, "edge_score")

    # sample (training_hybrid.py:72-83)
    smp = sampling.sample_edges(p_full, batch.prob, q, False, coef, validate=False)
    g_s = g_full.subgraph(smp.sel, ascending=True)
    if pipeline == "hybrid":
        # edge_probs_full[mask] with grad (training_hybrid.py:86): backward over the q edges only
        p_sel = ops.gather_selected(p_full, None, smp.sel, ops.SAMPLE_RAW, 0.0, None)[0]
        p_s = scorer.score(out, g_full, ids=smp.sel, precomputed=p_sel, seed=seed_sc)
    elif pipeline == "two_pass":
        # pass 3 (training_two_pass.py:75-81): the scorer runs again on the sampled subgraph itself -- message
        # passing over the q sampled edges, scoring of those q edges, fresh dropout masks, gradients enabled
        out2 = scorer.embed(batch.x, g_s)
        p_s = scorer.score(out2, g_s)
    elif pipeline == "straight_through":
        # sampled_edge_weight = (p * st)[mask].clamp(0,1) with dense gradient
        # (training_straight_through.py:60-75, sampling.py:137-155)
        p_full_g = scorer.score(out, g_full, precomputed=p_full, seed=seed_sc)
        p_s = ops.StraightThroughWeightsFn.apply(p_full_g, batch.prob, smp.sel, smp.S, SAMPLE_TRAIN, coef)
    else:
        raise ValueError(pipeline)

    learned_out = model(batch, g_s, p_s)

    update_edge_mlp = True
    acc_l = acc_r = None
    random_out = None
    with_edges = bool(args.reg1 or args.reg2)
    if args.conditional:
        random_out = model.gcn2(h_rand, g_rand, None) if h_rand is not None else model(batch, g_rand)
        # calculate_f1 x2 (training_hybrid.py:94-95): micro-F1 == accuracy; same denominator, so the
        # strict `>` of :98 compares the two correct-counts.
        acc_l = ops.loss_forward(learned_out.detach(), batch.y, tm_u8, g_s if with_edges else None,
                                 p_s.detach() if with_edges else None)
        acc_r = ops.loss_forward(random_out.detach(), batch.y, tm_u8)
        gate = torch.stack([acc_l[2], acc_r[2]])
        if dp:  # data-parallel over batches: one global decision on the summed correct-counts
            sdist.dist.all_reduce(gate)
        host = torch.cat([acc_l, acc_r, smp.state.double(), r.state.double(), gate]).cpu()
        _check_sampler(host[16:24], q)
        _check_sampler(host[24:32], q)
        update_edge_mlp = bool(host[32] > host[33])
        if getattr(args, "force_branch", None) == "learned":   # bench only: always time the full (learned-wins) step
            update_edge_mlp = True
        elif getattr(args, "force_branch", None) == "random":  # tests only: a random-wins step on demand
            update_edge_mlp = False
    elif checks is not None:
        # no gate read on this path: the sampler's invalid-input / count flags ride on the caller's loss read
        checks.append(("sampler", smp.state))
        if r is not None:
            checks.append(("sampler", r.state))
    else:
        _check_sampler(smp.state.cpu(), q)
        if r is not None:
            _check_sampler(r.state.cpu(), q)
    if update_edge_mlp:
        loss = ops.fused_loss(learned_out, batch.y, tm_u8, p_s if with_edges else None, g_s if with_edges else None,
                              args.regularizer1_coef, args.consist_reg_coef, bool(args.reg1), bool(args.reg2),
                              acc=acc_l)
    else:
        loss = _ce(criterion, random_out, batch, tm_u8, acc_r)
    backward_fn(loss)
    return loss, update_edge_mlp


def _defer_oob(batch, checks):
    """Queue the out-of-range flag of batch.edge_index (set once, when the edge list was narrowed to int32)."""
    f = ops.graph_of(batch.edge_index, batch.x.size(0)).take_oob_flag()
    if f is not None:
        checks.append(("oob", f))


def _read_loss_and_checks(loss, checks, q):
    """ONE device->host read for the step's loss value and every deferred validity flag; raises the way the
    reference does (torch.multinomial's RuntimeError, PyG's index assert) -- after the step, without an extra sync."""
    if not checks:
        return loss.item()
    parts = [loss.detach().reshape(1).double()] + [t.reshape(-1).double() for _, t in checks]
    host = torch.cat(parts).cpu()
    off = 1
    for kind, t in checks:
        k = t.numel()
        if kind == "sampler":
            _check_sampler(host[off:off + k], q)
        elif kind == "oob" and int(host[off]) != 0:
            raise RuntimeError("edge_index contains node ids outside [0, num_nodes)")
        off += k
    return float(host[0])


def _check_sampler(st, q):
    if int(st[5]) != 0:
        raise RuntimeError("probability tensor contains either `inf`, `nan` or element < 0")
    if int(st[7]) != q:
        raise RuntimeError(f"sampler selected {int(st[7])} edges, expected {q}")


def train_epoch(pipeline, args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion,
                cluster_loader, q=500, alternate_frequency=1):
    device = args.device
    mode = args.mode
    model.train()
    profiler = getattr(model, "gpu_profiler", None)
    total_loss = 0
    temperature = 1.0
    conditional_update = 0
    total_update = 0

    def _backward(loss):
        ops.seg_begin(profiler, "backward")
        loss.backward()
        ops.seg_end(profiler, "backward")

    dp_mode = sdist.is_dist() and bool(getattr(args, "data_parallel", False)) and mode == "learned"
    # host batches are uploaded one step ahead on a copy stream (loader.prefetch); device batches pass through
    for batch in sloader.prefetch(cluster_loader, device):
        checks = []
        if dp_mode and not isinstance(batch, sharded.ShardedBatch):
            # data-parallel ranks must issue the same collectives: agree on the control flow of this iteration first
            any_tr, all_tr, any_big, all_big = sdist.agree_on_path(_has_train(batch), batch.edge_index.shape[1] > q)
            if any_tr != all_tr or any_big != all_big:
                raise RuntimeError("data-parallel ranks disagree on this iteration's path (a batch without train "
                                   "nodes or with E <= q on some ranks only): give every rank batches of the same "
                                   "kind per iteration (main.py:41-67 cluster batches), or train them on one rank")
        if not _has_train(batch):
            continue
        total_update += 1
        optimizer_edge_prob.zero_grad()
        optimizer_gnn.zero_grad()

        if mode == "learned" and isinstance(batch, sharded.ShardedBatch):
            # one graph sharded by destination range over the ranks; q is the global budget
            if batch.num_edges_global <= q:
                raise RuntimeError("sharded mode expects more edges than the budget q")
            batch = batch.to(device)
            r_ = (args.t_init - args.t_min) / max_epoch
            temperature = max(args.t_min, args.t_init - epoch * r_)
            loss, update_edge_mlp = sharded.learned_step(pipeline, args, epoch, max_epoch, model, batch, criterion, q,
                                                         _backward)
            opts = (optimizer_edge_prob, optimizer_gnn) if update_edge_mlp else (optimizer_gnn,)
            seen, plist = set(), []
            for o in opts:
                for grp in o.param_groups:
                    for prm in grp["params"]:
                        if id(prm) not in seen:
                            seen.add(id(prm))
                            plist.append(prm)
            sharded.allreduce_partial_grads(plist, batch.comm)
            if update_edge_mlp:
                conditional_update += 1
                optimizer_edge_prob.step()
            optimizer_gnn.step()
        elif mode == "learned":
            if batch.edge_index.shape[1] > q:
                batch = batch.to(device)
                # temperature anneal: computed and returned, never used by the sampler
                # (training_hybrid.py:67-70)
                r_ = (args.t_init - args.t_min) / max_epoch
                temperature = max(args.t_min, args.t_init - epoch * r_)
                loss, update_edge_mlp = learned_step(pipeline, args, epoch, max_epoch, model, batch, criterion, q,
                                                     _backward, checks)
                if sdist.is_dist() and bool(getattr(args, "data_parallel", False)):
                    opts = (optimizer_edge_prob, optimizer_gnn) if update_edge_mlp else (optimizer_gnn,)
                    seen, plist = set(), []
                    for o in opts:
                        for grp in o.param_groups:
                            for prm in grp["params"]:
                                if id(prm) not in seen:
                                    seen.add(id(prm))
                                    plist.append(prm)
                    sdist.allreduce_grads(plist)
                if update_edge_mlp:
                    conditional_update += 1
                    optimizer_edge_prob.step()
                    optimizer_gnn.step()
                else:
                    optimizer_gnn.step()
            else:
                batch = batch.to(device)
                _defer_oob(batch, checks)
                out = model(batch, batch.edge_index)
                loss = _ce(criterion, out, batch, batch.train_mask.view(torch.uint8))
                _backward(loss)
                optimizer_gnn.step()
        elif mode in ("random", "edge", "full"):
            batch = batch.to(device)
            _defer_oob(batch, checks)
            ei = batch.edge_index
            if mode == "random" and ei.shape[1] > q:
                ei = sampling.random_edge_sampling(ei, q=q)
            elif mode == "edge" and ei.shape[1] > q:
                g_full = ops.graph_of(batch.edge_index, batch.x.size(0))
                ei = g_full.subgraph(sampling.sample_random(_softmax_prob(batch.prob), q).sel, ascending=True)
            out = model(batch, ei)
            loss = _ce(criterion, out, batch, batch.train_mask.view(torch.uint8))
            _backward(loss)
            optimizer.step()
        else:
            raise ValueError("Invalid mode. Choose 'learned', 'random', or 'full'.")

        total_loss += _read_loss_and_checks(loss, checks, q)

    return total_loss / len(cluster_loader), temperature, conditional_update, total_update
