import os, sys, time, torch, torch.nn as nn
sys.path.insert(0, '.')
import bench
from sgs_gnn_b200 import ops, synth, training_hybrid, _lib
dev = torch.device('cuda:0')
ops.set_precision(gemm='fp32', scorer='fp16')
batch = synth.make_graph('reddit', seed=42, device=dev)
q = int(batch.num_edges * 0.2)
model, og, oe, oa = bench.build_model(batch.x.size(1), batch.num_classes, dev, 0.3)
args = bench.make_args(dev, 0.3)
crit = nn.CrossEntropyLoss()
for ep in range(3):
    training_hybrid.train(args, ep + 1, 1000, model, og, oe, oa, crit, [batch], q=q)
torch.cuda.synchronize()
for ep in range(4):
    st0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    with ops.KernelTimer() as kt:
        training_hybrid.train(args, 10 + ep, 1000, model, og, oe, oa, crit, [batch], q=q)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tot = kt.totals_ms()
    st1 = torch.cuda.memory_stats()
    ksum = sum(v[0] for v in tot.values())
    print(f"step wall {1e3*(t1-t0):.1f} ms, timed kernels {ksum:.1f} ms, cudaMalloc segments +{st1['segment.all.allocated']-st0['segment.all.allocated']}, frees +{st1['segment.all.freed']-st0['segment.all.freed']}, retries +{st1['num_alloc_retries']-st0['num_alloc_retries']}, peak {torch.cuda.max_memory_allocated()/2**30:.1f} GiB reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB")
# host-side profile of one step
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
training_hybrid.train(args, 20, 1000, model, og, oe, oa, crit, [batch], q=q)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
