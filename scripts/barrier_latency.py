"""Latency of the symmetric-memory barrier the peer slab exchange uses (torchrun, one rank per GPU): 200 back-to-back
barriers timed with CUDA events on rank 0.  Debug aid for DESIGN section 6 (17 barriers per sharded step)."""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = symm_mem.empty(1024, dtype=torch.float32, device=torch.device("cuda", local))
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
for _ in range(20):
    hdl.barrier(channel=0)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    hdl.barrier(channel=0)
e1.record()
torch.cuda.synchronize()
if dist.get_rank() == 0:
    print(f"symmetric-memory barrier, world {dist.get_world_size()}: {1e3 * e0.elapsed_time(e1) / 200:.1f} us per barrier "
          f"(200 back to back, CUDA events)")
dist.destroy_process_group()
