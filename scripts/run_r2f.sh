# 8-GPU box: host / kernel profile of rank 0 of the sharded hybrid step (debug aid, never a bench value)
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
SGS_TORCH_PROFILE=gpurun_out/r2f_tprof8 SGS_CPROFILE=gpurun_out/r2f_cprof8 $TR8 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench8_prof.json 2> gpurun_out/r2f_bench8_prof.err
rm -f gpurun_out/r2f_tprof8.rank[1-7].txt gpurun_out/r2f_cprof8.rank[1-7].txt
head -c 3000 gpurun_out/r2f_tprof8.rank0.txt | cut -c1-180
