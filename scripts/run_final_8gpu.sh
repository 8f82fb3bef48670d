# 8-GPU box: BASELINE configs 3 / 5 + the headline line, then one torch-profiler / cProfile pass of rank 0 (debug aid:
# where the non-scaling part of the 8-GPU step goes; never a bench value)
bash scripts/run_configs_multigpu.sh 8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
SGS_TORCH_PROFILE=gpurun_out/r02_tprof8 SGS_CPROFILE=gpurun_out/r02_cprof8 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_tprof8.log 2>&1
rm -f gpurun_out/r02_tprof8.rank[1-7].txt gpurun_out/r02_cprof8.rank[1-7].txt
tail -2 gpurun_out/r02_tprof8.log | cut -c1-300
