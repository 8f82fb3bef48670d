# N-GPU box: sharded NCCL/peer parity test (N = 2 only) + BASELINE configs 3 / 5 at N GPUs
N=${1:-2}
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_sharded.py tests/test_gpu_gcn.py -m gpu -q -k "nccl or sddmm or fp16 or degree or pair or edge" > gpurun_out/r02_sharded_nccl_test.log 2>&1; tail -3 gpurun_out/r02_sharded_nccl_test.log
fi
bash scripts/run_configs_multigpu.sh $N
