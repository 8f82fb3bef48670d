# N-GPU box: (N = 2) NCCL / peer parity tests; headline line; barrier latency
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_sharded.py tests/test_gpu_gcn.py -m gpu -q -k "nccl or sddmm or fp16 or degree or pair or edge" > gpurun_out/r02_sharded_nccl_test.log 2>&1; tail -3 gpurun_out/r02_sharded_nccl_test.log
  python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r2o_sharded_all.log 2>&1; tail -3 gpurun_out/r2o_sharded_all.log
fi
$TR scripts/barrier_latency.py 2>&1 | grep "us per barrier" | tee gpurun_out/r02_barrier_${N}gpu.txt
$TR bench.py --gpus $N --steps 8 --warmup 3 --no-cpu 2> gpurun_out/r2o_bench$N.err | grep '^{' > gpurun_out/r2o_bench$N.json; tail -c 300 gpurun_out/r2o_bench$N.err
python - <<PY
import json
for line in open("gpurun_out/r2o_bench$N.json"):
    d=json.loads(line)
    print($N, round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), d["sel_checksum"]["hash"], d["clocks"])
    print("    ", d["kernel_time_share"])
PY
