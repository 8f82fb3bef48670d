# 2-GPU box: sharded step over NCCL + peer memory (parity), then bench N = 2 with / without the peer-memory exchange
python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; tail -25 gpurun_out/r2e_tests.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err; tail -c 600 gpurun_out/r2e_bench2.err
SGS_PEER=0 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2e_bench2_nccl.json 2> gpurun_out/r2e_bench2_nccl.err; tail -c 300 gpurun_out/r2e_bench2_nccl.err
for f in r2e_bench2 r2e_bench2_nccl; do python - <<PY
import json
for line in open("gpurun_out/$f.json"):
    if line.startswith("{"):
        d=json.loads(line)
        print("$f", round(d["ms_per_step"],2), d["sel_checksum"]["hash"], d["config"]["learned_wins_steps"])
        print("   ", d["kernel_time_share"])
PY
done
