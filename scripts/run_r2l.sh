# 2-GPU box: full GPU tests (incl. NCCL / peer parity), 1-GPU bench (BF lookahead), 2-GPU bench
python -m pytest tests -m gpu -q > gpurun_out/r2l_tests.log 2>&1; tail -4 gpurun_out/r2l_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2l_bench1.json 2> gpurun_out/r2l_bench1.err; tail -c 200 gpurun_out/r2l_bench1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
$TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/r2l_bench2.err | grep '^{' > gpurun_out/r2l_bench2.json; tail -c 300 gpurun_out/r2l_bench2.err
for f in r2l_bench1 r2l_bench2; do python - <<PY
import json
for line in open("gpurun_out/$f.json"):
    if line.startswith("{"):
        d=json.loads(line)
        ks={k["kernel"]:(round(k["avg_launch_ms"],3),k["launches"]) for k in d.get("kernels",[])}
        print("$f", round(d["ms_per_step"],2), d["sel_checksum"]["hash"], ks)
        print("    ", d["kernel_time_share"])
PY
done
