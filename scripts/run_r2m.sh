N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus $N --steps 8 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/r2m_bench$N.err | grep '^{' > gpurun_out/r2m_bench$N.json; tail -c 300 gpurun_out/r2m_bench$N.err
python - <<PY
import json
for line in open("gpurun_out/r2m_bench$N.json"):
    d=json.loads(line)
    print($N, round(d["ms_per_step"],2), d["sel_checksum"]["hash"])
    print("    ", d["kernel_time_share"])
PY
