# BASELINE.json configs 3 and 5 on N GPUs of one box: usage  bash scripts/run_configs_multigpu.sh N
N=${1:-8}
OUT=gpurun_out/r02_configs_${N}gpu.jsonl
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520"
run() { $TR bench.py --gpus $N --no-cpu "$@" 2>> gpurun_out/r02_configs_${N}gpu.err | grep '^{' >> $OUT; }
run --steps 5 --warmup 3                                                  # hybrid (headline), with e2e
run --steps 3 --warmup 3 --pipeline straight_through --no-e2e             # config 3
if [ "$N" = "8" ]; then run --steps 5 --warmup 3 --workload ogbn-products --no-e2e; fi   # config 5
python - <<PY
import json
for line in open("$OUT"):
    d = json.loads(line)
    c = d["config"]
    print(d["n_gpus"], c["workload"], c["pipeline"], "ms/step", round(d["ms_per_step"], 3), "value", f'{d["value"]:.3e}',
          "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"], 2), d["sel_checksum"]["hash"])
    print("   ", d["kernel_time_share"])
PY
