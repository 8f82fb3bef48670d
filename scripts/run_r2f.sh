# 8-GPU box: hybrid Reddit-shape epoch sharded over 8 and 4 GPUs + host / kernel profile of rank 0 at N = 8
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
$TR8 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench8.json 2> gpurun_out/r2f_bench8.err; tail -c 400 gpurun_out/r2f_bench8.err
SGS_TORCH_PROFILE=gpurun_out/r2f_tprof8 SGS_CPROFILE=gpurun_out/r2f_cprof8 $TR8 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench8_prof.json 2> gpurun_out/r2f_bench8_prof.err
rm -f gpurun_out/r2f_tprof8.rank[1-7].txt gpurun_out/r2f_cprof8.rank[1-7].txt
$TR4 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench4.json 2> gpurun_out/r2f_bench4.err; tail -c 400 gpurun_out/r2f_bench4.err
for f in r2f_bench8 r2f_bench4; do python - <<PY
import json
for line in open("gpurun_out/$f.json"):
    if line.startswith("{"):
        d=json.loads(line)
        print("$f", round(d["ms_per_step"],2), d["sel_checksum"]["hash"], d["config"]["learned_wins_steps"])
        print("   ", d["kernel_time_share"])
PY
done
