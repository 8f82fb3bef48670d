# 2 GPUs over NCCL: the e2e arm with the compact sharded upload
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 52 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu 2> gpurun_out/r2v.err | grep '^{' > gpurun_out/r2v_bench2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v_bench2.json").readline()); e=d["e2e"]
print(round(d["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "serial", round(e["serial_ms_per_step"],2), "h2d", e["h2d_bytes_per_step"], d["sel_checksum"]["hash"])
PY
tail -c 400 gpurun_out/r2v.err
