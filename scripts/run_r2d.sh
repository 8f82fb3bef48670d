# round-2 check at N = 1 and N = 2: new tests, e2e with the prefetching loader, sel_checksum equality, sharded ST
python -m pytest tests/test_gpu_step.py tests/test_gpu_gcn.py tests/test_gpu_tc.py tests/test_gpu_sampler.py -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; tail -4 gpurun_out/r2d_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2d_bench1.json 2> gpurun_out/r2d_bench1.err; tail -c 400 gpurun_out/r2d_bench1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2d_bench2.json 2> gpurun_out/r2d_bench2.err; tail -c 400 gpurun_out/r2d_bench2.err
$TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --pipeline straight_through > gpurun_out/r2d_bench2_st.json 2> gpurun_out/r2d_bench2_st.err; tail -c 400 gpurun_out/r2d_bench2_st.err
for f in r2d_bench1 r2d_bench2 r2d_bench2_st; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json"))
    print("$f", round(d["ms_per_step"],2), "e2e", d.get("e2e") and {k:d["e2e"][k] for k in ("ms_per_step","serial_ms_per_step","h2d_bytes_per_step")}, d["sel_checksum"] and d["sel_checksum"]["hash"], d["sel_checksum"] and d["sel_checksum"]["tau_bits"])
    print("   ", d["kernel_time_share"])
except Exception as e: print("$f", "failed", e)
PY
done
