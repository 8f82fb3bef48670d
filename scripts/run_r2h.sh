# 1-GPU box: validate + A/B the SpMM variants, sampler write kernel; then the full default bench line
for v in BULK PIPE; do
  env SGS_SPMM_$v=1 python -m pytest tests/test_gpu_gcn.py tests/test_gpu_benched_parity.py tests/test_gpu_step.py tests/test_gpu_sampler.py -m gpu -q -x > gpurun_out/r2h_tests_$v.log 2>&1; echo "tests $v:"; tail -2 gpurun_out/r2h_tests_$v.log
done
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2h_plain.json 2> gpurun_out/r2h_plain.err
SGS_SPMM_PIPE=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2h_pipe.json 2> gpurun_out/r2h_pipe.err
SGS_SPMM_BULK=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2h_bulk.json 2> gpurun_out/r2h_bulk.err; tail -c 300 gpurun_out/r2h_bulk.err
for f in r2h_plain r2h_pipe r2h_bulk; do python - <<PY
import json
for line in open("gpurun_out/$f.json"):
    if line.startswith("{"):
        d=json.loads(line)
        ks={k["kernel"]:round(k["avg_launch_ms"],3) for k in d["kernels"]}
        print("$f", round(d["ms_per_step"],2), ks)
PY
done
