# 1-GPU box: validate the pair sweep (+ PIPE v2), A/B benches
python -m pytest tests/test_gpu_gcn.py tests/test_gpu_benched_parity.py tests/test_gpu_step.py tests/test_gpu_eval.py tests/test_gpu_scorer_loss.py -m gpu -q > gpurun_out/r2j_tests.log 2>&1; tail -3 gpurun_out/r2j_tests.log
SGS_SPMM_PIPE=0 python -m pytest tests/test_gpu_gcn.py tests/test_gpu_benched_parity.py -m gpu -q > gpurun_out/r2j_tests_pipe.log 2>&1; echo "tests PIPE=0:"; tail -2 gpurun_out/r2j_tests_pipe.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2j_pair.json 2> gpurun_out/r2j_pair.err; tail -c 300 gpurun_out/r2j_pair.err
SGS_NO_PAIR=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2j_nopair.json 2> gpurun_out/r2j_nopair.err
SGS_SPMM_PIPE=0 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2j_pipe.json 2> gpurun_out/r2j_pipe.err
for f in r2j_pair r2j_nopair r2j_pipe; do python - <<PY
import json
for line in open("gpurun_out/$f.json"):
    if line.startswith("{"):
        d=json.loads(line)
        ks={k["kernel"]:(round(k["avg_launch_ms"],3),k["launches"]) for k in d.get("kernels",[])}
        print("$f", round(d["ms_per_step"],2), ks)
        print("    ", d["kernel_time_share"])
PY
done
