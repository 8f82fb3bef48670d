# 1-GPU box: full tests, PIPE variant tests + A/B, the default bench line (with e2e + CPU reference), BASELINE configs
python -m pytest tests -m gpu -q > gpurun_out/r2i_tests.log 2>&1; tail -4 gpurun_out/r2i_tests.log
SGS_SPMM_PIPE=1 python -m pytest tests/test_gpu_gcn.py tests/test_gpu_benched_parity.py -m gpu -q > gpurun_out/r2i_tests_pipe.log 2>&1; echo "tests PIPE:"; tail -2 gpurun_out/r2i_tests_pipe.log
SGS_SPMM_PIPE=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2i_pipe.json 2> gpurun_out/r2i_pipe.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench1.json 2> gpurun_out/r2i_bench1.err; tail -c 300 gpurun_out/r2i_bench1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err
for f in r2i_pipe r2i_bench1 r2i_ref; do python - <<PY
import json
for line in open("gpurun_out/$f.json"):
    if line.startswith("{"):
        d=json.loads(line)
        ks={k["kernel"]:round(k["avg_launch_ms"],3) for k in d.get("kernels",[])}
        print("$f", round(d["ms_per_step"],2), d.get("e2e"), ks, d.get("cpu_baseline"))
PY
done
bash scripts/run_configs_1gpu.sh
