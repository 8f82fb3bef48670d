# Round-2 closing run on ONE B200: full GPU tests, smoke, the default bench line (with e2e + cpu_baseline), the ncu
# evidence of the same command, the BASELINE.json one-GPU configs.  Outputs under gpurun_out/ (copied to profiles/ here).
python -m pytest tests -m gpu -q > gpurun_out/r2n_tests.log 2>&1; tail -4 gpurun_out/r2n_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -3 gpurun_out/r2n_smoke.log
# A/B: fused edge loss with 8 edges in flight per thread
SGS_LOSS_UN=8 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/r2n_un8.err | grep '^{' > gpurun_out/r2n_un8.json
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/r2n_un4.err | grep '^{' > gpurun_out/r2n_un4.json
for f in r2n_un8 r2n_un4; do python - <<PY
import json
for line in open("gpurun_out/$f.json"):
    d=json.loads(line)
    ks={k["kernel"]:(round(k["avg_launch_ms"],3),k["launches"]) for k in d.get("kernels",[])}
    print("$f", round(d["ms_per_step"],2), d["sel_checksum"]["hash"], ks)
PY
done
# the round's headline line
python bench.py --steps 10 --warmup 3 2> gpurun_out/r02_bench_1gpu.err | grep '^{' > gpurun_out/r02_bench_1gpu.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu.json").readline())
print("FINAL", round(d["ms_per_step"], 2), d["value"], "e2e", d["e2e"], "cpu", d["cpu_baseline"], d["roofline"], d["clocks"])
PY
cp gpurun_out/parity_benched.json gpurun_out/r02_parity_benched.json 2>/dev/null
bash scripts/ncu_round2.sh r02 > gpurun_out/r2n_ncu.log 2>&1; tail -5 gpurun_out/r2n_ncu.log
timeout 900 bash scripts/run_configs_1gpu.sh > gpurun_out/r2n_configs.log 2>&1; tail -30 gpurun_out/r2n_configs.log
