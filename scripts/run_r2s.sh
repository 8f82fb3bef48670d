# last check of the final tree: GPU tests, smoke, small-shape + headline lines (raw stream getter in ops._stream)
python -m pytest tests -m gpu -q > gpurun_out/r2s_tests.log 2>&1; tail -2 gpurun_out/r2s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; tail -2 gpurun_out/r2s_smoke.log
for w in smallcora arxiv-year; do
  python bench.py --steps 20 --warmup 5 --workload $w --no-cpu --no-e2e 2>> gpurun_out/r2s.err | grep '^{' >> gpurun_out/r2s_small.jsonl
done
python bench.py --steps 5 --warmup 3 --no-cpu 2>> gpurun_out/r2s.err | grep '^{' >> gpurun_out/r2s_small.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2s_small.jsonl"):
    d=json.loads(l); print(d["config"]["workload"][:24], round(d["ms_per_step"],3), d["e2e"] and round(d["e2e"]["ms_per_step"],2), d["clocks"].get("samples"))
PY
