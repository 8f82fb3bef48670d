# row-pointer upload: exactness test + the prefetched-training test + one e2e line
python -m pytest tests/test_gpu_step.py -m gpu -q -k "row_pointer or prefetched" > gpurun_out/r2t_tests.log 2>&1; tail -2 gpurun_out/r2t_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2t.err | grep '^{' > gpurun_out/r2t_bench.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench.json").readline()); e=d["e2e"]
print(round(d["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "serial", round(e["serial_ms_per_step"],2), "h2d", e["h2d_bytes_per_step"], d["sel_checksum"]["hash"])
PY
