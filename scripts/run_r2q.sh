# e2e variance check on one box: same command repeated, with / without the clock sampler, 5 and 10 steps
run() { tag=$1; shift; "$@" 2> gpurun_out/r2q_$tag.err | grep '^{' > gpurun_out/r2q_$tag.json; python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_$tag.json").readline()); e=d["e2e"]
print("$tag", round(d["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "serial", round(e["serial_ms_per_step"],2), d["clocks"].get("samples"))
PY
}
run a5 python bench.py --steps 5 --warmup 3 --no-cpu
run b5_noclk env SGS_NO_CLOCKS=1 python bench.py --steps 5 --warmup 3 --no-cpu
run c10 python bench.py --steps 10 --warmup 3 --no-cpu
run d5 python bench.py --steps 5 --warmup 3 --no-cpu
