# e2e spread with the host thread / pinned memory bound to the GPU's NUMA node: same command four times + the default line
run() { tag=$1; shift; "$@" 2> gpurun_out/r2r_$tag.err | grep '^{' > gpurun_out/r2r_$tag.json; python - <<PY
import json
d=json.loads(open("gpurun_out/r2r_$tag.json").readline()); e=d["e2e"]
print("$tag", round(d["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "serial", round(e["serial_ms_per_step"],2), d["clocks"], e["how"][-60:], d.get("cpu_baseline") and (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
PY
}
run a5 python bench.py --steps 5 --warmup 3 --no-cpu
run b5 python bench.py --steps 5 --warmup 3 --no-cpu
run c10 python bench.py --steps 10 --warmup 3 --no-cpu
run d5 python bench.py --steps 5 --warmup 3 --no-cpu
run e10 python bench.py --steps 10 --warmup 3
nvidia-smi topo -m > gpurun_out/r2r_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" > gpurun_out/r2r_lscpu.txt
