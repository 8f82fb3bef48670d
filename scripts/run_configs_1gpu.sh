# BASELINE.json configs that need one GPU (lines go to gpurun_out/r02_configs_1gpu.jsonl; copied to profiles/ afterwards)
OUT=gpurun_out/r02_configs_1gpu.jsonl
: > $OUT
run() { python bench.py --no-cpu "$@" 2>> gpurun_out/r02_configs_1gpu.err | grep '^{' >> $OUT; }
run --steps 10 --warmup 3 --gate natural                                     # Reddit shape, the reference's own gate
run --steps 5 --warmup 3 --pipeline straight_through --no-e2e                 # config 3 at N = 1
run --steps 3 --warmup 3 --precision tf32 --no-e2e                            # tensor-core fp32-parity mode
run --steps 5 --warmup 3 --workload ogbn-products                             # config 5 at N = 1
run --steps 20 --warmup 5 --workload smallcora                                # config 0
run --steps 3 --warmup 3 --clusters 230 --no-e2e                              # the reference's own regime (main.py:41-67): 230 cluster batches
for w in amazon-ratings arxiv-year; do                                        # config 4: 10-50 % edge budget
  for sp in 0.1 0.2 0.3 0.4 0.5; do run --steps 20 --warmup 5 --workload $w --sample-perc $sp --no-e2e; done
done
# the reference itself: stock torch-eager on this GPU and on the host cores, arxiv-year shape
python bench.py --impl reference --ref-device cuda --workload arxiv-year --cpu-scale 1.0 --steps 3 --warmup 1 2>> gpurun_out/r02_configs_1gpu.err | grep '^{' >> $OUT
python bench.py --impl reference --workload arxiv-year --cpu-scale 1.0 --steps 2 --warmup 1 2>> gpurun_out/r02_configs_1gpu.err | grep '^{' >> $OUT
python - <<'PY'
import json
for line in open("gpurun_out/r02_configs_1gpu.jsonl"):
    d = json.loads(line)
    c = d["config"]
    print(d.get("impl", "b200"), c.get("workload"), c.get("pipeline"), c.get("sample_perc"), c.get("gate", "")[:8], c.get("scorer_precision"),
          "ms/step", round(d["ms_per_step"], 3), "value", f'{d["value"]:.3e}', "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step", 0), 2),
          "learned", c.get("learned_wins_steps"))
PY
