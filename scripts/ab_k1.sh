run() { timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],1), 'K1', round(d['roofline']['avg_launch_ms'],2), 'learned', d['config']['learned_wins_steps'], {k:round(v*d['ms_per_step'],1) for k,v in list(d['kernel_time_share'].items())[:4]})"; }
timeout 200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -x -q 2>&1 | tail -1
run pair_consumer_fence
SGS_K1_SINGLE_CTA=1 run single_consumer_fence
