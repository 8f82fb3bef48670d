CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r01i_plain.log 2>&1 &&
SGS_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01i_launches.csv $CMD > gpurun_out/r01i_ncu1.log 2>&1
tail -2 gpurun_out/r01i_ncu1.log | cut -c1-200
