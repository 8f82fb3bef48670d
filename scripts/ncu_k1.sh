CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r01c_plain.log 2>&1 &&
SGS_CUDA_PROFILER=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'edge_score_tc2' -c 1 -o /tmp/r01c_k1 -f $CMD > gpurun_out/r01c_ncu.log 2>&1
ncu -i /tmp/r01c_k1.ncu-rep --page raw --csv > gpurun_out/r01c_k1_raw.csv 2>/dev/null
ncu -i /tmp/r01c_k1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r01c_k1_source.csv.gz
ncu -i /tmp/r01c_k1.ncu-rep --page details 2>/dev/null | head -400 > gpurun_out/r01c_k1_details.txt
ls -la gpurun_out
