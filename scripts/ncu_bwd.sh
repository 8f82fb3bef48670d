# source-level ncu capture of the scorer-backward kernels (one B200)
TAG=${1:-r01f}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
SGS_CUDA_PROFILER=1 timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'edge_score_bwd_d|edge_score_tc2' -c 4 -o /tmp/${TAG}_bwd -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_bwd.ncu-rep --page raw --csv > gpurun_out/${TAG}_bwd_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_bwd.ncu-rep --page source --csv -k regex:edge_score_bwd_df 2>/dev/null | gzip > gpurun_out/${TAG}_bf_source.csv.gz
ncu -i /tmp/${TAG}_bwd.ncu-rep --page source --csv -k regex:edge_score_bwd_dw 2>/dev/null | gzip > gpurun_out/${TAG}_bw_source.csv.gz
ncu -i /tmp/${TAG}_bwd.ncu-rep --page source --csv -k regex:edge_score_tc2 2>/dev/null | gzip > gpurun_out/${TAG}_tc2_source.csv.gz
ls -la gpurun_out/${TAG}*
