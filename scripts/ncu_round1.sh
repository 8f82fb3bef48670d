# Round-1 ncu evidence (run under gpurun on ONE B200).  Reports are written to /tmp and only CSV
# summaries come back (gpurun_out/ is capped at 64 MiB).  Never a bench value: kernels are replayed ~40x.
# usage: bash scripts/ncu_round1.sh TAG [nospmm]
TAG=${1:-r01z}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
# (1) launch list of one timed step (cold-cache, serialised: compare SHARES with bench.py's kernel_time_share)
SGS_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# (2) full metric set + source correlation for the hot kernels (the sampler's kernels come first in a step)
SGS_CUDA_PROFILER=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'edge_score_tc2|edge_score_bwd_d|loss_edges_fused|topq_' -c 34 -o /tmp/${TAG}_hot -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > gpurun_out/${TAG}_hot_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv -k regex:edge_score_tc2 2>/dev/null | gzip > gpurun_out/${TAG}_k1_ba_source.csv.gz
ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv -k regex:edge_score_bwd_df 2>/dev/null | gzip > gpurun_out/${TAG}_bf_source.csv.gz
ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv -k regex:edge_score_bwd_dw 2>/dev/null | gzip > gpurun_out/${TAG}_bw_source.csv.gz
ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv -k regex:loss_edges_fused 2>/dev/null | gzip > gpurun_out/${TAG}_loss_source.csv.gz
ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv -k regex:topq_keys_window 2>/dev/null | gzip > gpurun_out/${TAG}_topq_keys_source.csv.gz
if [ "$2" != "nospmm" ]; then
  SGS_CUDA_PROFILER=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_kernel|sddmm|gemm_tf32|gemm_fp32' -c 10 -o /tmp/${TAG}_spmm -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
  ncu -i /tmp/${TAG}_spmm.ncu-rep --page raw --csv > gpurun_out/${TAG}_spmm_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_spmm.ncu-rep --page source --csv -k regex:spmm_kernel -c 1 2>/dev/null | gzip > gpurun_out/${TAG}_spmm_source.csv.gz
fi
ls -la gpurun_out/${TAG}*
