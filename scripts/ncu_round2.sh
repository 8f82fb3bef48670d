# Round-2 ncu evidence (run under gpurun on ONE B200).  Reports are written to /tmp and only CSV summaries come back
# (gpurun_out/ is capped at 64 MiB).  Never a bench value: kernels are replayed ~40x.
# usage: bash scripts/ncu_round2.sh TAG
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
# (1) launch list of one timed step (cold-cache, serialised: compare SHARES with bench.py's kernel_time_share)
SGS_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# (2) full metric set + source correlation for the hot kernels
SGS_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'edge_score_tc2|edge_score_bwd_d|loss_edges_fused|topq_|spmm_|sddmm|gemm_tf32' -c 60 -o /tmp/${TAG}_hot -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > gpurun_out/${TAG}_hot_raw.csv 2>/dev/null
for k in edge_score_tc2:k1_ba edge_score_bwd_df:bf edge_score_bwd_dw:bw loss_edges_fused:loss topq_keys_window:topq_keys topq_write:topq_write spmm_h16:spmm_h16 sddmm_h16:sddmm_h16; do
  ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv -k regex:${k%%:*} -c 1 2>/dev/null | gzip > gpurun_out/${TAG}_${k##*:}_source.csv.gz
done
ls -la gpurun_out/${TAG}*
