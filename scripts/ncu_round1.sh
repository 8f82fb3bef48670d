# Round-1 ncu evidence (run under gpurun on ONE B200).  Reports are written to /tmp and only CSV
# summaries come back (gpurun_out/ is capped at 64 MiB).
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r01b_plain.log 2>&1 &&
SGS_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/r01b_ncu1.log 2>&1
SGS_CUDA_PROFILER=1 timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'edge_score_tc|edge_score_bwd|topq_keys|topq_hist|topq_write|sddmm|loss_edges_bwd' -c 18 -o /tmp/r01b_hot -f $CMD > gpurun_out/r01b_ncu2.log 2>&1
SGS_CUDA_PROFILER=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_kernel|gemm_tf32|gemm_fp32' -c 8 -o /tmp/r01b_spmm -f $CMD > gpurun_out/r01b_ncu3.log 2>&1
for r in r01b_hot r01b_spmm; do
  ncu -i /tmp/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
ncu -i /tmp/r01b_hot.ncu-rep --page source --csv -k regex:edge_score_tc 2>/dev/null | gzip > gpurun_out/r01b_k1_source.csv.gz
ncu -i /tmp/r01b_hot.ncu-rep --page source --csv -k regex:edge_score_bwd_da 2>/dev/null | gzip > gpurun_out/r01b_ba_source.csv.gz
ncu -i /tmp/r01b_hot.ncu-rep --page source --csv -k regex:edge_score_bwd_df 2>/dev/null | gzip > gpurun_out/r01b_bf_source.csv.gz
ncu -i /tmp/r01b_hot.ncu-rep --page source --csv -k regex:edge_score_bwd_dw 2>/dev/null | gzip > gpurun_out/r01b_bw_source.csv.gz
ncu -i /tmp/r01b_spmm.ncu-rep --page source --csv -k regex:spmm_kernel -c 1 2>/dev/null | gzip > gpurun_out/r01b_spmm_source.csv.gz
ls -la gpurun_out/ /tmp/*.ncu-rep
