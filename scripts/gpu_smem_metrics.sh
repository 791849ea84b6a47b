# usage: bash scripts/gpu_smem_metrics.sh variant...  -- shared-memory wavefront / conflict counters of the backward main kernel (ncu, one launch)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
M=l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_shared_ld.sum,l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum
for v in "$@"; do
  export FASN_LIBRARY=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n/libfasn_$v.so
  [ "$v" = "base" ] && export FASN_LIBRARY=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n/libfasn.so
  timeout 300 ncu --metrics $M --clock-control none -k regex:fasn_bwd_kernel -s 3 -c 1 --csv --log-file gpurun_out/smem_$v.csv python bench.py --workload ${WL:-c3} --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
  echo "== $v"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/smem_$v.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print("  %-75s %s" % (r[-3], r[-1]))
PY
done
