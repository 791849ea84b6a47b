# end-to-end leg (host buffers) of the headline workload against the number of pipeline chunks
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for c in 4 8 16 32 64; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --sustained-seconds 0 --e2e-chunks $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('chunks $c: e2e %.1f TFLOP/s  %.2f ms/step  (copies alone H2D %.1f D2H %.1f both %.1f GB/s)' % (e['value'], e['ms_per_step'], e['pinned_copy_GBs_per_gpu_min']['h2d_alone'], e['pinned_copy_GBs_per_gpu_min']['d2h_alone'], e['pinned_copy_GBs_per_gpu_min']['both_directions_sum']))"
done
