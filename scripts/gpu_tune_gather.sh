# sweep of the gather launch shape (targets per task x records per tile) on three workloads; prints ms per launch
mkdir -p gpurun_out
NTS=${NTS:-"64 128"}; TILES=${TILES:-"32 64 128"}
for w in config2_viscous_gubser_bsq_2p1d:20 config4_milne_bsq_table_3p1d:1e6:5 config3_trento_pbpb_table_2p1d:2e5:5 config5_cartesian_blob_3p1d:10; do
  name=${w%:*}; steps=${w##*:}
  for nt in $NTS; do for tile in $TILES; do
    CCB_GATHER_NT=$nt CCB_GATHER_TILE=$tile timeout 300 python bench.py --workload $name --steps $steps --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['per_kernel']
print('$name nt=$nt tile=$tile  fields %.3f  grad %.3f  step %.3f ms' % (k['03_smooth_fields']['ms_per_launch'], k['09_smooth_gradients']['ms_per_launch'], d['ms_per_step']))"
  done; done
done
