TAG=r02ah
mkdir -p gpurun_out
for combo in "4 4" "3 4" "5 4" "4 6"; do
set -- $combo
CCB_EOS_ROOT_MINB=$1 CCB_EOS_EPI_MINB=$2 timeout 300 python bench.py --workload config4_milne_bsq_table_3p1d:2e6 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); pk=d['roofline']['per_kernel']
print('root $1 epi $2: thermo %.3f ms  step %.2f' % (pk['04_thermodynamics']['ms_per_launch'], d['ms_per_step']))"
done
