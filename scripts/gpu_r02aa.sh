TAG=r02aa
mkdir -p gpurun_out
for regs in 0 2 3; do
for w in config4_milne_bsq_table_3p1d config1_ideal_gubser_2p1d config2_viscous_gubser_bsq_2p1d config3_trento_pbpb_table_2p1d config5_cartesian_blob_3p1d; do
CCB_GATHER_REGS=$regs timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${w}_r$regs.json 2> gpurun_out/${TAG}_bench_${w}_r$regs.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${w}_r$regs.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("regs=$regs ${w:0:7}: %.3f ms/step  grad %.3f fields %.3f" % (d["ms_per_step"], pk["09_smooth_gradients"]["ms_per_launch"], pk["03_smooth_fields"]["ms_per_launch"]))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${TAG}_bench_${w}_r$regs.err").read()[-500:])
PY
done
done
