TAG=r02ac
mkdir -p gpurun_out
for mode in 3 4 0; do
for w in config4_milne_bsq_table_3p1d config3_trento_pbpb_table_2p1d; do
CCB_EOS_FAST=$mode timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${w}_m$mode.json 2> gpurun_out/${TAG}_bench_${w}_m$mode.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${w}_m$mode.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("fast=$mode ${w:0:7}: %.3f ms/step  thermo %.3f" % (d["ms_per_step"], pk["04_thermodynamics"]["ms_per_launch"]), d["config"].get("eos_chain_members_rank0"))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${TAG}_bench_${w}_m$mode.err").read()[-500:])
PY
done
done
(time timeout 2000 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
