TAG=r02t
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for w in config4_milne_bsq_table_3p1d config1_ideal_gubser_2p1d config2_viscous_gubser_bsq_2p1d; do
timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$w.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("$w: %.3f ms/step " % d["ms_per_step"], {k[:3]: round(v["ms_per_launch"],3) for k,v in pk.items()})
PY
done
