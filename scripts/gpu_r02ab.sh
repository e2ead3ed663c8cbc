TAG=r02ab
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for w in config4_milne_bsq_table_3p1d config1_ideal_gubser_2p1d config2_viscous_gubser_bsq_2p1d config3_trento_pbpb_table_2p1d config5_cartesian_blob_3p1d; do
timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$w.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("${w:0:7}: %.3f ms/step  %.3e e2e %.3e fp64 %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["fp64"]["frac"] or 0), {k[:3]: round(v["ms_per_launch"],3) for k,v in pk.items()})
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${TAG}_bench_$w.err").read()[-500:])
PY
done
