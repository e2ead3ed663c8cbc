TAG=r02ag
mkdir -p gpurun_out
for mode in 1 0; do
CCB_EOS_FAST=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4_m$mode.json 2> gpurun_out/${TAG}_bench_config4_m$mode.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_config4_m$mode.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("fast=$mode: %.3f ms/step  thermo %.3f" % (d["ms_per_step"], pk["04_thermodynamics"]["ms_per_launch"]), {k[:3]: round(v["ms_per_launch"],2) for k,v in pk.items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/${TAG}_bench_config4_m$mode.err").read()[-800:])
PY
done
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
