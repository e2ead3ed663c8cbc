TAG=r02e
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_bench_config4.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_config4.json").read().strip().splitlines()[-1])
    print("config4: %.3e p-steps/s  %.2f ms/step" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"],2) for k, v in d["roofline"]["per_kernel"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_bench_config4.err").read()[-1500:])
PY
