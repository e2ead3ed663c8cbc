TAG=r02c
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -q -x 2>&1 | tail -60) > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
for SH in "64 64 0" "64 32 0" "128 64 0" "64 64 1" "128 64 1"; do
  set -- $SH
  CCB_GATHER_NT=$1 CCB_GATHER_TILE=$2 CCB_GATHER_REGS=$3 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$1_$2_$3.json 2> gpurun_out/${TAG}_bench_$1_$2_$3.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$1_$2_$3.json").read().strip().splitlines()[-1])
    print("nt=$1 tile=$2 regs=$3 config4: %.3e p-steps/s  %.2f ms/step" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"],2) for k, v in d["roofline"]["per_kernel"].items()})
except Exception as e:
    print("bench $SH failed", e); print(open("gpurun_out/${TAG}_bench_$1_$2_$3.err").read()[-1500:])
PY
done
