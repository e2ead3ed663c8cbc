TAG=r02i
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$name.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("$name: %.2f ms/step  fields %.2f  grad %.2f" % (d["ms_per_step"], pk["03_smooth_fields"]["ms_per_launch"], pk["09_smooth_gradients"]["ms_per_launch"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_bench_$name.err").read()[-800:])
PY
}
run base A=1
run nbuf3 CCB_GATHER_NBUF=3 CCB_FIELDS_NBUF=3
run nbuf3_regs2 CCB_GATHER_NBUF=3 CCB_GATHER_REGS=2 CCB_FIELDS_NBUF=3
run nbuf3_regs1 CCB_GATHER_NBUF=3 CCB_GATHER_REGS=1 CCB_FIELDS_NBUF=3
run nbuf4 CCB_GATHER_NBUF=4 CCB_FIELDS_NBUF=4
run nbuf4_regs1 CCB_GATHER_NBUF=4 CCB_GATHER_REGS=1 CCB_FIELDS_NBUF=4
(CCB_GATHER_NBUF=3 CCB_FIELDS_NBUF=3 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_parity.py tests/test_slabs.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/${TAG}_pytest_nbuf3.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_nbuf3.log
(CCB_GATHER_NBUF=4 CCB_FIELDS_NBUF=4 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_parity.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/${TAG}_pytest_nbuf4.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_nbuf4.log
