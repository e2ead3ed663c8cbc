TAG=r02j
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$name.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("$name: %.2f ms/step " % d["ms_per_step"], {k[:2]: round(v["ms_per_launch"],2) for k,v in pk.items()})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_bench_$name.err").read()[-800:])
PY
}
run base A=1
run eom3 CCB_EOM_MINB=3
run eom4 CCB_EOM_MINB=4
(timeout 1500 python -m pytest tests/test_gpu_gubser_check.py tests/test_slabs.py tests/test_dropin.py -m gpu -q -x -s 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
