TAG=r02l
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
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
run hitmask A=1
run nomask CCB_HITMASK=0

(time timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
for w in config1_ideal_gubser_2p1d config2_viscous_gubser_bsq_2p1d config3_trento_pbpb_table_2p1d config5_cartesian_blob_3p1d; do
  for hm in 1 0; do
    CCB_HITMASK=$hm timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${w}_hm$hm.json 2> gpurun_out/${TAG}_bench_${w}_hm$hm.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${w}_hm$hm.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("$w hm=$hm: %.3f ms/step " % d["ms_per_step"], {k[:2]: round(v["ms_per_launch"],3) for k,v in pk.items()})
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${TAG}_bench_${w}_hm$hm.err").read()[-600:])
PY
  done
done
