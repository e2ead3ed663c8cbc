# GPU tests + one bench line per BASELINE config (own arm); outputs under gpurun_out/<tag>_*
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "out of range" > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
for w in config2_viscous_gubser_bsq_2p1d:40 config1_ideal_gubser_2p1d:10 config3_trento_pbpb_table_2p1d:5 config4_milne_bsq_table_3p1d:5 config5_cartesian_blob_3p1d:10; do
  name=${w%%:*}; steps=${w##*:}
  timeout 600 python bench.py --workload $name --steps $steps --warmup 3 ${2:---no-cpu-baseline} > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$name", "%.3e p-steps/s  %.3f ms/step  e2e %.3e  n=%d ngb=%.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["particles"], d["config"]["mean_neighbours"]))
    print("   ", {k: v["ms_per_launch"] for k, v in r["per_kernel"].items()}, "fp64frac", r["fp64"]["frac"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/${TAG}_bench_$name.err").read()[-1500:])
EOF
done
