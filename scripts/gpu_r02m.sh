TAG=r02m
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_bench_config4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_config4.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("config4: %.2f ms/step " % d["ms_per_step"], {k[:2]: round(v["ms_per_launch"],2) for k,v in pk.items()}, "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_list.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|eom_rk_kernel|eos_chain_stage_kernel|eos_local_kernel|permute_columns|prologue_kernel|radix_scatter" -s 40 -c 14 -o gpurun_out/${TAG}_prof_stage_config4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
ls -la gpurun_out/${TAG}_*
