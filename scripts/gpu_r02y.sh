TAG=r02y
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke") | tail -5
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("default: %.2f ms/step %.3e  e2e %.3e  cpu %.3e (par04 %.3e) launches %d clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["phase04_openmp"]["value"], d["gpu_launches"], d["clocks"]))
print({k[:3]: round(v["ms_per_launch"],2) for k,v in pk.items()}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), "fp64", round(d["roofline"]["fp64"]["frac"],3), "traffic", d["roofline"]["traffic"])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 400 gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
