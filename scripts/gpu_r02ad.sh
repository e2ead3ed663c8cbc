TAG=r02ad
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke") | tail -4
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("default: %.2f ms/step %.3e  e2e %.3e  cpu %.3e (par04 %.3e) launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["phase04_openmp"]["value"], d["gpu_launches"]))
print({k[:3]: round(v["ms_per_launch"],2) for k,v in pk.items()}, "fp64", round(d["roofline"]["fp64"]["frac"],3), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm:', d['value'], d['cpu_baseline']['cores'], d['config']['workload'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4_n2.json 2> gpurun_out/${TAG}_bench_config4_n2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_config4_n2.json").read().strip().splitlines()[-1])
    print("n2: %.2f ms/step  %.3e p-steps/s" % (d["ms_per_step"], d["value"]), d["config"]["parallelism"][:22], "parity", d["parity_check"]["ok"], d["parity_check"]["worst_rel_err"])
except Exception as e:
    print("n2 failed", e); print(open("gpurun_out/${TAG}_bench_config4_n2.err").read()[-1500:])
PY
