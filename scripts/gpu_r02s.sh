TAG=r02s
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_bench_config4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_config4.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("config4: %.2f ms/step " % d["ms_per_step"], {k[:3]: round(v["ms_per_launch"],2) for k,v in pk.items()}, "e2e", d["e2e"]["value"], d["config"].get("eos_chain_members_rank0"))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4_n2.json 2> gpurun_out/${TAG}_bench_config4_n2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_config4_n2.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("n2: %.2f ms/step  %.3e p-steps/s" % (d["ms_per_step"], d["value"]), {k[:3]: round(v["ms_per_launch"],3) for k,v in pk.items()}, d["config"]["parallelism"][:22], d["per_rank"], "parity", d["parity_check"]["ok"], d["parity_check"]["worst_rel_err"], d["parity_check"].get("slab_axis"))
except Exception as e:
    print("n2 failed", e); print(open("gpurun_out/${TAG}_bench_config4_n2.err").read()[-1500:])
PY
