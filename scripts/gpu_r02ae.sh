TAG=r02ae
mkdir -p gpurun_out
np=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus $np --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_config4_n8.json 2> gpurun_out/${TAG}_bench_config4_n8.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_config4_n8.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("n8: %.2f ms/step  %.3e p-steps/s e2e %.3e" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k[:3]: round(v["ms_per_launch"],2) for k,v in pk.items()}, d["config"]["parallelism"][:22], "parity", d["parity_check"]["ok"], d["parity_check"]["worst_rel_err"], "cpu", d["cpu_baseline"])
    for r,x in enumerate(d["per_rank"]): print("   rank", r, x["particles"], "compute %.2f  exchange+wait %.2f" % (x["compute_ms_per_step"], x["slab_exchange_ms_per_step"]))
except Exception as e:
    print("n8 failed", e); print(open("gpurun_out/${TAG}_bench_config4_n8.err").read()[-1500:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29589 bench.py --impl reference --gpus $np --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n8.json 2> gpurun_out/${TAG}_bench_ref_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_ref_n8.json').read().strip().splitlines()[-1]); print('reference arm N=8:', d['value'], 'cores', d['cpu_baseline']['cores'], d['n_gpus'])"
