TAG=r02w
mkdir -p gpurun_out
run() { # name nproc env...
  name=$1; np=$2; shift; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 295$np$np bench.py --gpus $np --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4_$name.json 2> gpurun_out/${TAG}_bench_config4_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_config4_$name.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("$name: %.2f ms/step  %.3e p-steps/s e2e %.3e" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k[:3]: round(v["ms_per_launch"],2) for k,v in pk.items()}, d["config"]["parallelism"][:22], "parity", d["parity_check"]["ok"], d["parity_check"]["worst_rel_err"], d["parity_check"]["slab_axis"])
    for r,x in enumerate(d["per_rank"]): print("   rank", r, x["particles"], "compute %.2f  exchange+wait %.2f" % (x["compute_ms_per_step"], x["slab_exchange_ms_per_step"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_bench_config4_$name.err").read()[-1500:])
PY
}
run n8_strong 8 A=1
run n4_strong 4 A=1
