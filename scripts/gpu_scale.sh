# bench.py at N GPUs of one box (weak scaling of the default workload; optionally a second workload / strong scaling)
N=${1:-2}; TAG=${2:-x}; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/${TAG}_bench_config2_n${N}_weak.json 2> gpurun_out/${TAG}_n${N}.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/${TAG}_bench_config2_n${N}_weak.json").read().strip().splitlines()[-1])
print("N=$N weak config2: %.3e p-steps/s  %.3f ms/step  e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: v["ms_per_launch"] for k, v in d["roofline"]["per_kernel"].items()})
EOF
if [ -n "$1" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 --workload config4_milne_bsq_table_3p1d --scaling strong > gpurun_out/${TAG}_bench_config4_n${N}_strong.json 2> gpurun_out/${TAG}_n${N}_c4.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/${TAG}_bench_config4_n${N}_strong.json").read().strip().splitlines()[-1])
print("N=$N strong config4: %.3e p-steps/s  %.3f ms/step  e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: v["ms_per_launch"] for k, v in d["roofline"]["per_kernel"].items()})
EOF
fi
