# round 2, first GPU run: parity suite, default bench (config 4, 8M particles) at both register budgets, ncu captures at full size
TAG=r02a
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
for R in 0 1; do
  CCB_GATHER_REGS=$R timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4_regs$R.json 2> gpurun_out/${TAG}_bench_config4_regs$R.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_config4_regs$R.json").read().strip().splitlines()[-1])
    print("regs=$R config4: %.3e p-steps/s  %.2f ms/step  e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: v["ms_per_launch"] for k, v in d["roofline"]["per_kernel"].items()})
except Exception as e:
    print("bench regs=$R failed", e); print(open("gpurun_out/${TAG}_bench_config4_regs$R.err").read()[-1500:])
PY
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|eom_rk_kernel|eos_chain_stage_kernel|eos_local_kernel|permute_columns|radix_scatter|prologue_kernel" -s 40 -c 12 -o gpurun_out/${TAG}_prof_config4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
