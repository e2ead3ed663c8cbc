TAG=r02o
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_bench_config4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_config4.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("config4: %.2f ms/step " % d["ms_per_step"], {k[:2]: round(v["ms_per_launch"],2) for k,v in pk.items()}, "e2e", d["e2e"]["value"])
PY
for w in config2_viscous_gubser_bsq_2p1d config3_trento_pbpb_table_2p1d; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$w.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("$w: %.3f ms/step " % d["ms_per_step"], {k[:2]: round(v["ms_per_launch"],3) for k,v in pk.items()})
PY
done
timeout 900 ncu --set full --clock-control none -k regex:"eom_rk_kernel|eos_chain_stage_kernel" -s 6 -c 2 -o /tmp/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
python scripts/ncu_summary.py /tmp/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_eom_eos_summary.txt 2>/dev/null
grep -E "Kernel Name|gpu__time_duration|local_op|dram__bytes|l1tex__t_sector_hit|lts__t_sector_hit" gpurun_out/${TAG}_ncu_eom_eos_summary.txt | cut -c1-140
