TAG=r02n
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_bench_config4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|eom_rk_kernel|eos_chain_stage_kernel|eos_local_kernel|permute_columns|prologue_kernel|radix_scatter" -s 40 -c 14 -o /tmp/${TAG}_prof_stage_config4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
python scripts/ncu_summary.py /tmp/${TAG}_prof_stage_config4.ncu-rep --traffic > gpurun_out/${TAG}_ncu_stage_summary_config4.txt 2> gpurun_out/${TAG}_traffic_config4.json
grep -c "Kernel Name" gpurun_out/${TAG}_ncu_stage_summary_config4.txt
for k in 0 1 2 3 4 5 6 7 8 9 10 11 12 13; do python scripts/ncu_regions.py /tmp/${TAG}_prof_stage_config4.ncu-rep $k 1.0 > gpurun_out/${TAG}_regions_$k.txt 2>&1; done
ls -la gpurun_out/${TAG}_* /tmp/${TAG}_prof_stage_config4.ncu-rep
