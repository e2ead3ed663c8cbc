TAG=r02af
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"FieldsPolicy|eos_table_fast_kernel|eos_chain_stage_kernel|GradPolicy" -s 12 -c 5 -o /tmp/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-160
python scripts/ncu_summary.py /tmp/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_final_kernels_config4.txt 2>/dev/null
grep -E "Kernel Name|gpu__time_duration|registers_per_thread|warps_active|issue_active|fp64_cycles|local_op|lts__t_sector_hit" gpurun_out/${TAG}_ncu_final_kernels_config4.txt | cut -c1-150
