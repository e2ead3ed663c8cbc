# One-GPU evidence run for profiles/: default bench line, ncu launch list of the same command, ncu --set full of one launch
# of each of the heaviest kernels. TAG names the round (e.g. r01c).
TAG=${1:-r01x}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_bench_config2.err; tail -c 600 gpurun_out/${TAG}_bench_config2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_config2_reference.json 2> /dev/null; tail -c 300 gpurun_out/${TAG}_bench_config2_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"eom_rk_kernel|gather_kernel|eos_local_kernel" -s 8 -c 4 -o gpurun_out/${TAG}_prof_top_config2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_top.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_top.log
