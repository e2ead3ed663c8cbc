TAG=r02d
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_kernel" -s 14 -c 2 -o gpurun_out/${TAG}_prof_gather_config4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/
