# ncu --set full capture of the three heaviest kernels of one bench step (1 GPU; never a multi-rank command)
mkdir -p gpurun_out
TAG=${1:-a}
ncu --set full --clock-control none --import-source on -k regex:"eom_rk_kernel|gather_kernel" -s 6 -c 3 \
    -o gpurun_out/prof_top_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_top_$TAG.log 2>&1
tail -3 gpurun_out/ncu_top_$TAG.log
