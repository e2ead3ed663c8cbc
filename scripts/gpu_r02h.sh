TAG=r02h
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/${TAG}_pytest.log 2>&1
tail -12 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_bench_config4.err
tail -c 3000 gpurun_out/${TAG}_bench_config4.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_list.log
