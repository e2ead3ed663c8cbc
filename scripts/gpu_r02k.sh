TAG=r02k
mkdir -p gpurun_out
nvidia-smi -L
(timeout 1500 python -m pytest tests/test_gpu_gubser_check.py tests/test_multi_gpu.py tests/test_slabs.py -m gpu -q -s 2>&1 | tail -25) > gpurun_out/${TAG}_pytest.log 2>&1
tail -12 gpurun_out/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config4_n2.json 2> gpurun_out/${TAG}_bench_config4_n2.err
tail -c 1500 gpurun_out/${TAG}_bench_config4_n2.json; tail -5 gpurun_out/${TAG}_bench_config4_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err
tail -c 600 gpurun_out/${TAG}_bench_ref_n2.json
