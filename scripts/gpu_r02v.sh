TAG=r02v
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_dropin.py tests/test_slabs.py tests/test_multi_gpu.py tests/test_gpu_preinverted.py -m gpu -q 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
