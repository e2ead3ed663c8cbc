TAG=r02f
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests/test_gpu_preinverted.py -m gpu -q -x 2>&1 | tail -60) > gpurun_out/${TAG}_pytest.log 2>&1
tail -30 gpurun_out/${TAG}_pytest.log
