TAG=r02g
mkdir -p gpurun_out
(time timeout 3000 python -m pytest tests/test_gpu_freeze_out.py tests/test_gpu_preinverted.py -m gpu -q 2>&1 | tail -70) > gpurun_out/${TAG}_pytest_new.log 2>&1
tail -40 gpurun_out/${TAG}_pytest_new.log
(time timeout 3000 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_freeze_out.py --deselect tests/test_gpu_preinverted.py 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
