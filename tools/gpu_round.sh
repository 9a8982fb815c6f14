timeout 900 python -m pytest tests/test_gpu_dropin_link.py -m gpu -x -q > gpurun_out/pytest_link.log 2>&1
