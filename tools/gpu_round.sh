timeout 900 python -m pytest tests -m gpu -x -q -k "tiny or caps or time_limit or large_coord" > gpurun_out/pytest_gpu.log 2>&1
