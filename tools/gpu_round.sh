timeout 900 python -m pytest tests -m gpu -x -q -k "extra_mileage or csv_all" > gpurun_out/pytest_gpu.log 2>&1
