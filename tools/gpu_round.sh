timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python tools/configs_bench.py small,c4 > gpurun_out/configs_small.jsonl 2>&1
