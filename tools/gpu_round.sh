timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 900 python tools/fi100k.py > gpurun_out/fi100k.jsonl 2>&1
timeout 600 python tools/configs_bench.py c2 > gpurun_out/configs_c2.jsonl 2>&1
