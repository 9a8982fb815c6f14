timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python tools/shardtime.py > gpurun_out/shardtime.jsonl 2>&1
timeout 600 python tools/autoshape.py > gpurun_out/autoshape.jsonl 2>&1
