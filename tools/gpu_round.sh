timeout 900 python -m pytest tests -m gpu -x -q -k "nn_batch or greedy_iter" > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python tools/configs_bench.py nnb > gpurun_out/nnb.jsonl 2>&1
