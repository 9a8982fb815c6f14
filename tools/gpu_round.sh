timeout 900 python -m pytest tests -m gpu -x -q -k "nn or tiny or dropin or smoke" > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python tools/configs_bench.py nn > gpurun_out/configs_nn.jsonl 2>&1
