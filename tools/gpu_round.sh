set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "matrix" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 600 python tools/matbench.py > gpurun_out/matbench.jsonl 2>&1; echo "rc=$?"
