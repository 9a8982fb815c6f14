set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "tabu or dropin or masked" > gpurun_out/pytest_tabu.log 2>&1; echo "pytest rc=$?"
