timeout 1500 python tools/dropin_e2e.py > gpurun_out/dropin_e2e.jsonl 2>&1
