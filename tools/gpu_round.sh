python tools/batch_mgpu.py FI > gpurun_out/batch_g1.jsonl 2>&1
python tools/batch_mgpu.py BI >> gpurun_out/batch_g1.jsonl 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29521 tools/batch_mgpu.py FI > gpurun_out/batch_g2.jsonl 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29522 tools/batch_mgpu.py BI >> gpurun_out/batch_g2.jsonl 2>&1
