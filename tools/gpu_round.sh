set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 6000 60 > gpurun_out/mgpu2.log 2>&1; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py 100000 40 > gpurun_out/mgpu2_100k.log 2>&1; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "rc=$?"
