set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 20000 40 > gpurun_out/mgpu${N}.log 2>&1; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_g${N}.json 2> gpurun_out/bench_g${N}.err; echo "rc=$?"
