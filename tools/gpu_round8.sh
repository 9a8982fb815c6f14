set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_g${N}.json 2> gpurun_out/bench_g${N}.err; echo "rc=$?"
