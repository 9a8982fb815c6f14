# round-2 8-GPU session: parity of the sharded paths on 8 ranks, then the bench at N = 8 and N = 4 (each rank checks its tour
# against the committed single-GPU move log)
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tools/mgpu_check.py 20000 60 > gpurun_out/mgpu${N}_20k.log 2>&1; echo rc=$?
grep MGPU_CHECK gpurun_out/mgpu${N}_20k.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_g${N}.json 2> gpurun_out/bench_r2_g${N}.err; echo rc=$?
tail -2 gpurun_out/bench_r2_g${N}.err
if [ "$N" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/bench_r2_g4.json 2> gpurun_out/bench_r2_g4.err; echo rc=$?
fi
