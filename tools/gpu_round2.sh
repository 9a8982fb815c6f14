# round-2 multi-GPU session (2 GPUs): parity of the sharded paths, bench N=2 with the per-pass breakdown
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/t_r2_mgpu.log 2>&1; echo rc=$?
tail -15 gpurun_out/t_r2_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_check.py 20000 100 > gpurun_out/mgpu2_20k.log 2>&1; echo rc=$?
grep MGPU_CHECK gpurun_out/mgpu2_20k.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_g2.json 2> gpurun_out/bench_r2_g2.err; echo rc=$?
tail -3 gpurun_out/bench_r2_g2.err
