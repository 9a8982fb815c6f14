#!/bin/bash
# first GPU trip: each group in its own process with a hard timeout so that one hang cannot hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -k "$@" > gpurun_out/t_$name.log 2>&1; echo "exit $?"; tail -5 gpurun_out/t_$name.log; }
run matrix "matrix"
run bi "bi_ or berlin52"
run fi "fi_"
run batch "batch"
run nn "nn_ or tour_costs"
run misc "dropin or errors or duplicate or non_fp32 or random_start"
run big "uni4000 or uni10000"
echo "=== bench small"
timeout -s KILL 300 python bench.py --n 20000 --steps 20 --warmup 3 > gpurun_out/bench_20k.json 2> gpurun_out/bench_20k.err; echo "exit $?"; cat gpurun_out/bench_20k.json; tail -3 gpurun_out/bench_20k.err
