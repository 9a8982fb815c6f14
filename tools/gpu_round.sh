timeout 900 python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"
