set -x
timeout 900 python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "rc=$?"
