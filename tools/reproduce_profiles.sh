#!/usr/bin/env bash
# Reproduces the evidence under profiles/ on a B200 box (run from the repository root, e.g. under `gpurun -- bash tools/reproduce_profiles.sh`).
# ncu steps follow /opt/skills/guides/B200_PROFILING.md: the same command first exits 0 without ncu, one GPU, --clock-control none.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json                                                   # profiles/r1_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json   # profiles/r1_bench_reference_arm.json
python tools/matbench.py > gpurun_out/matbench.jsonl                                          # profiles/r1_matbench.jsonl
python tools/autoshape.py > gpurun_out/bi_pass_by_n.jsonl                                     # profiles/r1_bi_pass_by_n.jsonl
python tools/configs_bench.py small > gpurun_out/tsplib_small.jsonl                           # profiles/r1_tsplib_small.jsonl
python tools/fi100k.py > gpurun_out/fi_large_n.jsonl                                          # profiles/r1_fi_large_n.jsonl
python tools/dropin_e2e.py > gpurun_out/dropin_e2e.jsonl                                      # profiles/r1_dropin_e2e.jsonl
python tools/fuzz.py 150 1 > gpurun_out/fuzz.jsonl                                            # profiles/r1_fuzz.jsonl
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tlo > gpurun_out/bench_short.json &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tlo > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_bench.csv > gpurun_out/launches_bench.txt    # profiles/r1_launches_bench.txt
python tools/prof.py 100000 6 matrix > gpurun_out/prof_plain.log &&
  ncu --set full --clock-control none --import-source on -k regex:'bi_scan_kernel|dist_matrix_kernel' -s 2 -c 5 -f -o gpurun_out/prof \
      python tools/prof.py 100000 6 matrix > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > gpurun_out/ncu_full.txt            # profiles/r1_ncu_full_*.txt
# multi-GPU (gpurun --gpus N): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
#     bench.py --gpus N --steps 100 --warmup 5            -> profiles/r1_bench_nN.json
#   ... tools/mgpu_check.py 20000 40 / tools/batch_mgpu.py FI  (parity of the sharded paths, profiles/r1_ga_batch_1_and_2_gpus.jsonl)
