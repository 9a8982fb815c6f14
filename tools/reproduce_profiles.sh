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

# ---- round 2 (files profiles/r2_*) ----------------------------------------------------------------------------------------
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_g1.json                            # profiles/r2_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json           # profiles/r2_bench_reference_arm.json
bash tools/gpu_profiles_r2.sh                                                                  # r2_launches_bench.txt, r2_ncu_full_*.txt, r2_*_hotspots.txt
python tools/prune_probe.py 10000 20000 > gpurun_out/prune_probe.jsonl                         # r2_prune_probe_n10k_n20k.jsonl (PRUNE_SHAPES=... for r2_prune_probe_small_tiles.jsonl)
python tools/prune_probe.py 100000 > gpurun_out/prune_probe_100k.jsonl                         # r2_prune_probe_n100k.jsonl
python tools/r2_probe.py b c d > gpurun_out/r2_probe.jsonl                                     # r2_fi_latency.jsonl, r2_pruned_pass_phases_n10k.jsonl, r2_nn_grid_walk.jsonl
python tools/shuf_ab.py > gpurun_out/shuf_ab.jsonl                                             # r2_row_shuffle_ab.jsonl
python tools/occ_ab.py > gpurun_out/occ_ab.jsonl                                               # r2_blocks_per_sm_ab.jsonl (second build: make OUT=../lib_alt EXTRA=-DTSPB_BI_MINBLOCKS64=8, TSPB200_LIB=...)
for w in 8 4 2; do python tools/shardshape.py $w; done > gpurun_out/shardshape.jsonl           # r2_shard_shapes_row_shuffle.jsonl
python tools/tail_probe.py > gpurun_out/tail_probe.jsonl; python tools/tj_sweep.py > gpurun_out/tj_sweep.jsonl   # r2_tail_probe.jsonl, r2_tile_width_sweep.jsonl
python tools/fuzz.py 150 21 > gpurun_out/fuzz.jsonl                                            # r2_fuzz.jsonl (one line per run)
# multi-GPU (gpurun --gpus 2 / 8): bash tools/gpu_round2.sh, bash tools/gpu_round8.sh         -> r2_bench_n2/4/8.json, r2_mgpu_check.txt
