# round-2 profiles (1 GPU): launch list of the bench command, ncu --set full of the scan kernel, the matrix kernel, the one-block
# BI kernel and the pruning path.  Every command first exits 0 without ncu (B200_PROFILING.md).
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tlo --no-extras > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tlo --no-extras > gpurun_out/r2_ncu_launches.log 2>&1
python tools/prof.py 100000 6 matrix > gpurun_out/r2_prof_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'bi_scan_kernel|dist_matrix_kernel' -s 2 -c 5 -f -o gpurun_out/r2_prof_scan \
      python tools/prof.py 100000 6 matrix > gpurun_out/r2_ncu_scan.log 2>&1
python tools/prof_batch.py 592 40 > gpurun_out/r2_prof_batch_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'two_opt_batch_bi_kernel|tile_boxes_kernel|tile_filter_kernel|bi_scan_kernel' -c 8 -f -o gpurun_out/r2_prof_batch \
      python tools/prof_batch.py 592 2 > gpurun_out/r2_ncu_batch.log 2>&1
python tools/prof.py 10000 60 > gpurun_out/r2_prof10k_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:bi_scan_kernel -s 20 -c 2 -f -o gpurun_out/r2_prof_10k python tools/prof.py 10000 60 > gpurun_out/r2_ncu_10k.log 2>&1
ls -la gpurun_out/*.ncu-rep
