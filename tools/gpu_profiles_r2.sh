# round-2 profiles (1 GPU): launch list of the bench command, ncu --set full of the scan kernel, the matrix kernel, the one-block
# BI kernel and the pruning path.  Every command first exits 0 without ncu (B200_PROFILING.md).  The .ncu-rep files are turned
# into text summaries on the box and deleted: gpurun_out/ must stay under 64 MiB to travel back.
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tlo --no-extras > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tlo --no-extras > gpurun_out/r2_ncu_launches.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench.txt
python tools/prof.py 100000 6 matrix exhaustive > gpurun_out/r2_prof_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'bi_scan_kernel|dist_matrix_kernel' -s 3 -c 4 -f -o gpurun_out/r2_prof_scan \
      python tools/prof.py 100000 6 matrix exhaustive > gpurun_out/r2_ncu_scan.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_prof_scan.ncu-rep > gpurun_out/r2_ncu_full_scan_and_matrix.txt
python tools/ncu_hotspots.py gpurun_out/r2_prof_scan.ncu-rep bi_scan_kernel 40 > gpurun_out/r2_scan_hotspots.txt 2>&1
rm -f gpurun_out/r2_prof_scan.ncu-rep
python tools/prof_batch.py 592 40 > gpurun_out/r2_prof_batch_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'two_opt_batch_bi_kernel|tile_boxes_kernel|tile_filter_kernel|bi_scan_kernel|apply_move' -c 9 -f -o gpurun_out/r2_prof_batch \
      python tools/prof_batch.py 592 2 > gpurun_out/r2_ncu_batch.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_prof_batch.ncu-rep > gpurun_out/r2_ncu_full_batch_bi_and_pruning.txt
python tools/ncu_hotspots.py gpurun_out/r2_prof_batch.ncu-rep two_opt_batch_bi_kernel 40 > gpurun_out/r2_batch_bi_hotspots.txt 2>&1
rm -f gpurun_out/r2_prof_batch.ncu-rep
python tools/prof.py 10000 60 > gpurun_out/r2_prof10k_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:bi_scan_kernel -s 20 -c 2 -f -o gpurun_out/r2_prof_10k python tools/prof.py 10000 60 > gpurun_out/r2_ncu_10k.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_prof_10k.ncu-rep > gpurun_out/r2_ncu_full_scan_n10k_pruned.txt
python tools/ncu_hotspots.py gpurun_out/r2_prof_10k.ncu-rep bi_scan_kernel 60 > gpurun_out/r2_scan_n10k_hotspots.txt 2>&1
rm -f gpurun_out/r2_prof_10k.ncu-rep
du -sh gpurun_out
