"""Dev tool: turn a .ncu-rep (ncu --set full) and a launch list (ncu --metrics gpu__time_duration.sum --csv) into the
small text summaries committed under profiles/.  Runs here (no GPU): `ncu -i <rep> --page raw --csv`."""
import csv
import subprocess
import sys
from collections import OrderedDict

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}  (launch id {r[0]})")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:85s} {r[i]:>18s} {units[i]}")


def launches(path):
    rows = list(csv.reader(open(path)))
    h0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[h0]
    d = OrderedDict()
    for r in rows[h0 + 1:]:
        if len(r) < len(h):
            continue
        rec = dict(zip(h, r))
        d.setdefault((rec["Kernel Name"], rec["Grid Size"], rec["Block Size"]), []).append(float(rec["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print(f"{'kernel':90s} {'grid':>16s} {'block':>14s} {'launches':>8s} {'mean_us':>12s} {'sum_us':>14s} {'share':>8s}")
    for (k, g, b), v in d.items():
        print(f"{k[:90]:90s} {g:>16s} {b:>14s} {len(v):8d} {sum(v)/len(v)/1e3:12.2f} {sum(v)/1e3:14.2f} {sum(v)/tot:8.4f}")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2])
    else:
        launches(sys.argv[2])
