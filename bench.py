#!/usr/bin/env python
"""bench.py — 2-opt move evaluations per second on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): synthetic uniform EUC_2D instance uni<n> (n = 100000, SURVEY.md §8(c) generator),
nearest-neighbour start tour built on the GPU, best-improvement 2-opt with on-the-fly distances.
One STEP = one best-improvement pass = all n(n-3)/2 move deltas evaluated + argmin + move applied.
  value   = passes * n(n-3)/2 / device time, tour resident in HBM (CUDA events on the engine's stream)
  e2e     = the same through the host-buffer C-ABI call tspb200_two_opt(): coordinates + tour uploaded from
            pinned host memory, K passes, tour downloaded, wall clock around the call
  N > 1   : the pair tiles are dealt round-robin over the ranks (strong scaling, same instance), one 8-byte
            NCCL min-allreduce per pass selects the move; max over ranks of the device time.
--impl reference times the reference's own CPU code (oracle/_ref: its calc_dist driven over rows of one
best-improvement scan, all host threads) on the same instance and start tour.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "2opt_move_evals_per_sec"
UNIT = "evals/s"
FP32_INSTR_PER_EVAL = 18  # per-unit figure of the on-the-fly 2-opt roofline


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tlo", action="store_true", help="skip the time-to-local-optimum runs (uni100000 on N GPUs, uni10000)")
    ap.add_argument("--rows-per-thread", type=int, default=0)
    ap.add_argument("--tile-cols", type=int, default=0)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed range
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


def pinned(arr: np.ndarray) -> np.ndarray:
    import torch
    t = torch.empty(arr.shape, dtype=getattr(torch, str(arr.dtype)), pin_memory=torch.cuda.is_available())
    out = t.numpy()
    out[...] = arr
    out_holder.append(t)
    return out


out_holder = []


def cpu_baseline_sample(xy, succ, seconds_budget: float):
    """Rows of ONE best-improvement scan on the host cores: the reference's compiled calc_dist when
    oracle/_ref is present (kind "reference"), else the oracle port."""
    from oracle.oracle import Oracle, RefLib, REF_SO
    n = len(xy)
    threads = os.cpu_count() or 1
    if os.path.exists(REF_SO):
        lib, kind = RefLib(), "reference"
    else:
        lib, kind = Oracle(), "port"
    # calibrate on a few rows, then size the sample for the time budget
    ev, sec, _ = lib.bi_scan_rows_mt(xy, 0, succ, 0, min(n - 1, 16 * threads), threads)
    rate = ev / max(sec, 1e-6)
    want = rate * seconds_budget
    total_pairs = n * (n - 3) // 2
    if want >= total_pairs:
        rows = n - 1
    else:  # rows r with r*n - r^2/2 ~= want
        rows = int(n - np.sqrt(max(0.0, float(n) * n - 2.0 * want)))
        rows = max(16 * threads, min(n - 1, rows))
    ev, sec, _ = lib.bi_scan_rows_mt(xy, 0, succ, 0, rows, threads)
    return {"value": ev / sec, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"rows [0,{rows}) of one best-improvement scan of uni{n} from the NN start = {ev} pair evaluations "
                      f"in {sec:.2f} s; {'reference calc_dist (oracle/_ref)' if kind == 'reference' else 'oracle port'}, "
                      f"{threads} threads, rows dealt in blocks of 16"}, rows


def run_reference(args, rank):
    """--impl reference: the reference CPU implementation on the same workload (rank 0 only)."""
    if rank != 0:
        return
    from oracle.oracle import Oracle
    from tsp_optimization_b200.instances import uniform_instance
    n = args.n
    xy = uniform_instance(n)
    succ = start_tour_cpu_or_cached(xy, n)
    total = max(1, args.steps + args.warmup)
    per_step = max(0.3, min(8.0, 150.0 / total))
    base, rows = cpu_baseline_sample(xy, succ, per_step)
    from oracle.oracle import RefLib, REF_SO
    lib = RefLib() if os.path.exists(REF_SO) else Oracle()
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        lib.bi_scan_rows_mt(xy, 0, succ, 0, rows, threads)
    ev_sum, sec_sum = 0, 0.0
    for _ in range(args.steps):
        ev, sec, _ = lib.bi_scan_rows_mt(xy, 0, succ, 0, rows, threads)
        ev_sum += ev
        sec_sum += sec
    val = ev_sum / sec_sum
    base["value"] = val
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec_sum / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"uni{n} EUC_2D, NN start, best-improvement 2-opt scan (reference CPU code)", "n": n},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def start_tour_cpu_or_cached(xy, n):
    """NN start tour without a GPU (reference arm).  greedy() is O(n^2) on the CPU (10^10 calc_dist calls at n = 100 000),
    so the committed fixture tests/golden/nn_uni<n>.npz (generated by tests/golden/make_nn_uni100000.py with the oracle; the
    GPU nearest-neighbour kernel is tested against it) is used when it exists, then a cache from an earlier run."""
    from oracle.oracle import Oracle
    fixture = os.path.join(ROOT, "tests", "golden", f"nn_uni{n}.npz")
    if os.path.exists(fixture):
        return np.load(fixture)["succ"].astype(np.int32)
    cache = os.path.join(ROOT, "gpurun_out", f"nn_uni{n}.npy")
    if os.path.exists(cache):
        return np.load(cache)
    succ, _ = Oracle().nn_tour(xy, 0, 0)
    try:
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.save(cache, succ)
    except OSError:
        pass
    return succ


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION in this image) off it
    # (every level from VERSION up prints it, so the variable is dropped rather than raised to WARN)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        del os.environ["NCCL_DEBUG"]
    import torch
    import torch.distributed as dist

    from tsp_optimization_b200 import BI, Engine
    from tsp_optimization_b200.dist import attach_engine_comm, init_process_group_from_env
    from tsp_optimization_b200.instances import uniform_instance

    if world > 1:
        init_process_group_from_env("nccl")
    torch.cuda.set_device(local)
    n = args.n
    xy = uniform_instance(n)
    pairs = n * (n - 3) // 2

    eng = Engine(local)
    if args.rows_per_thread:
        eng.set_option("rows_per_thread", args.rows_per_thread)
    if args.tile_cols:
        eng.set_option("tile_cols", args.tile_cols)
    eng.set_instance(xy, 0)
    t0 = time.perf_counter()
    succ0, nn_cost = eng.nn_tour(0)  # identical on every rank (deterministic)
    nn_s = time.perf_counter() - t0
    if rank == 0:
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.save(os.path.join(ROOT, "gpurun_out", f"nn_uni{n}.npy"), succ0)
        except OSError:
            pass
    if world > 1:
        attach_engine_comm(eng, rank, world)
    eng.tour_upload(succ0)

    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # matrix kernel timing only

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident-tour timing: W warm-up passes, then exactly K timed passes in ONE engine call --------------------
    # l2_flush_bytes: the engine writes 256 MB (> 126 MB L2) before every pass and times each pass with its own CUDA
    # event pair on its stream; value = K * pairs / sum of the per-pass intervals (max over ranks).  No host round trip
    # between passes, so multi-GPU ranks stay in step through the device-side exchange only.
    eng.bi_run(args.warmup)
    if not args.no_flush:
        eng.set_option("l2_flush_bytes", 256 << 20)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    st = eng.bi_run(args.steps)
    barrier()
    wall_s = time.perf_counter() - wall0
    clocks = sampler.stop()
    eng.set_option("l2_flush_bytes", 0)
    gpu_ms, launches, moves, done_passes = st.gpu_ms, st.launches, st.moves, st.passes
    if world > 1:
        t = torch.tensor([gpu_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gpu_ms = float(t.item())
    value = done_passes * pairs / (gpu_ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region -------------------------------
    h_xy = pinned(xy)
    h_succ = pinned(succ0.astype(np.int32))
    e2e_passes = max(1, args.steps)
    eng.set_instance(h_xy, 0)           # warm-up of the same call sequence
    eng.two_opt(BI, h_succ, 0.0, max_iters=2)
    barrier()
    w0 = time.perf_counter()
    eng.set_instance(h_xy, 0)
    s_out, obj_out, st_e, _ = eng.two_opt(BI, h_succ, 0.0, max_iters=e2e_passes)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = st_e.passes * pairs / e2e_s
    h2d = 16 * n + 4 * n
    d2h = 4 * n + 8

    # ---- e2e, strictest reading: EVERY step is its own public call with its own host<->device copies -----------------
    # K calls of tspb200_two_opt(max_iters = 1): each uploads the step's input tour from pinned host memory (4n bytes),
    # runs one pass, downloads the resulting tour and its cost (4n + 8 bytes), which is the next step's input.  The
    # instance (coordinates) stays resident like any constant of the job.
    step_calls = max(1, min(args.steps, 50))
    cur = h_succ.copy()
    eng.two_opt(BI, cur, 0.0, max_iters=1)
    barrier()
    w0 = time.perf_counter()
    done_step = 0
    for _ in range(step_calls):
        cur, _, st_s, _ = eng.two_opt(BI, cur, 0.0, max_iters=1)
        done_step += st_s.passes
    torch.cuda.synchronize()
    step_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([step_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s = float(t.item())
    e2e_step = {"value": done_step * pairs / step_s, "unit": UNIT, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4 * n + 8,
                "steps": step_calls, "seconds": step_s, "call": "tspb200_two_opt(BI, host succ[], max_iters=1) once per step"}

    # ---- time to local optimum (BASELINE metric, second half): the same instance and start tour, run to the end ----
    tlo = None
    if not args.no_tlo:
        eng.set_instance(h_xy, 0)
        barrier()
        w0 = time.perf_counter()
        s_out, obj_out, st_f, _ = eng.two_opt(BI, h_succ, 0.0)
        torch.cuda.synchronize()
        tlo_s = time.perf_counter() - w0
        if world > 1:
            t = torch.tensor([tlo_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tlo_s = float(t.item())
        tlo = {"workload": f"uni{n} NN start -> best-improvement 2-opt local optimum on {world} GPU(s), host buffers in and out",
               "time_to_local_optimum_s": tlo_s, "passes": st_f.passes, "moves": st_f.moves, "gpu_ms": st_f.gpu_ms,
               "start_cost": nn_cost, "final_cost": obj_out, "evals": st_f.evals, "evals_per_s": st_f.evals / tlo_s}
        if world == 1 and n != 10000:  # BASELINE configs[2]: uni10000, greedy start + 2-opt to the local optimum, 1 B200
            xy2 = uniform_instance(10000)
            eng.set_instance(xy2, 0)
            s2, c2 = eng.nn_tour(0)
            eng.two_opt(BI, s2, 0.0, max_iters=4)  # warm-up of this tile shape
            for mode, nm in ((BI, "BI"), (1 - BI, "FI")):
                w0 = time.perf_counter()
                _, o2, st2, _ = eng.two_opt(mode, s2, c2)
                d2 = time.perf_counter() - w0
                tlo[f"uni10000_{nm}"] = {"time_to_local_optimum_s": d2, "passes": st2.passes, "moves": st2.moves,
                                         "final_cost": o2, "evals": st2.evals}
            eng.set_instance(h_xy, 0)
            eng.tour_upload(h_succ)
    if rank != 0:
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peaks_src = measured_peaks()
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    num_sms = eng.info("num_sms")
    per_gpu = value / world
    # Binding unit of bi_scan_kernel: the SFU (MUFU.SQRT, 16 results/clk/SM).  Algorithmic minimum = ONE distance
    # (one sqrt) per evaluated move: every D[p][q] is shared by the two moves that use it.
    sqrt_peak = num_sms * 16 * sm_mhz * 1e6
    fp32_peak = num_sms * 128 * sm_mhz * 1e6
    R = eng.info("rows_per_thread")
    roofline = {"bound": "sfu_sqrt", "achieved": per_gpu / 1e12, "peak": sqrt_peak / 1e12, "unit": "T sqrt/s (= T evals/s)",
                "frac": per_gpu / sqrt_peak, "traffic": 1.726e6, "kernel": "bi_scan_kernel",
                "per_unit": f"1 MUFU.SQRT per evaluated move (algorithmic minimum: every distance serves two moves); the kernel "
                            f"issues (R+1)/R = {(R + 1) / R:.4f} with R = {R} rows per thread",
                "peak_source": f"{num_sms} SMs x 16 MUFU/clk x {sm_mhz:.0f} MHz (nvidia-smi median under load); "
                               f"MEASURED_PEAKS.json ({peaks_src}) holds HBM and bf16 figures only, neither bounds this kernel",
                "ncu": "profiles/r1_ncu_full_bi_scan_64x8_final.txt (bi_scan_kernel<64,8>, n = 100 000): XU pipe 89 % of peak, "
                       "issue slots 66 %, 6.66 thread instructions per evaluated move, DRAM 1.7 MB read / 0 written per launch",
                "fp32_issue_view": {"survey_per_unit": FP32_INSTR_PER_EVAL,
                                    "frac_vs_survey_18_instr_model": per_gpu * FP32_INSTR_PER_EVAL / fp32_peak,
                                    "executed_thread_instr_per_eval": 6.66,
                                    "frac_issue_slots": per_gpu * 6.66 / fp32_peak,
                                    "note": "SURVEY.md §8d's 18 lane-instr/eval model evaluates two fresh distances per move with "
                                            "scalar FP32; sharing each distance between its two moves and packed FP32x2 arithmetic "
                                            "bring the executed count to 6.66, so the 18-instr fraction exceeds 1 and is reported "
                                            "for reference only"},
                "traffic_note": "dram__bytes_read.sum per launch (ncu --set full, cold L2): the tour records once, 16 B/node; compute-bound"}
    # secondary kernel: distance matrix, HBM-store-bound (4*n*ld bytes written per launch)
    mat = None
    if world == 1:
        nm = 20000  # 1.6 GB of int32 (> L2); the size the ncu --set full capture in profiles/ was taken at
        eng_m = eng
        eng_m.set_instance(uniform_instance(nm), 0)
        eng_m.dist_matrix_build()
        ms_list = []
        for _ in range(5):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()
            ms_list.append(eng_m.dist_matrix_build())
        ld = eng_m.info("matrix_ld")
        eng_m.dist_matrix_free()
        ms = float(np.median(ms_list))
        gbs = 4.0 * nm * ld / (ms * 1e-3) / 1e9
        mat = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
               "traffic": 1.5446e9, "traffic_note": "ncu --set full at this n: dram__bytes_write.sum 1.544 GB + dram__bytes_read.sum 0.3 MB per launch, of "
                                                    "1.600 GB algorithmic (the rest is still dirty in L2 at kernel end); gpu__dram_throughput 82 % "
                                                    "of peak (profiles/r1_ncu_full_bi_scan_128x16_and_matrix.txt)",
               "kernel": "dist_matrix_kernel", "n": nm, "ms": ms,
               "per_unit": "4 bytes written per matrix entry (int32), reads O(n)", "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks_src})"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": gpu_ms / max(1, done_passes), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"uni{n} EUC_2D (SURVEY.md §8c generator), GPU nearest-neighbour start, best-improvement "
                                   f"2-opt passes with on-the-fly distances (BASELINE configs[3])",
                       "n": n, "pairs_per_step": pairs, "block_threads": eng.info("block_threads"), "rows_per_thread": eng.info("rows_per_thread"),
                       "tile_cols": eng.info("tile_cols"), "grid": eng.info("grid_bi"), "tiles": eng.info("ntiles"),
                       "sharding": ("tiles round-robin over ranks; per pass each rank's 8-byte argmin key is " +
                                    ("stored into every peer's slots over NVLink by the scan kernel (CUDA IPC peer memory)"
                                     if eng.info("exchange_p2p") else "min-allreduced by NCCL")) if world > 1 else "single GPU",
                       "l2": "flushed before every timed pass (256 MB write on the engine stream, outside the per-pass event pairs)"
                             if not args.no_flush else "not flushed (5.7 MB working set)",
                       "nn_start_s": nn_s},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_passes, "d2h_bytes_per_step": d2h / e2e_passes,
                    "passes": st_e.passes, "seconds": e2e_s,
                    "call": "tspb200_set_instance + tspb200_two_opt(BI, host succ[], max_iters=steps)",
                    "one_call_per_step": e2e_step},
            "gpu_launches": int(launches), "moves_applied": int(moves), "wall_s": wall_s,
            "clocks": clocks, "roofline": roofline}
    if mat:
        line["roofline_matrix"] = mat
    if tlo:
        line["time_to_local_optimum"] = tlo
    if not args.no_cpu_baseline and world == 1:
        base, _ = cpu_baseline_sample(xy, succ0, 15.0)
        line["cpu_baseline"] = base
    print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
