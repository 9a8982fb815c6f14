"""Round-2 latency probes (dev tool): (a) per-block time line of a pruned pass at n = 10 000 / 20 000, (b) first-improvement
latency per move as a function of the gap between moves at n = 100 000.  python tools/r2_probe.py [a] [b]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from tsp_optimization_b200 import BI, FI, Engine  # noqa: E402
from tsp_optimization_b200.instances import uniform_instance  # noqa: E402

TM = ["tm_gap", "tm_scan", "tm_spread", "tm_tail", "tm_xwait", "tm_apply_gap", "tm_apply"]


def probe_a(eng):
    for n in (10000, 20000):
        eng.set_instance(uniform_instance(n), 0)
        succ, _ = eng.nn_tour(0)
        for prune in (1, 0):
            eng.set_option("prune", prune)
            for skip in (20, 400):
                eng.set_option("timing", 1)
                eng.tour_upload(succ)
                eng.bi_run(skip)
                eng.set_option("timing", 2)
                st = eng.bi_run(1)
                bt = eng.block_times().astype(np.int64)
                ok = bt[:, 0] > 0
                t0 = bt[ok, 0].min()
                start, end = (bt[ok, 0] - t0) / 1e3, (bt[ok, 1] - t0) / 1e3
                dur = end - start
                order = np.argsort(end)
                ids = np.nonzero(ok)[0]
                print(json.dumps({"probe": "a", "n": n, "prune": prune, "after_passes": skip, "blocks": int(ok.sum()),
                                  "shape": [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")],
                                  "tiles_scanned": st.tiles_scanned, "tiles_total": st.tiles_total,
                                  "start_us_pct": np.percentile(start, [0, 50, 100]).round(2).tolist(),
                                  "end_us_pct": np.percentile(end, [0, 10, 25, 50, 75, 90, 99, 100]).round(2).tolist(),
                                  "dur_us_pct": np.percentile(dur, [0, 10, 25, 50, 75, 90, 99, 100]).round(2).tolist(),
                                  "busy_blocks_gt2us": int((dur > 2.0).sum()),
                                  "slowest_blocks": ids[order[-8:]].tolist(), "slowest_end_us": end[order[-8:]].round(2).tolist(),
                                  "slowest_dur_us": dur[order[-8:]].round(2).tolist()}), flush=True)
        eng.set_option("timing", 0)
        eng.set_option("prune", -1)


def probe_b(eng):
    n = 100000
    eng.set_instance(uniform_instance(n), 0)
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "nn_uni100000.npz"))
    succ = z["succ"]
    eng.tour_upload(succ)
    eng.fi_run(50)
    eng.tour_upload(succ)
    tot = 0
    while True:
        st = eng.fi_run(1000)
        tot += st.moves
        print(json.dumps({"probe": "b", "moves_so_far": tot, "moves": st.moves, "gpu_ms": round(st.gpu_ms, 3), "sweeps_done": st.passes,
                          "pairs_swept": st.evals, "launches": st.launches,
                          "us_per_move": round(1e3 * st.gpu_ms / max(1, st.moves), 2),
                          "pairs_per_move": round(st.evals / max(1, st.moves))}), flush=True)
        if st.moves < 1000:
            break


def probe_c(eng):
    """phase stamps of a pruned pass (option timing = 2): where does a block with ONE small tile spend its microseconds?"""
    import ctypes as C
    for n in (10000,):
        eng.set_instance(uniform_instance(n), 0)
        succ, _ = eng.nn_tour(0)
        eng.set_option("prune", 1)
        for skip in (20, 400):
            eng.set_option("timing", 1)
            eng.tour_upload(succ)
            eng.bi_run(skip)
            eng.set_option("timing", 2)
            eng.bi_run(1)
            out = np.zeros((4096, 8), dtype=np.uint64)
            eng._ck(eng.L.tspb200_debug_fetch(eng.h, b"block_phases", out.ctypes.data, out.nbytes))
            ph = out[:eng.info("grid_bi")].astype(np.int64)
            busy = ph[:, 5] > 0
            t0 = ph[:, 0].min()
            def pct(a):
                return np.percentile(a, [0, 25, 50, 75, 90, 100]).round(2).tolist() if len(a) else []
            b = ph[busy]
            e = ph[~busy]
            print(json.dumps({"probe": "c", "n": n, "after_passes": skip, "busy_blocks": int(busy.sum()), "idle_blocks": int((~busy).sum()),
                              "busy_start_us": pct((b[:, 0] - t0) / 1e3), "busy_start_to_drawn_us": pct((b[:, 1] - b[:, 0]) / 1e3),
                              "busy_drawn_to_loaded_us": pct((b[:, 2] - b[:, 1]) / 1e3), "busy_loaded_to_scanned_us": pct((b[:, 3] - b[:, 2]) / 1e3),
                              "busy_scanned_to_folded_us": pct((b[:, 4] - b[:, 3]) / 1e3), "busy_folded_to_ticket_us": pct((b[:, 6] - b[:, 4]) / 1e3),
                              "busy_tiles": pct(b[:, 5]), "busy_total_us": pct((b[:, 6] - b[:, 0]) / 1e3),
                              "idle_start_to_drawn_us": pct((e[:, 1] - e[:, 0]) / 1e3), "idle_drawn_to_folded_us": pct((e[:, 4] - e[:, 1]) / 1e3),
                              "idle_folded_to_ticket_us": pct((e[:, 6] - e[:, 4]) / 1e3)}), flush=True)
        eng.set_option("timing", 0)
        eng.set_option("prune", -1)


def probe_d(eng):
    """nearest neighbour: bucket-grid walk vs the grid-wide scan, same tour"""
    for n in (2000, 20000, 100000):
        eng.set_instance(uniform_instance(n), 0)
        res = {}
        for mode in (1, 0):
            eng.set_option("nn_grid", mode)
            eng.nn_tour(0)
            t0 = time.perf_counter()
            s, c = eng.nn_tour(0)
            res[mode] = (time.perf_counter() - t0, s, c)
        eng.set_option("nn_grid", -1)
        print(json.dumps({"probe": "d", "n": n, "grid_walk_s": round(res[1][0], 5), "full_scan_s": round(res[0][0], 5),
                          "same_tour": bool((res[0][1] == res[1][1]).all()), "same_cost": res[0][2] == res[1][2], "cost": res[1][2]}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["a", "b"]
    eng = Engine(0)
    if "a" in which:
        probe_a(eng)
    if "b" in which:
        probe_b(eng)
    if "c" in which:
        probe_c(eng)
    if "d" in which:
        probe_d(eng)
    eng.close()
