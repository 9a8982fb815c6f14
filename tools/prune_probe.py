"""Exact tile pruning and the per-pass breakdown on one GPU: ms per pass, tiles scanned, and where a pass's time goes
(option "timing": %globaltimer stamps inside the kernels).  python tools/prune_probe.py [n ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from tsp_optimization_b200 import BI, Engine  # noqa: E402
from tsp_optimization_b200.instances import uniform_instance  # noqa: E402

TM = ["tm_gap", "tm_scan", "tm_spread", "tm_tail", "tm_xwait", "tm_apply_gap", "tm_apply", "tm_count"]


def breakdown(eng):
    c = max(1, eng.info("tm_count"))
    return {k[3:] + "_us": round(eng.info(k) / c / 1e3, 2) for k in TM[:-1]} | {"passes_timed": eng.info("tm_count")}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [10000, 100000]
    eng = Engine(0)
    for n in sizes:
        xy = uniform_instance(n)
        eng.set_instance(xy, 0)
        succ0, _ = eng.nn_tour(0)
        for prune, cap in ((0, 300 if n > 20000 else -1), (1, -1)):
            shapes = [None, (64, 8, 128), (64, 8, 64), (64, 4, 128), (64, 4, 64), (128, 8, 128)]
            if os.environ.get("PRUNE_SHAPES"):  # e.g. PRUNE_SHAPES=64x2x64,64x4x32
                shapes = [None] + [tuple(int(v) for v in t.split("x")) for t in os.environ["PRUNE_SHAPES"].split(",")]
            for shape in ([None] if prune == 0 else shapes):
                if shape:
                    eng.set_option("block_threads", shape[0]); eng.set_option("rows_per_thread", shape[1]); eng.set_option("tile_cols", shape[2])
                else:
                    eng.set_option("block_threads", 0); eng.set_option("rows_per_thread", 0); eng.set_option("tile_cols", 0)
                eng.set_option("prune", prune)
                eng.set_option("timing", 1)
                eng.tour_upload(succ0)
                eng.bi_run(8)  # warm-up
                eng.tour_upload(succ0)
                t0 = time.perf_counter()
                st = eng.bi_run(cap)
                wall = time.perf_counter() - t0
                s, cost = eng.tour_download()
                rec = {"n": n, "prune": prune, "shape": [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")],
                       "passes": st.passes, "moves": st.moves, "gpu_ms": round(st.gpu_ms, 3), "wall_s": round(wall, 4),
                       "us_per_pass": round(1e3 * st.gpu_ms / max(1, st.passes), 2), "cost": cost,
                       "tiles_scanned": st.tiles_scanned, "tiles_total": st.tiles_total,
                       "live_frac": round(st.tiles_scanned / st.tiles_total, 4) if st.tiles_total else None,
                       "breakdown": breakdown(eng), "cold_calls": eng.info("cold_calls")}
                print(json.dumps(rec), flush=True)
        eng.set_option("timing", 0)
        eng.set_option("prune", -1)
    eng.close()


if __name__ == "__main__":
    main()
