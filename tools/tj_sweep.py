"""Tile width vs pass time for one rank's share of the tiles (debug_shard emulation on one GPU): does a tile count just under a
multiple of the resident blocks beat the continuous cost model?  python tools/tj_sweep.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tsp_optimization_b200 import Engine  # noqa: E402
from tsp_optimization_b200.instances import uniform_instance  # noqa: E402

eng = Engine(0)
eng.set_option("prune", 0)
n = 100000
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
plan = {8: (96, 104, 112, 116, 120, 124, 128, 132, 136, 140, 148, 160), 4: (176, 184, 192, 200, 208, 216, 224, 240), 2: (232, 240, 248, 256, 264, 272),
        1: (240, 248, 252, 256, 260, 264, 272)}
for world, tjs in plan.items():
    for tj in tjs:
        for split in (0, 4):
            if split and tj % 16:
                continue
            eng.set_option("debug_shard", (world << 8) | (world - 1) if world > 1 else 0)
            eng.set_option("tile_cols", tj)
            eng.set_option("tail_split", split)
            eng.tour_upload(succ)
            eng.bi_run(5)
            best = 1e9
            for rep in range(2):
                st = eng.bi_run(50)
                best = min(best, 1e3 * st.gpu_ms / st.passes)
            tiles = eng.info("ntiles")
            per_rank = (tiles + world - 1) // world
            print(json.dumps({"world": world, "TJ": tj, "split": split, "us_per_pass": round(best, 2), "tiles_rank": per_rank,
                              "rounds": round(per_rank / eng.info("grid_bi"), 3), "ideal_us": round(1396.0 / world, 1)}), flush=True)
eng.close()
