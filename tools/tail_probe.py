"""Where a best-improvement pass loses time at its end: one rank's share of the tiles emulated on ONE GPU (debug_shard), with
and without tail smoothing (sub-tiles for the last draws), plus the per-block time line of a one-wave pass (n = 10 000).
python tools/tail_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from tsp_optimization_b200 import Engine  # noqa: E402
from tsp_optimization_b200.instances import uniform_instance  # noqa: E402

TM = ["tm_gap", "tm_scan", "tm_spread", "tm_tail", "tm_xwait", "tm_apply_gap", "tm_apply"]


def run(eng, succ, passes, **opts):
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.set_option("timing", 1)
    eng.tour_upload(succ)
    eng.bi_run(4)
    eng.tour_upload(succ)
    st = eng.bi_run(passes)
    cnt = max(1, eng.info("tm_count"))
    return {"us_per_pass": round(1e3 * st.gpu_ms / st.passes, 2), "shape": [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")],
            "grid": eng.info("grid_bi"), "tiles": eng.info("ntiles"), **{k[3:]: round(eng.info(k) / cnt / 1e3, 2) for k in TM}}


def main():
    eng = Engine(0)
    eng.set_option("prune", 0)
    n = 100000
    eng.set_instance(uniform_instance(n), 0)
    succ, _ = eng.nn_tour(0)
    for world in (1, 8, 4, 2):
        shard = (world << 8) | (world - 1) if world > 1 else 0
        for tj in ((0, 128, 64) if world == 8 else (0,)):
            for split, tiles in ((0, 0), (2, 0), (4, 0), (4, 1184), (4, 296), (2, 1184)):
                if tj == 64 and split == 4 and False:
                    continue
                r = run(eng, succ, 60, debug_shard=shard, tile_cols=tj, tail_split=split, tail_tiles=tiles)
                print(json.dumps({"n": n, "world": world, "tail_split": split, "tail_tiles": tiles, "ideal_us": round(1396.0 / world, 1), **r}), flush=True)
    eng.set_option("debug_shard", 0)
    eng.set_option("tile_cols", 0)
    eng.set_option("tail_split", -1)
    eng.set_option("tail_tiles", 0)
    # one-wave pass: per-block time line
    for n in (10000, 20000):
        eng.set_instance(uniform_instance(n), 0)
        succ, _ = eng.nn_tour(0)
        for opts in ({}, {"tile_cols": 64}, {"tile_cols": 48}, {"rows_per_thread": 4, "tile_cols": 88}, {"fuse_apply": 1}):
            eng.set_option("tile_cols", 0); eng.set_option("rows_per_thread", 0); eng.set_option("fuse_apply", -1)
            r = run(eng, succ, 200, **opts)
            print(json.dumps({"n": n, "opts": opts, **r}), flush=True)
        eng.set_option("tile_cols", 0); eng.set_option("rows_per_thread", 0); eng.set_option("fuse_apply", -1)
        eng.set_option("timing", 2)
        eng.tour_upload(succ)
        eng.bi_run(20)
        bt = eng.block_times().astype(np.int64)
        t0 = bt[:, 0].min()
        start, end = (bt[:, 0] - t0) / 1e3, (bt[:, 1] - t0) / 1e3
        dur = end - start
        order = np.argsort(end)
        print(json.dumps({"n": n, "blocks": len(bt), "start_us_pct": np.percentile(start, [0, 50, 90, 100]).round(2).tolist(),
                          "end_us_pct": np.percentile(end, [0, 10, 50, 90, 99, 100]).round(2).tolist(),
                          "dur_us_pct": np.percentile(dur, [0, 10, 50, 90, 99, 100]).round(2).tolist(),
                          "slowest_blocks": order[-12:].tolist(), "slowest_end_us": end[order[-12:]].round(2).tolist(),
                          "fastest_blocks": order[:6].tolist()}), flush=True)
        eng.set_option("timing", 0)
    eng.close()


if __name__ == "__main__":
    main()
