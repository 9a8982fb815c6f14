"""GPU tuning sweep (dev tool): BI pass throughput vs tile shape / grid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time, json
import numpy as np
from tsp_optimization_b200 import Engine, BI
from tsp_optimization_b200.instances import uniform_instance

def main():
    ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [100000]
    eng = Engine(0)
    for n in ns:
        xy = uniform_instance(n)
        eng.set_instance(xy, 0)
        succ, _ = eng.nn_tour(0)
        pairs = n * (n - 3) // 2
        shapes = [(0, 0, 0)] + [(T, R, TJ) for T, R in ((256, 8), (128, 8), (64, 8), (64, 4), (256, 4)) for TJ in (32, 64, 96, 128, 256)]
        for T, R, TJ in shapes:
            if True:
                for fuse in ((0, 1) if n <= 20000 else (0,)):
                    grid = 0
                    eng.set_option("block_threads", T); eng.set_option("rows_per_thread", R); eng.set_option("tile_cols", TJ)
                    eng.set_option("fuse_apply", fuse)
                    eng.tour_upload(succ)
                    eng.bi_run(3)
                    k = 20 if n >= 50000 else 100
                    st = eng.bi_run(k)
                    print(json.dumps({"n": n, "T": eng.info("block_threads"), "R": eng.info("rows_per_thread"), "TJ": eng.info("tile_cols"), "auto": T == 0, "fuse": fuse, "grid": eng.info("grid_bi"), "tiles": eng.info("ntiles"),
                                      "ms_per_pass": st.gpu_ms / st.passes, "Gevals_s": st.passes * pairs / st.gpu_ms / 1e6}), flush=True)
main()
