"""A/B of resident blocks per SM for the 64-thread scan kernel (dev tool): grid sizes of 8 / 9 / 10 blocks per SM, for the
library named by TSPB200_LIB (default build: register cap for 8 blocks; alt build: -DTSPB_BI_MINBLOCKS64=10).
python tools/occ_ab.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance

eng = Engine(0)
n = 100000
eng.set_instance(uniform_instance(n), 0)
succ = np.load("tests/golden/nn_uni100000.npz")["succ"]
pairs = n * (n - 3) // 2
eng.set_option("prune", 0)
for world in (1, 8):
    eng.set_option("debug_shard", (world << 8) if world > 1 else 0)
    for grid in (0, 1480, 1628, 1776, 0, 1480, 1628, 1776):
        eng.set_option("grid", grid)
        eng.tour_upload(succ)
        eng.bi_run(5)
        eng.set_option("l2_flush_bytes", 256 << 20)
        st = eng.bi_run(30)
        eng.set_option("l2_flush_bytes", 0)
        print(json.dumps({"lib": os.environ.get("TSPB200_LIB", "default"), "world": world, "grid": eng.info("grid_bi"),
                          "shape": [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")],
                          "us_per_pass": round(1e3 * st.gpu_ms / st.passes, 2)}), flush=True)
eng.close()
