import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
n = 100000
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
for world in (1, 2, 4, 8):
    for rank in sorted(set((0, world - 1))):
        eng.set_option("debug_shard", (world << 8) | rank if world > 1 else 0)
        eng.tour_upload(succ)
        eng.bi_run(3)
        st = eng.bi_run(40)
        print(json.dumps({"world": world, "rank": rank, "T": eng.info("block_threads"), "R": eng.info("rows_per_thread"), "TJ": eng.info("tile_cols"),
                          "grid": eng.info("grid_bi"), "tiles": eng.info("ntiles"), "us_per_pass": st.gpu_ms * 1e3 / st.passes,
                          "ideal_us": 1474.0 / world}), flush=True)
