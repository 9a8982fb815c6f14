import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
n = 100000
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng.set_option("debug_shard", (world << 8) | 0)
eng.set_option("prune", 0)  # exhaustive passes: this is about the throughput shapes
shapes = [(0, 0, 0)] + [(64, 8, tj) for tj in (64, 96, 128, 160, 192, 224, 256)] + [(64, 4, tj) for tj in (128, 256)] + \
         [(128, 16, tj) for tj in (128, 192)] + [(128, 8, tj) for tj in (128, 256)]  # (64 x R: row-shuffle variant unless row_shuffle = 0)
for T, R, TJ in shapes:
    eng.set_option("block_threads", T); eng.set_option("rows_per_thread", R); eng.set_option("tile_cols", TJ)
    eng.tour_upload(succ)
    eng.bi_run(3)
    st = eng.bi_run(40)
    nt = eng.info("ntiles"); g = eng.info("grid_bi")
    print(json.dumps({"world": world, "T": eng.info("block_threads"), "R": eng.info("rows_per_thread"), "TJ": eng.info("tile_cols"), "row_shuffle": eng.info("row_shuffle"), "auto": T == 0,
                      "grid": g, "tiles_per_rank": -(-nt // world), "waves": -(-nt // world) / g, "us_per_pass": st.gpu_ms * 1e3 / st.passes}), flush=True)
