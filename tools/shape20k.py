import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
for n in (20000, 30000):
    eng.set_instance(uniform_instance(n), 0)
    succ, _ = eng.nn_tour(0)
    for T, R, TJ in [(0, 0, 0)] + [(64, 8, tj) for tj in (64, 80, 96, 128, 176, 256)] + [(128, 8, 176), (128, 8, 256)]:
        eng.set_option("block_threads", T); eng.set_option("rows_per_thread", R); eng.set_option("tile_cols", TJ)
        eng.tour_upload(succ)
        eng.bi_run(3)
        st = eng.bi_run(100)
        print(json.dumps({"n": n, "T": eng.info("block_threads"), "R": eng.info("rows_per_thread"), "TJ": eng.info("tile_cols"), "auto": T == 0,
                          "tiles": eng.info("ntiles"), "grid": eng.info("grid_bi"), "us_per_pass": round(st.gpu_ms * 1e3 / st.passes, 1)}), flush=True)
