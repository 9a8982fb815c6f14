import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
for n, k in ((1000, 100), (2000, 100), (5000, 200), (10000, 300), (20000, 200), (50000, 60), (100000, 40)):
    eng.set_instance(uniform_instance(n), 0)
    succ, _ = eng.nn_tour(0)
    pairs = n * (n - 3) // 2
    eng.tour_upload(succ)
    eng.bi_run(3)
    st = eng.bi_run(k)
    print(json.dumps({"n": n, "T": eng.info("block_threads"), "R": eng.info("rows_per_thread"), "TJ": eng.info("tile_cols"), "grid": eng.info("grid_bi"),
                      "us_per_pass": st.gpu_ms * 1e3 / st.passes, "Gevals_s": st.passes * pairs / st.gpu_ms / 1e6}), flush=True)
