"""Dev tool: exhaustive best-improvement pass at mid sizes, automatic shape vs forced rows per thread."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
eng.set_option("prune", 0)
for n in (10000, 20000, 40000):
    eng.set_instance(uniform_instance(n), 0)
    succ, _ = eng.nn_tour(0)
    for R in (0, 8, 4, 0, 8, 4):
        eng.set_option("rows_per_thread", R)
        eng.tour_upload(succ)
        eng.bi_run(10)
        st = eng.bi_run(200)
        print(json.dumps({"n": n, "forced_R": R, "shape": [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")],
                          "grid": eng.info("grid_bi"), "tiles": eng.info("ntiles"), "us_per_pass": round(1e3 * st.gpu_ms / st.passes, 2)}), flush=True)
    eng.set_option("rows_per_thread", 0)
eng.close()
