import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
n = 100000
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
pairs = n * (n - 3) // 2
for T, R, TJ in ((256, 8, 256), (128, 16, 256), (128, 16, 128), (256, 16, 256), (128, 8, 256), (64, 8, 256), (128, 16, 512), (64, 8, 512), (256, 8, 512)):
    eng.set_option("block_threads", T); eng.set_option("rows_per_thread", R); eng.set_option("tile_cols", TJ)
    eng.tour_upload(succ)
    eng.bi_run(3)
    st = eng.bi_run(30)
    print(json.dumps({"T": T, "R": R, "TJ": TJ, "grid": eng.info("grid_bi"), "ms_per_pass": st.gpu_ms / st.passes, "Gevals_s": st.passes * pairs / st.gpu_ms / 1e6}), flush=True)
