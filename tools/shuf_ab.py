"""A/B of the best-improvement scan with and without the row shuffle (dev tool): exhaustive passes at n = 100 000 (the headline
shape and a few others) and pruned full runs at 10 000 / 100 000.  python tools/shuf_ab.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance

eng = Engine(0)
n = 100000
eng.set_instance(uniform_instance(n), 0)
succ = np.load("tests/golden/nn_uni100000.npz")["succ"]
pairs = n * (n - 3) // 2
for shape in ((0, 0, 0), (64, 8, 256), (64, 8, 128), (64, 4, 256), (64, 8, 512)):
    for shuf in (0, 1, 0, 1):
        eng.set_option("prune", 0)
        eng.set_option("block_threads", shape[0]); eng.set_option("rows_per_thread", shape[1]); eng.set_option("tile_cols", shape[2])
        eng.set_option("row_shuffle", shuf)
        eng.tour_upload(succ)
        eng.bi_run(5)
        eng.set_option("l2_flush_bytes", 256 << 20)
        st = eng.bi_run(30)
        eng.set_option("l2_flush_bytes", 0)
        s, cost = eng.tour_download()
        print(json.dumps({"n": n, "shape": [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")], "row_shuffle": eng.info("row_shuffle"),
                          "tile_rows": eng.info("tile_rows"), "us_per_pass": round(1e3 * st.gpu_ms / st.passes, 2),
                          "evals_per_s": round(st.passes * pairs / (st.gpu_ms * 1e-3) / 1e12, 4), "cost_after_35": cost}), flush=True)
eng.set_option("block_threads", 0); eng.set_option("rows_per_thread", 0); eng.set_option("tile_cols", 0)
for n2 in (10000, 100000):
    eng.set_instance(uniform_instance(n2), 0)
    s0 = succ if n2 == 100000 else eng.nn_tour(0)[0]
    for prune in (1, 0):
        if prune == 0 and n2 == 100000:
            continue
        for shuf in (0, 1):
            eng.set_option("prune", prune)
            eng.set_option("row_shuffle", shuf)
            eng.tour_upload(s0)
            eng.bi_run(8)
            eng.tour_upload(s0)
            st = eng.bi_run(-1)
            s, cost = eng.tour_download()
            print(json.dumps({"n": n2, "prune": prune, "row_shuffle": eng.info("row_shuffle"), "shape": [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")],
                              "passes": st.passes, "gpu_ms": round(st.gpu_ms, 2), "us_per_pass": round(1e3 * st.gpu_ms / st.passes, 2), "cost": cost,
                              "tiles_scanned": st.tiles_scanned, "tiles_total": st.tiles_total}), flush=True)
eng.close()
