import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
for n, k in ((10000, 300), (20000, 200), (100000, 40)):
    eng.set_instance(uniform_instance(n), 0)
    succ, _ = eng.nn_tour(0)
    pairs = n * (n - 3) // 2
    for seed, pdl in ((1, 0), (1, 1), (0, 1)):
        eng.set_option("seed_hint", seed)
        eng.set_option("pdl", pdl)
        eng.tour_upload(succ)
        eng.bi_run(3)
        st = eng.bi_run(k)
        print(json.dumps({"n": n, "seed": seed, "pdl": pdl, "us_per_pass": st.gpu_ms * 1e3 / st.passes, "Gevals_s": st.passes * pairs / st.gpu_ms / 1e6, "cold": eng.info("cold_calls")}), flush=True)
