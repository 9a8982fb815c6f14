import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_optimization_b200 import Engine, FI
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
for n in (20000, 100000):
    xy = uniform_instance(n)
    eng.set_instance(xy, 0)
    succ, cost = eng.nn_tour(0)
    t0 = time.perf_counter()
    s, obj, st, _ = eng.two_opt(FI, succ, cost)
    dt = time.perf_counter() - t0
    print(json.dumps({"n": n, "mode": "FI", "wall_s": dt, "sweeps": st.passes, "moves": st.moves, "pairs_swept": st.evals, "cost": obj,
                      "us_per_move": dt * 1e6 / max(1, st.moves), "launches": st.launches}), flush=True)
