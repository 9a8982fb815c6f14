"""Dev tool: cold-path calls and device time per BI pass."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
n = int(sys.argv[1]); passes = int(sys.argv[2]); seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = Engine(0)
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
eng.set_option("seed_hint", seed)
eng.tour_upload(succ)
prev = 0
rows = []
for k in range(passes):
    st = eng.bi_run(1)
    c = eng.info("cold_calls")
    rows.append((k, st.gpu_ms * 1e3, c - prev))
    prev = c
us = np.array([r[1] for r in rows]); cc = np.array([r[2] for r in rows])
print(json.dumps({"n": n, "seed": seed, "us_median": float(np.median(us)), "us_mean": float(us.mean()), "us_p10": float(np.percentile(us, 10)),
                  "us_p90": float(np.percentile(us, 90)), "cold_median": float(np.median(cc)), "cold_mean": float(cc.mean()), "cold_max": int(cc.max())}))
print(" ".join(f"{int(u)}/{c}" for _, u, c in rows[:60]))
