"""Dev tool: the one-block first-improvement kernel on the GA population (1024 x uni1000), repeated, kernel time vs wall time."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_optimization_b200 import BI, FI, Engine
from tsp_optimization_b200.instances import order_to_succ, reference_random_population, uniform_instance
eng = Engine(0)
xy = uniform_instance(1000)
eng.set_instance(xy, 0)
pop = reference_random_population(1000, 1024, 123)
succ = np.stack([order_to_succ(o) for o in pop])
costs = eng.tour_costs(succ, as_order=False)
for tours in (1024, 512, 128, 1024, 1024):
    for mode, nm in ((FI, "FI"), (BI, "BI")):
        t0 = time.perf_counter()
        sb, ob, st = eng.two_opt_batch(mode, succ[:tours], costs[:tours])
        wall = time.perf_counter() - t0
        print(json.dumps({"mode": nm, "tours": tours, "wall_ms": round(wall * 1e3, 2), "gpu_ms": round(st.gpu_ms, 2), "moves": st.moves, "sum": float(ob.sum())}), flush=True)
eng.close()
