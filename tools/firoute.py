import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine, FI, BI
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
for n in (128, 200, 300, 500, 1000, 2000, 3000, 4000, 6000):
    xy = uniform_instance(n)
    eng.set_instance(xy, 0)
    succ, cost = eng.nn_tour(0)
    row = {"n": n}
    for mode, nm in ((FI, "FI"), (BI, "BI")):
        for route in (0, 1):
            eng.set_option("single_block", route)
            eng.two_opt(mode, succ, cost)
            t0 = time.perf_counter()
            s, obj, st, _ = eng.two_opt(mode, succ, cost)
            row[f"{nm}_{'block' if route else 'grid'}_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
    print(json.dumps(row), flush=True)
