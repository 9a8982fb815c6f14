"""Dev tool: uni4000 HEU_2opt_greedy through the drop-in, three times in a row (first-use costs vs steady state), and the same
pieces through the engine API."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import RefLib
from tsp_optimization_b200 import Engine, FI
from tsp_optimization_b200.instances import uniform_instance
gpu = RefLib(gpu_link=True)
xy = uniform_instance(4000)
for k in range(3):
    t0 = time.perf_counter(); sg = gpu.run_method("HEU_2opt_greedy", xy, 0); tg = time.perf_counter() - t0
    print(json.dumps({"dropin_run": k, "s": round(tg, 5), "obj": sg[2]}), flush=True)
eng = Engine(0)
for k in range(3):
    t0 = time.perf_counter(); eng.set_instance(xy, 0); t1 = time.perf_counter()
    succ, cost = eng.nn_tour(0); t2 = time.perf_counter()
    s, obj, st, _ = eng.two_opt(FI, succ, cost); t3 = time.perf_counter()
    eng.dist_matrix(); t4 = time.perf_counter()
    print(json.dumps({"engine_run": k, "set_instance_ms": round((t1 - t0) * 1e3, 2), "nn_ms": round((t2 - t1) * 1e3, 2), "fi_ms": round((t3 - t2) * 1e3, 2),
                      "fi_gpu_ms": round(st.gpu_ms, 2), "moves": st.moves, "matrix_to_host_ms": round((t4 - t3) * 1e3, 2), "obj": obj}), flush=True)
    xy = xy + 0.0  # same values, new buffer
eng.close()
