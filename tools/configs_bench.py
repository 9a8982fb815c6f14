"""Dev tool: time-to-local-optimum and evals/s on the BASELINE.json configs other than the headline one."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_optimization_b200 import Engine, BI, FI
from tsp_optimization_b200.instances import uniform_instance, random_tours

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c4", "small"]
eng = Engine(0)
z = np.load("tests/golden/instances.npz")

def out(**kw):
    print(json.dumps(kw), flush=True)

if "small" in which:
    for nm in ("berlin52", "pr299", "att532", "gr666", "pr1002", "dsj1000"):
        xy, wt = z[nm + "__xy"], int(z[nm + "__wt"])
        eng.set_instance(xy, wt)
        eng.dist_matrix_build()
        ms = [eng.dist_matrix_build() for _ in range(5)]
        out(cfg="tsplib", inst=nm, n=len(xy), weight_type=wt, matrix_kernel_us=float(np.median(ms)) * 1e3)
        if wt != 4:
            eng.dist_matrix_free()  # GEO (all-FP64 trig) runs its 2-opt on the resident matrix
        succ, cost = eng.nn_tour(0)
        for mode, name in ((BI, "BI"), (FI, "FI")):
            eng.two_opt(mode, succ, cost)
            t0 = time.perf_counter()
            s, obj, st, _ = eng.two_opt(mode, succ, cost)
            dt = time.perf_counter() - t0
            out(cfg="tsplib", inst=nm, mode=name, wall_ms=dt * 1e3, gpu_ms=st.gpu_ms, passes=st.passes, moves=st.moves, cost=obj,
                evals_per_s=st.evals / max(st.gpu_ms, 1e-9) * 1e3)
if "c2" in which:
    n = 10000
    xy = uniform_instance(n)
    eng.set_instance(xy, 0)
    succ, cost = eng.nn_tour(0)
    for mode, name in ((BI, "BI"), (FI, "FI")):
        t0 = time.perf_counter()
        s, obj, st, _ = eng.two_opt(mode, succ, cost)
        dt = time.perf_counter() - t0
        out(cfg="c2 uni10000 NN->local optimum", mode=name, wall_s=dt, gpu_ms=st.gpu_ms, passes=st.passes, moves=st.moves, cost=obj,
            evals=st.evals, evals_per_s=st.evals / st.gpu_ms * 1e3, us_per_pass=st.gpu_ms * 1e3 / max(1, st.passes), launches=st.launches)
if "c4" in which:
    n, B = 1000, 1024
    xy = uniform_instance(n)
    eng.set_instance(xy, 0)
    tours = random_tours(n, B, 7)
    costs = eng.tour_costs(tours, as_order=False)
    for mode, name in ((BI, "BI"), (FI, "FI")):
        t0 = time.perf_counter()
        s, o, st = eng.two_opt_batch(mode, tours, costs)
        dt = time.perf_counter() - t0
        out(cfg="c4 GA batch uni1000 x1024 random tours -> local optimum", mode=name, wall_s=dt, gpu_ms=st.gpu_ms, passes=st.passes,
            moves=st.moves, evals=st.evals, evals_per_s=st.evals / st.gpu_ms * 1e3, mean_cost=float(o.mean()))
if "nnb" in which:
    for nm in ("pr1002", "gr666"):
        xy, wt = z[nm + "__xy"], int(z[nm + "__wt"])
        eng.set_instance(xy, wt)
        if wt == 4:
            eng.dist_matrix_build()
        eng.greedy_iter()
        t0 = time.perf_counter()
        best, succ, cost = eng.greedy_iter()
        out(cfg="greedy_iter (n NN runs)", inst=nm, wall_ms=(time.perf_counter() - t0) * 1e3, best_start=best, cost=cost)
        eng.dist_matrix_free()
if "nn" in which:
    for n in (10000, 100000):
        eng.set_instance(uniform_instance(n), 0)
        eng.nn_tour(0)
        t0 = time.perf_counter()
        eng.nn_tour(0)
        out(cfg="nn", n=n, wall_s=time.perf_counter() - t0)
