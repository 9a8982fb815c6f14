"""Dev tool: a short tour through every kernel for compute-sanitizer (small sizes; results still checked against the oracle)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle
from tsp_optimization_b200 import Engine, BI, FI
from tsp_optimization_b200.instances import uniform_instance, random_tours

orc = Oracle()
eng = Engine(0)
z = np.load("tests/golden/instances.npz")
for nm in ("berlin52", "pr299", "att48", "ulysses22"):
    xy, wt = z[nm + "__xy"], int(z[nm + "__wt"])
    eng.set_instance(xy, wt)
    assert (eng.dist_matrix() == orc.dist_matrix(xy, wt)).all()
    if wt != 4:
        eng.dist_matrix_free()
    succ, cost = eng.nn_tour(0)
    for route in (0, 1):
        eng.set_option("single_block", route)
        for mode in (BI, FI):
            s, obj, st, log = eng.two_opt(mode, succ, cost if mode == FI else 0.0, log_cap=500)
            if mode == BI:
                es, eobj, _, elog = orc.two_opt_bi(xy, wt, succ, log_cap=500)
            else:
                es, eobj, _, elog = orc.two_opt_fi(xy, wt, succ, cost, log_cap=500)
            assert (s == es).all() and obj == eobj and log.tolist() == elog.tolist(), (nm, route, mode)
    eng.set_option("single_block", -1)
    sb, costs = eng.nn_tour_batch(np.arange(min(len(xy), 16), dtype=np.int32))
    eng.dist_matrix_free()
xy = uniform_instance(3000)
eng.set_instance(xy, 0)
succ, cost = eng.nn_tour(0)
for T, R, TJ in ((0, 0, 0), (256, 8, 64), (128, 16, 64), (64, 2, 32)):
    eng.set_option("block_threads", T); eng.set_option("rows_per_thread", R); eng.set_option("tile_cols", TJ)
    s, obj, st, log = eng.two_opt(BI, succ, 0.0, max_iters=6, log_cap=16)
    es, eobj, _, elog = orc.two_opt_bi(xy, 0, succ, max_passes=6, log_cap=16)
    assert (s == es).all() and log.tolist() == elog.tolist()
eng.set_option("block_threads", 0); eng.set_option("rows_per_thread", 0); eng.set_option("tile_cols", 0)
eng.tour_upload(succ)
eng.fi_run(20)
n = 120
xy = uniform_instance(n)
eng.set_instance(xy, 0)
mask = np.zeros(n * (n - 1) // 2, dtype=np.int32); mask[::7] = 3
eng.two_opt_tabu(orc.nn_tour(xy, 0, 0)[0], mask, 10, 4)
tours = random_tours(n, 8, 1)
eng.two_opt_batch(FI, tours, eng.tour_costs(tours, as_order=False))
eng.two_opt_batch(BI, tours)
eng.close()
print("sanitize tour OK")
