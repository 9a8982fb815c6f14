"""Differential fuzzing of the CUDA path against the oracle: random instances (several coordinate regimes and metrics), random
start tours, random entry point / tile shape / caps.  Usage: python tools/fuzz.py <seconds> [seed]."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle
from tsp_optimization_b200 import Engine, BI, FI
from tsp_optimization_b200.instances import order_to_succ

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
orc, eng = Oracle(), Engine(0)
SHAPES = [(0, 0, 0), (64, 2, 32), (64, 4, 64), (64, 8, 88), (128, 8, 128), (128, 16, 64), (256, 8, 256), (256, 2, 36), (128, 4, 48)]


def coords(n):
    kind = rng.integers(0, 9)
    if kind == 0:
        return rng.integers(0, 10000, size=(n, 2)).astype(np.float64), "int1e4"
    if kind == 1:
        return rng.integers(0, 60, size=(n, 2)).astype(np.float64), "int60-ties"
    if kind == 2:
        return rng.integers(0, 2_000_000, size=(n, 2)).astype(np.float64), "int2e6"
    if kind == 3:
        return rng.integers(0, 4000, size=(n, 2)).astype(np.float64) / 2.0, "half"
    if kind == 4:
        return np.round(rng.random((n, 2)) * 5000.0, 3), "dec3"
    if kind == 5:
        c = rng.integers(0, 100000, size=(8, 2))
        return (c[rng.integers(0, 8, size=n)] + rng.integers(-30, 30, size=(n, 2))).astype(np.float64), "clustered"
    if kind == 6:
        return rng.integers(-5000, 5000, size=(n, 2)).astype(np.float64), "neg"
    if kind == 8:  # FP32-exact but NOT integer, 1e5..1e6 range: the matrix fast path without the integer shortcut
        return rng.integers(800_000, 8_000_000, size=(n, 2)).astype(np.float64) / 8.0, "eighths1e6"
    x = rng.integers(0, 3000, size=(n, 1)).astype(np.float64)
    return np.hstack([x, np.zeros((n, 1))]), "collinear"


t0 = time.time()
cases = fails = 0
stats = {}
while time.time() - t0 < budget:
    n = int(rng.choice([4, 5, 6, 7, 9, 16, 33, 64, 65, 100, 129, 257, 300, 500, 777, 1200, 2000],
                       p=[.06, .06, .05, .06, .06, .08, .08, .08, .06, .08, .06, .07, .06, .05, .04, .03, .02]))
    xy, kind = coords(n)
    wt = int(rng.choice([0, 0, 0, 3, 5, 4, 1, 2]))
    if wt == 4:
        xy = np.round((xy % 180.0) - 90.0 + (xy % 60) / 100.0, 2)
    succ = order_to_succ(rng.permutation(n).astype(np.int32)) if rng.random() < 0.5 else orc.nn_tour(xy, wt, int(rng.integers(0, n)))[0]
    cost0 = orc.succ_cost(xy, wt, succ)
    mode = BI if rng.random() < 0.6 else FI
    T, R, TJ = SHAPES[int(rng.integers(0, len(SHAPES)))]
    route = int(rng.choice([-1, 0, 1]))
    cap = -1 if ((rng.random() < 0.6 and n <= 777) or route == 1) else int(rng.integers(1, 30))
    if n > 777 and route == 1:
        route = 0
        cap = int(rng.integers(1, 30))
    eng.set_option("block_threads", T); eng.set_option("rows_per_thread", R); eng.set_option("tile_cols", TJ)
    eng.set_option("single_block", route)
    eng.set_option("prune", int(rng.choice([-1, 0, 1, 1])))          # exact tile pruning must not change a single move
    eng.set_option("batch_kernel", int(rng.integers(0, 2)))          # one-block BI: position-space / node-space kernel
    eng.set_option("row_shuffle", int(rng.choice([-1, 0, 1])))       # scan kernel: row below a lane's rows computed / shuffled from the next lane
    eng.set_option("fi_late", int(rng.integers(0, 2)))               # first improvement: winner selected by the apply launch / by the search kernel's tail
    eng.set_option("grid", int(rng.choice([0, 0, 1, 3, 17, 64])))  # few blocks -> many rounds of dynamically drawn tiles
    eng.set_instance(xy, wt)
    use_matrix = wt in (1, 2, 4) or rng.random() < 0.15
    if use_matrix:
        eng.dist_matrix_build()
        eng.set_option("force_path", 2 if rng.random() < 0.5 or wt in (1, 2, 4) else -1)
    try:
        if n <= 600 and rng.random() < 0.3:
            assert (eng.dist_matrix() == orc.dist_matrix(xy, wt)).all(), "matrix"
            if not use_matrix:
                eng.dist_matrix_free()
        if mode == BI:
            es, eobj, est, elog = orc.two_opt_bi(xy, wt, succ, max_passes=cap, log_cap=4096)
            s, obj, st, log = eng.two_opt(BI, succ, 0.0, max_iters=cap, log_cap=4096)
        else:
            es, eobj, est, elog = orc.two_opt_fi(xy, wt, succ, cost0, max_moves=cap, log_cap=4096)
            s, obj, st, log = eng.two_opt(FI, succ, cost0, max_iters=cap, log_cap=4096)
        ok = (s == es).all() and obj == eobj and log.tolist() == elog.tolist() and st.moves == est.moves
        assert ok, "2opt"
    except AssertionError as e:
        fails += 1
        np.savez(f"gpurun_out/fuzz_fail_{seed}_{cases}.npz", xy=xy, succ=succ)
        print(json.dumps({"FAIL": str(e), "case": cases, "n": n, "kind": kind, "wt": wt, "mode": mode, "shape": [T, R, TJ], "route": route, "cap": cap,
                          "matrix": bool(use_matrix)}), flush=True)
    eng.set_option("force_path", -1)
    eng.dist_matrix_free() if use_matrix else None
    cases += 1
    stats[kind] = stats.get(kind, 0) + 1
print(json.dumps({"cases": cases, "fails": fails, "seconds": time.time() - t0, "seed": seed, "by_kind": stats}), flush=True)
sys.exit(1 if fails else 0)
