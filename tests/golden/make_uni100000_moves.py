"""Generates tests/golden/uni100000_bi_moves.npz — the move log bench.py checks every rank's tour against.

Run on a B200 box (python tests/golden/make_uni100000_moves.py): uni100000 (SURVEY.md §8c generator), GPU nearest-neighbour
start (must equal the committed oracle fixture nn_uni100000.npz), best-improvement 2-opt on ONE GPU:
  moves[P, 3]   the first P = 1200 applied moves (i, j, delta), exhaustive scan
  final_*       passes / moves / cost / sha256 of the tour at the local optimum; the exhaustive run and the run with exact
                tile pruning must agree before anything is written
  oracle_moves  how many leading moves were re-derived with the CPU oracle (one full multi-threaded scan each) and found equal
Only this script and the tests use oracle/ here; the bench reads the .npz alone."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tsp_optimization_b200 import BI, Engine  # noqa: E402
from tsp_optimization_b200.instances import apply_moves, uniform_instance  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()


def main():
    n, P = 100000, 1200
    oracle_moves = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    xy = uniform_instance(n)
    eng = Engine(0)
    eng.set_instance(xy, 0)
    succ0, c0 = eng.nn_tour(0)
    nn = np.load(os.path.join(ROOT, "tests", "golden", "nn_uni100000.npz"))
    assert (succ0 == nn["succ"]).all(), "GPU nearest-neighbour tour differs from the oracle fixture"
    eng.set_option("prune", 0)
    s_p, _, st_p, log = eng.two_opt(BI, succ0, 0.0, max_iters=P, log_cap=P)
    assert st_p.passes == P and len(log) == P
    assert (apply_moves(succ0, log) == s_p).all(), "host replay of the move log differs from the device tour"
    finals = []
    for prune in (1, 0):
        eng.set_option("prune", prune)
        s, obj, st, lg = eng.two_opt(BI, succ0, 0.0, log_cap=P)
        assert lg.tolist() == log.tolist()
        finals.append((sha(s), st.passes, st.moves, obj))
        print("prune", prune, finals[-1], "gpu_ms", st.gpu_ms, flush=True)
    assert finals[0] == finals[1], finals
    checked = 0
    if oracle_moves:
        from oracle.oracle import Oracle
        orc = Oracle()
        cur = succ0.copy()
        for k in range(oracle_moves):
            ev, sec, key = orc.bi_scan_rows_mt(xy, 0, cur, 0, n - 1, os.cpu_count() or 1)
            assert [key[1], key[2], key[0]] == log[k].tolist(), (k, key, log[k])
            cur = apply_moves(cur, log[k:k + 1])
            checked += 1
            print("oracle move", k, key, f"{sec:.1f} s", flush=True)
    out = os.path.join(ROOT, "tests", "golden", "uni100000_bi_moves.npz")
    np.savez_compressed(out, moves=log.astype(np.int32), final_passes=finals[0][1], final_moves=finals[0][2], final_cost=finals[0][3],
                        final_sha256=finals[0][0], nn_sha256=sha(succ0), nn_cost=c0, oracle_moves=checked)
    # the bench writes results under gpurun_out/ on the GPU box; copy the fixture there so that it travels back
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "uni100000_bi_moves.npz"), moves=log.astype(np.int32), final_passes=finals[0][1],
                        final_moves=finals[0][2], final_cost=finals[0][3], final_sha256=finals[0][0], nn_sha256=sha(succ0), nn_cost=c0,
                        oracle_moves=checked)
    print("wrote", out)


if __name__ == "__main__":
    main()
