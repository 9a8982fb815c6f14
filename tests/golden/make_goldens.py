"""Generates tests/golden/{instances.npz,goldens.json} IN THE BUILD CONTAINER from the unmodified reference
(oracle/_ref/libtspref.so, compiled from /root/reference/src) and the reference's published result CSVs.
The GPU box has no /root/reference, so GPU-side parity tests read these committed fixtures.

    python tests/golden/make_goldens.py
"""
import csv
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, RefLib  # noqa: E402

REF = "/root/reference"
HEUR = ["ali535", "att532", "d493", "d657", "dsj1000", "gr431", "gr666", "lin318", "p654", "pcb442", "pr1002",
        "pr439", "rat575", "rat783", "rd400", "u574", "u724", "vm1084"]
TOP = ["berlin52", "pr299", "att48", "burma14", "ulysses16", "ulysses22", "gr96", "gr202", "gr229", "eil51", "a280"]
ALL_GEO_EXTRA = {"gr137": "data/all/gr137.tsp"}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref, orc = RefLib(), Oracle()
    paths = {}
    for nm in HEUR:
        paths[nm] = f"{REF}/data/heuristics/{nm}.tsp"
    for nm in TOP:
        paths[nm] = f"{REF}/data/{nm}.tsp"
    for nm, p in ALL_GEO_EXTRA.items():
        if os.path.exists(f"{REF}/{p}"):
            paths[nm] = f"{REF}/{p}"
    inst, gold = {}, {}
    for nm, p in sorted(paths.items()):
        xy, wt = ref.parse(p)
        inst[nm + "__xy"] = xy
        inst[nm + "__wt"] = np.int32(wt)
        n = len(xy)
        g = {"n": n, "wt": int(wt)}
        m = ref.dist_matrix(xy, wt)
        g["matrix_sum"] = int(m.sum(dtype=np.int64))
        g["matrix_sha256"] = sha(m)
        succ, nn_cost = ref.nn_tour(xy, wt, 0)
        g["nn_cost"] = nn_cost
        g["nn_sha256"] = sha(succ)
        fs, fc = ref.two_opt_fi(xy, wt, succ, nn_cost)
        g["fi_cost"] = fc
        g["fi_sha256"] = sha(fs)
        o_fs, o_fc, o_st, _ = orc.two_opt_fi(xy, wt, succ, nn_cost)
        assert (o_fs == fs).all() and o_fc == fc, nm
        g["fi_moves"], g["fi_sweeps"], g["fi_evals"] = int(o_st.moves), int(o_st.passes), int(o_st.evals)
        if n <= 700 or nm in ("pr1002", "dsj1000"):
            bs, bc = ref.two_opt_bi(xy, wt, succ)
            o_bs, o_bc, o_bst, _ = orc.two_opt_bi(xy, wt, succ)
            assert (o_bs == bs).all() and o_bc == bc, nm
            g["bi_cost"], g["bi_sha256"] = bc, sha(bs)
            g["bi_moves"], g["bi_evals"] = int(o_bst.moves), int(o_bst.evals)
        gold[nm] = g
        print(nm, g)
    # full move logs on berlin52 (oracle restatement, validated against the reference's final state above)
    xy, wt = inst["berlin52__xy"], int(inst["berlin52__wt"])
    succ, c = ref.nn_tour(xy, wt, 0)
    gold["berlin52"]["bi_log"] = orc.two_opt_bi(xy, wt, succ, log_cap=1000)[3].tolist()
    gold["berlin52"]["fi_log"] = orc.two_opt_fi(xy, wt, succ, c, log_cap=1000)[3].tolist()
    # the reference's own published goldens (results/*.csv, current code revision)
    csvg = {}
    for fn, col, key in (("constructive_heuristics_new.csv", "GREEDY", "GREEDY"),
                         ("constructive_heuristics_2opt_new.csv", "2OPT_GREEDY", "2OPT_GREEDY")):
        with open(f"{REF}/results/{fn}") as f:
            rd = csv.reader(f)
            head = next(rd)
            ci = head.index(col)
            for row in rd:
                nm = os.path.basename(row[0]).replace(".tsp", "")
                csvg.setdefault(nm, {})[key] = float(row[ci])
    for nm in HEUR:
        assert csvg[nm]["GREEDY"] == gold[nm]["nn_cost"], (nm, csvg[nm], gold[nm]["nn_cost"])
        assert csvg[nm]["2OPT_GREEDY"] == gold[nm]["fi_cost"], (nm, csvg[nm], gold[nm]["fi_cost"])
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "instances.npz"), **inst)
    with open(os.path.join(here, "goldens.json"), "w") as f:
        json.dump({"instances": gold, "reference_csv": csvg,
                   "source": "oracle/_ref/libtspref.so built from /root/reference/src; results/constructive_heuristics_new.csv, "
                             "results/constructive_heuristics_2opt_new.csv"}, f, indent=1, sort_keys=True)
    print("wrote", len(gold), "instances")


if __name__ == "__main__":
    main()
