"""CPU tests: the oracle restatement (oracle/tsp_oracle.c) against (a) the committed golden fixtures that
were generated from the unmodified reference, (b) the reference's own published CSV goldens, and (c) the
compiled reference itself when oracle/_ref is present."""
import hashlib

import numpy as np
import pytest

from tsp_optimization_b200.instances import is_tour, order_to_succ, uniform_instance


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SMALL = ["berlin52", "att48", "burma14", "ulysses16", "ulysses22", "eil51", "gr96", "pr299", "a280", "gr202", "gr229"]


def test_golden_set_covers_all_metrics(goldens):
    wts = {g["wt"] for g in goldens["instances"].values()}
    assert {0, 3, 4, 5} <= wts  # EUC_2D, CEIL_2D, GEO, ATT


def test_matrix_matches_golden(oracle, instances, goldens):
    for nm, (xy, wt) in instances.items():
        g = goldens["instances"][nm]
        m = oracle.dist_matrix(xy, wt)
        assert int(m.sum(dtype=np.int64)) == g["matrix_sum"], nm
        assert sha(m) == g["matrix_sha256"], nm


def test_geo_diagonal_is_one(oracle, instances):
    xy, wt = instances["gr666"]
    assert wt == 4
    assert oracle.dist(xy, wt, 0, 0) == 1.0  # reference distutil.c:69 "+ 1.0" then nint


def test_man_max_follow_the_reference_bug(oracle):
    xy = np.array([[0.0, 0.0], [3.0, 40.0]])
    assert oracle.dist(xy, 2, 0, 1) == 3.0  # MAN_2D: dy = fabs(p2.y - p2.y) == 0
    assert oracle.dist(xy, 1, 0, 1) == 3.0  # MAX_2D
    assert oracle.dist(xy, 99, 0, 1) == 40.0  # unknown weight type -> EUC_2D: nint(40.11)


def test_nn_fi_bi_match_golden(oracle, instances, goldens):
    for nm, (xy, wt) in instances.items():
        g = goldens["instances"][nm]
        succ, cost = oracle.nn_tour(xy, wt, 0)
        assert cost == g["nn_cost"] and sha(succ) == g["nn_sha256"], nm
        if g["n"] > 700:
            continue
        fs, fc, fst, _ = oracle.two_opt_fi(xy, wt, succ, cost)
        assert fc == g["fi_cost"] and sha(fs) == g["fi_sha256"], nm
        assert (fst.moves, fst.passes, fst.evals) == (g["fi_moves"], g["fi_sweeps"], g["fi_evals"]), nm
        if nm in SMALL and "bi_cost" in g:
            bs, bc, bst, _ = oracle.two_opt_bi(xy, wt, succ)
            assert bc == g["bi_cost"] and sha(bs) == g["bi_sha256"], nm
            assert (bst.moves, bst.evals) == (g["bi_moves"], g["bi_evals"]), nm


def test_reference_csv_goldens(oracle, instances, goldens):
    """results/constructive_heuristics_new.csv GREEDY and ..._2opt_new.csv 2OPT_GREEDY (18 instances)."""
    csvg = goldens["reference_csv"]
    assert len(csvg) == 18
    for nm in ["lin318", "rd400", "pcb442", "att532", "ali535", "gr431", "u574"]:
        xy, wt = instances[nm]
        succ, cost = oracle.nn_tour(xy, wt, 0)
        assert cost == csvg[nm]["GREEDY"], nm
        _, fc, _, _ = oracle.two_opt_fi(xy, wt, succ, cost)
        assert fc == csvg[nm]["2OPT_GREEDY"], nm


def test_berlin52_move_logs(oracle, instances, goldens):
    xy, wt = instances["berlin52"]
    succ, cost = oracle.nn_tour(xy, wt, 0)
    assert cost == 8980
    bs, bc, bst, blog = oracle.two_opt_bi(xy, wt, succ, log_cap=100)
    assert bc == 7842 and bst.moves == 11 and bst.evals == 15288
    assert blog.tolist() == goldens["instances"]["berlin52"]["bi_log"]
    assert blog[:3].tolist() == [[10, 50, -257], [28, 45, -267], [0, 20, -231]]
    fs, fc, fst, flog = oracle.two_opt_fi(xy, wt, succ, cost, log_cap=100)
    assert fc == 8083 and fst.moves == 20 and fst.passes == 5 and fst.evals == 6380
    assert flog.tolist() == goldens["instances"]["berlin52"]["fi_log"]
    assert is_tour(bs) and is_tour(fs)


def test_oracle_equals_compiled_reference(oracle, reflib, instances):
    for nm in ["berlin52", "att48", "ulysses22", "pr299", "gr96"]:
        xy, wt = instances[nm]
        assert (oracle.dist_matrix(xy, wt) == reflib.dist_matrix(xy, wt)).all()
        s1, c1 = oracle.nn_tour(xy, wt, 0)
        s2, c2 = reflib.nn_tour(xy, wt, 0)
        assert (s1 == s2).all() and c1 == c2
        f1 = oracle.two_opt_fi(xy, wt, s1, c1)
        f2 = reflib.two_opt_fi(xy, wt, s1, c1)
        assert (f1[0] == f2[0]).all() and f1[1] == f2[1]
        b1 = oracle.two_opt_bi(xy, wt, s1, want_prev=True)
        b2 = reflib.two_opt_bi(xy, wt, s1, want_prev=True)
        assert (b1[0] == b2[0]).all() and b1[1] == b2[1] and (b1[4] == b2[2]).all()


def test_oracle_equals_reference_on_random_tours_and_coords(oracle, reflib):
    rng = np.random.default_rng(7)
    for wt, n in [(0, 120), (3, 90), (5, 100), (4, 60)]:
        if wt == 4:
            xy = np.round(rng.uniform(-80, 80, size=(n, 2)), 2)
        else:
            xy = rng.integers(0, 3000, size=(n, 2)).astype(np.float64)
            xy[::7] += 0.5  # non-integer but FP32-exact
        succ = order_to_succ(rng.permutation(n).astype(np.int32))
        c = oracle.succ_cost(xy, wt, succ)
        f1 = oracle.two_opt_fi(xy, wt, succ, c)
        f2 = reflib.two_opt_fi(xy, wt, succ, c)
        assert (f1[0] == f2[0]).all() and f1[1] == f2[1]
        b1 = oracle.two_opt_bi(xy, wt, succ)
        b2 = reflib.two_opt_bi(xy, wt, succ)
        assert (b1[0] == b2[0]).all() and b1[1] == b2[1]


def test_tabu_mask_restatement(oracle, reflib):
    """alg_2opt_tabu with a tabu list: same tour AND same mutated list (lazy expiry, tabusearch.c:83-92)."""
    rng = np.random.default_rng(11)
    n = 60
    xy = rng.integers(0, 1000, size=(n, 2)).astype(np.float64)
    succ = order_to_succ(rng.permutation(n).astype(np.int32))
    ncols = n * (n - 1) // 2
    mask = np.where(rng.random(ncols) < 0.2, rng.integers(1, 30, size=ncols), 0).astype(np.int32)
    m1, m2 = mask.copy(), mask.copy()
    b1 = oracle.two_opt_bi(xy, 0, succ, skip_edge=m1, iter_=30, tenure=12)
    b2 = reflib.two_opt_bi(xy, 0, succ, skip_edge=m2, iter_=30, tenure=12)
    assert (b1[0] == b2[0]).all() and b1[1] == b2[1]
    assert (m1 == m2).all()


def test_synthetic_known_answers(oracle):
    """uni2000 NN 406727 -> FI 345191 (6 sweeps, 630 moves)."""
    xy = uniform_instance(2000)
    succ, cost = oracle.nn_tour(xy, 0, 0)
    assert cost == 406727
    _, fc, st, _ = oracle.two_opt_fi(xy, 0, succ, cost)
    assert fc == 345191 and st.passes == 6 and st.moves == 630 and st.evals == 11982230


def test_bi_scan_rows_mt_matches_first_move(oracle, instances):
    xy, wt = instances["pr299"]
    succ, _ = oracle.nn_tour(xy, wt, 0)
    ev, sec, key = oracle.bi_scan_rows_mt(xy, wt, succ, 0, 299, 3)
    _, _, st, log = oracle.two_opt_bi(xy, wt, succ, max_passes=1, log_cap=1)
    assert ev == 299 * 296 // 2 == st.evals
    assert [key[1], key[2], key[0]] == log[0].tolist()


def test_csv_goldens_cover_every_deterministic_column(goldens):
    """18 instances x {GREEDY, GREEDY_ITER, EXTR_MILE, 2OPT_GREEDY, 2OPT_GREEDY_ITER, 2OPT_EXTR_MIL}: every cell was re-derived
    with the compiled reference when the fixture was made (tests/golden/make_goldens.py, make_csv_goldens.py)."""
    assert len(goldens["reference_csv"]) == 18
    for nm, row in goldens["reference_csv"].items():
        assert set(row) == {"GREEDY", "GREEDY_ITER", "EXTR_MILE", "2OPT_GREEDY", "2OPT_GREEDY_ITER", "2OPT_EXTR_MIL"}, nm


def test_large_fixture_spot_check(oracle):
    """tests/golden/large.npz / goldens_large.json (decimal and 10^6-range coordinates): the cheap cells re-derived here."""
    import json
    import os
    from conftest import GOLD_DIR
    z = np.load(os.path.join(GOLD_DIR, "large.npz"))
    g = json.load(open(os.path.join(GOLD_DIR, "goldens_large.json")))["instances"]
    assert set(g) == {"fl3795", "pla7397", "usa13509", "stefano_8k"}
    xy, wt = z["fl3795__xy"], int(z["fl3795__wt"])
    succ, cost = oracle.nn_tour(xy, wt, 0)
    assert cost == g["fl3795"]["nn_cost"]
    s, obj, st, log = oracle.two_opt_bi(xy, wt, succ, max_passes=3, log_cap=8)
    assert log.tolist() == g["fl3795"]["bi_log"][:3]


def test_constructive_restatements_equal_reference_and_csv(oracle, reflib, instances, goldens):
    """orc_greedy_iter / orc_extra_mileage (reference HEU_Greedy_iter, HEU_extramileage) vs the compiled reference (tour and
    cost) and vs the published GREEDY_ITER / EXTR_MILE columns."""
    for nm in ("berlin52", "pr299", "att532", "ulysses22", "gr96", "eil51"):
        xy, wt = instances[nm]
        b, s, c = oracle.greedy_iter(xy, wt)
        st, rs, rc = reflib.run_method("HEU_Greedy_iter", xy, wt)
        assert st == 0 and c == rc and (s == rs).all(), nm
        s, c = oracle.extra_mileage(xy, wt)
        st, rs, rc = reflib.run_method("HEU_extramileage", xy, wt)
        assert st == 0 and c == rc and (s == rs).all(), nm
        if nm in goldens["reference_csv"]:
            assert c == goldens["reference_csv"][nm]["EXTR_MILE"]
    rng = np.random.default_rng(4)
    for n in (2, 3, 5, 30):
        xy = rng.integers(0, 7, size=(n, 2)).astype(np.float64)  # ties and coincident nodes
        s, c = oracle.extra_mileage(xy, 0)
        st, rs, rc = reflib.run_method("HEU_extramileage", xy, 0)
        assert c == rc and (s == rs).all(), n


def test_vns_kick_restatement_equals_the_reference_kick(oracle, reflib):
    """oracle.vns_kick (indices handed in) against the reference's own kick() (src/vns.c:11), which draws its three indices
    from glibc random(): the draws are replayed with the restated generator.  Seeds whose largest index is n-1 are skipped —
    there the reference reads one element past its tour array."""
    from tsp_optimization_b200.instances import GlibcRandom, is_tour, uniform_instance
    n = 300
    xy = uniform_instance(n)
    succ, _ = oracle.nn_tour(xy, 0, 0)
    compared = 0
    for seed in range(1, 40):
        g = GlibcRandom(seed)
        i1 = g.rand_choice(0, n)
        i2 = i3 = i1
        while i2 == i1 or abs(i1 - i2) <= 1:
            i2 = g.rand_choice(0, n)
        while i3 == i1 or i3 == i2 or abs(i1 - i3) <= 1 or abs(i2 - i3) <= 1:
            i3 = g.rand_choice(0, n)
        if max(i1, i2, i3) == n - 1:
            continue
        s_ref, c_ref = reflib.vns_kick_stock(xy, 0, succ, seed)
        s_o, c_o = oracle.vns_kick(xy, 0, succ, i1, i2, i3)
        assert (s_ref == s_o).all() and c_ref == c_o and is_tour(s_o), seed
        compared += 1
    assert compared >= 30
