"""Resident sessions (include/tspb200.h "resident sessions"): the perturbation steps of the reference's VNS / tabu / GA
drivers applied to device-resident state, replayed iteration by iteration against the CPU oracle with the reference's own
random stream (glibc random(), restated in tsp_optimization_b200/instances.py GlibcRandom and pinned to libc in
tests/test_host_logic.py)."""
import numpy as np
import pytest

from tsp_optimization_b200 import engine as eng
from tsp_optimization_b200.instances import (GlibcRandom, is_tour, order_to_succ, reference_random_population, succ_to_order,
                                             uniform_instance)

BI, FI = eng.BI, eng.FI
pytestmark = pytest.mark.gpu


def _draw_kick(g, n):
    """reference src/vns.c:25-31"""
    i1 = g.rand_choice(0, n)
    i2 = i3 = i1
    while i2 == i1 or abs(i1 - i2) <= 1:
        i2 = g.rand_choice(0, n)
    while i3 == i1 or i3 == i2 or abs(i1 - i3) <= 1 or abs(i2 - i3) <= 1:
        i3 = g.rand_choice(0, n)
    return i1, i2, i3


@pytest.mark.parametrize("n,wt,iters", [(400, 0, 30), (150, 5, 25), (2500, 0, 6)])
def test_vns_loop_replayed_on_the_resident_tour(engine, oracle, n, wt, iters):
    """HEU_VNS (reference src/vns.c:125-177): kick the incumbent, alg_2opt, keep it if better, else restore — the tour never
    leaves the device between the steps; every iteration is compared with the oracle (tour, kicked cost, 2-opt cost)."""
    rng = np.random.default_rng(n)
    xy = uniform_instance(n) if wt == 0 else rng.integers(0, 5000, size=(n, 2)).astype(np.float64)
    engine.set_instance(xy, wt)
    engine.set_option("single_block", 0)
    try:
        start, c0 = oracle.nn_tour(xy, wt, 0)
        best_succ, best_obj, _, _ = oracle.two_opt_fi(xy, wt, start, c0)
        engine.tour_upload(start)
        st = engine.fi_run(-1)
        assert c0 + st.obj_delta == best_obj and engine.tour_cost() == best_obj
        engine.tour_save(0)
        g = GlibcRandom(123)
        improved = 0
        for it in range(iters):
            i1, i2, i3 = _draw_kick(g, n)
            if it == 3:  # the wrap-around case: the last tour index as the largest one
                i1, i2, i3 = 0, n // 2, n - 1
            ks, kc = oracle.vns_kick(xy, wt, best_succ, i1, i2, i3)
            assert engine.vns_kick(i1, i2, i3) == kc, (it, i1, i2, i3)
            s_dev, c_dev = engine.tour_download()
            assert (s_dev == ks).all() and c_dev == kc and is_tour(s_dev), it
            os_, oobj, ost, _ = oracle.two_opt_fi(xy, wt, ks, kc)
            st = engine.fi_run(-1)
            s_dev, c_dev = engine.tour_download()
            assert st.status == eng.LOCAL_OPTIMUM and st.moves == ost.moves and st.passes == ost.passes, it
            assert (s_dev == os_).all() and c_dev == oobj == kc + st.obj_delta, it
            if oobj < best_obj:
                best_obj, best_succ = oobj, os_
                engine.tour_save(0)
                improved += 1
            else:
                engine.tour_restore(0)
                s_dev, c_dev = engine.tour_download()
                assert (s_dev == best_succ).all() and c_dev == best_obj
        assert improved >= 1
    finally:
        engine.set_option("single_block", -1)


def test_tabu_loop_replayed_with_the_list_resident(engine, oracle):
    """tabu() (reference src/tabusearch.c:228-311) for 35 iterations: masked best improvement, the random kick with its
    check_tenure tests and lazy expiry, the two removed edges stamped with the iteration, a stepping tenure.  The device
    keeps tour and list; the host only draws the candidates (one per call, like the reference's loop, so the random stream
    stays in step).  Compared with the oracle every iteration (tour, cost, accepted candidate) and at the end (the list)."""
    n = 220
    xy = uniform_instance(n)
    engine.set_instance(xy, 0)
    succ_o, _ = oracle.nn_tour(xy, 0, 0)
    tl_o = np.zeros(n * (n - 1) // 2, dtype=np.int32)
    engine.tour_upload(succ_o)
    engine.tabu_begin()
    g = GlibcRandom(7)
    tenure, best = 3, None
    for it in range(1, 36):
        succ_o, obj_o, ost, _ = oracle.two_opt_bi(xy, 0, succ_o, skip_edge=tl_o, iter_=it, tenure=tenure)
        st = engine.tabu_run(it, tenure)
        s_dev, _ = engine.tour_download()
        assert (s_dev == succ_o).all() and st.cost == obj_o and st.moves == ost.moves and st.passes == ost.passes, it
        if best is None or obj_o < best:
            best = obj_o
            engine.tour_save(1)
        tries = 0
        while True:
            a, b = g.rand_choice(0, n), g.rand_choice(0, n)
            tries += 1
            acc_o, succ_o, tl_o = oracle.tabu_kick(succ_o, tl_o, [[a, b]], it, tenure)
            acc_d = engine.tabu_kick([[a, b]], it, tenure)
            assert acc_d == acc_o, (it, tries)
            if acc_o == 0:
                break
        s_dev, _ = engine.tour_download()
        assert (s_dev == succ_o).all(), it
        if it % 6 == 0:
            tenure = 15 if tenure == 3 else 3
    # several candidates in one call: the first acceptable one wins, the earlier ones leave their lazy-expiry traces
    cands = [[5, 5], [int(succ_o[9]), 9], [9, int(succ_o[9])], [g.rand_choice(0, n), g.rand_choice(0, n)], [3, 100], [7, 150]]
    acc_o, succ_o, tl_o = oracle.tabu_kick(succ_o, tl_o, cands, 36, tenure)
    assert engine.tabu_kick(cands, 36, tenure) == acc_o and acc_o >= 3
    s_dev, _ = engine.tour_download()
    assert (s_dev == succ_o).all()
    tl_d = engine.tabu_end(want_list=True)
    assert (tl_d == tl_o).all(), int((tl_d != tl_o).sum())
    engine.tour_restore(1)
    assert engine.tour_cost() == best


def test_resident_population_fitness_repair_and_offspring(engine, oracle):
    """GA (reference src/genetic.c): a population generated like random_generation() stays in HBM; fitness(), the 2-opt
    repair of selected individuals (alg_2opt on the chromosome's edges, :426-443) and the replacement of individuals by
    offspring only move indices, costs and the new chromosomes."""
    n, pop = 300, 48
    xy = uniform_instance(n)
    engine.set_instance(xy, 0)
    orders = reference_random_population(n, pop, 123)
    engine.population_upload(orders, as_order=True)
    fit = engine.population_costs(pop)
    assert fit.tolist() == [oracle.order_cost(xy, 0, o) for o in orders]
    sel = np.array([3, 17, 40, 0], dtype=np.int32)
    obj, st = engine.population_two_opt(FI, slots=sel, obj=fit[sel])
    back = engine.population_download(slots=sel, as_order=True)
    for k, slot in enumerate(sel):
        es, eobj, _, _ = oracle.two_opt_fi(xy, 0, order_to_succ(orders[slot]), fit[slot])
        assert obj[k] == eobj and (back[k] == succ_to_order(es)).all(), slot  # chromosome read off from node 0 (genetic.c:437-441)
    # untouched individuals are unchanged
    rest = engine.population_download(slots=[1, 2], as_order=False)
    assert (rest[0] == order_to_succ(orders[1])).all() and (rest[1] == order_to_succ(orders[2])).all()
    # offspring replace two individuals; only their fitness is recomputed
    kids = reference_random_population(n, 2, 999)
    engine.population_upload(kids, slots=[5, 9], as_order=True)
    assert engine.population_costs(slots=[5, 9]).tolist() == [oracle.order_cost(xy, 0, k) for k in kids]
    # best improvement over the whole population == the batched host-buffer call
    obj_all, _ = engine.population_two_opt(BI, count=pop)
    cur = orders.copy()
    cur[[5, 9]] = kids
    succs = np.stack([order_to_succ(o) for o in cur])
    for k, slot in enumerate(sel):
        succs[slot] = order_to_succ(back[k])
    sb, ob, _ = engine.two_opt_batch(BI, succs)
    assert (obj_all == ob).all() and (engine.population_download(pop, as_order=False) == sb).all()
    with pytest.raises(eng.TspB200Error):
        engine.population_costs(slots=[pop])
