"""The drop-in boundary, end to end: the UNMODIFIED reference objects (heuristics.c, tabusearch.c, ... compiled from
/root/reference/src by oracle/Makefile) linked against the product's calc_dist / alg_2opt / alg_2opt_tabu /
reverse_path (csrc/dropin.cpp -> libtspb200.so -> CUDA), exactly the INTEGRATION.md §2 recipe.  The reference's own
drivers (HEU_2opt_greedy = `-method 2OPT_GREEDY`, src/heuristics.c:572; HEU_2opt_extramileage :596;
HEU_2opt_greedy_iter :584) must produce the same tour and objective as the pure reference build and as the
reference's published CSV goldens."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpulib():
    from oracle.oracle import REF_GPU_SO, RefLib
    if not os.path.exists(REF_GPU_SO):
        pytest.skip("oracle/_ref/libtspref_gpu.so not built (needs /root/reference at build time)")
    return RefLib(gpu_link=True)


def test_strong_symbols_come_from_the_dropin(gpulib):
    # tspb200_dropin_layout only exists in csrc/dropin.cpp: its presence next to the reference's HEU_* drivers shows
    # which definitions the link picked
    assert hasattr(gpulib.L, "tspb200_dropin_layout") and hasattr(gpulib.L, "HEU_2opt_greedy")
    # the drop-in's definitions are laid out contiguously (one object file): the weakened reference symbols greedy,
    # HEU_Greedy_iter and HEU_extramileage must resolve into that range, HEU_2opt_greedy (reference code) must not
    import ctypes as C
    addr = lambda nm: C.cast(getattr(gpulib.L, nm), C.c_void_p).value
    lo = min(addr(nm) for nm in ("alg_2opt", "calc_dist", "reverse_path", "tspb200_dropin_layout"))
    hi = max(addr(nm) for nm in ("alg_2opt", "calc_dist", "reverse_path", "tspb200_dropin_layout", "tspb200_dropin_reset"))
    for nm in ("greedy", "HEU_Greedy_iter", "HEU_extramileage", "alg_2opt_tabu"):
        assert lo - 4096 <= addr(nm) <= hi + 4096, nm
    assert not (lo - 4096 <= addr("HEU_2opt_greedy") <= hi + 4096)


@pytest.mark.parametrize("nm", ["berlin52", "pr299", "att532", "gr666", "dsj1000", "pr1002", "ulysses22"])
def test_reference_2opt_greedy_on_cuda_path(gpulib, reflib, instances, goldens, nm):
    xy, wt = instances[nm]
    st_g, succ_g, obj_g = gpulib.run_method("HEU_2opt_greedy", xy, wt)
    st_r, succ_r, obj_r = reflib.run_method("HEU_2opt_greedy", xy, wt)
    assert st_g == st_r == 0
    assert (succ_g == succ_r).all() and obj_g == obj_r
    csv = goldens["reference_csv"].get(nm)  # results/constructive_heuristics_2opt_new.csv, column 2OPT_GREEDY
    if csv:
        assert obj_g == csv["2OPT_GREEDY"]


@pytest.mark.parametrize("method,nm", [("HEU_2opt_extramileage", "lin318"), ("HEU_2opt_extramileage", "att532"),
                                       ("HEU_2opt_greedy_iter", "berlin52"), ("HEU_greedy", "gr666"),
                                       ("HEU_extramileage", "pr299"), ("HEU_2opt_greedy_iter", "gr431"),
                                       ("HEU_Greedy_iter", "dsj1000"), ("HEU_extramileage", "gr666"),
                                       ("HEU_2opt_extramileage", "dsj1000")])
def test_other_reference_drivers_on_cuda_path(gpulib, reflib, instances, method, nm):
    xy, wt = instances[nm]
    st_g, succ_g, obj_g = gpulib.run_method(method, xy, wt)
    st_r, succ_r, obj_r = reflib.run_method(method, xy, wt)
    assert st_g == st_r
    assert (succ_g == succ_r).all() and obj_g == obj_r


def test_reference_plain_bi_through_dropin(gpulib, reflib, instances):
    """alg_2opt_tabu(inst, NULL, prev, 1, 1) incl. the exported prev[] (src/tabusearch.c:173-175)."""
    xy, wt = instances["pr299"]
    succ0, _ = reflib.nn_tour(xy, wt, 0)
    s_g, o_g, p_g = gpulib.two_opt_bi(xy, wt, succ0, want_prev=True)
    s_r, o_r, p_r = reflib.two_opt_bi(xy, wt, succ0, want_prev=True)
    assert (s_g == s_r).all() and o_g == o_r and (p_g == p_r).all()


def test_reference_masked_bi_through_dropin(gpulib, reflib, instances):
    """alg_2opt_tabu WITH a tabu list through the reference-named symbol: same tour, cost and mutated list."""
    xy, wt = instances["berlin52"]
    n = len(xy)
    rng = np.random.default_rng(3)
    succ0, _ = reflib.nn_tour(xy, wt, 0)
    ncols = n * (n - 1) // 2
    mask = np.where(rng.random(ncols) < 0.15, rng.integers(1, 40, size=ncols), 0).astype(np.int32)
    m_g, m_r = mask.copy(), mask.copy()
    s_g, o_g = gpulib.two_opt_bi(xy, wt, succ0, skip_edge=m_g, iter_=40, tenure=9)
    s_r, o_r = reflib.two_opt_bi(xy, wt, succ0, skip_edge=m_r, iter_=40, tenure=9)
    assert (s_g == s_r).all() and o_g == o_r and (m_g == m_r).all()
