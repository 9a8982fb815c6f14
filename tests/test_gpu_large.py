"""GPU parity on four bigger reference instances with numeric regimes of their own (tests/golden/make_goldens_large.py):
decimal coordinates that FP32 cannot hold (fl3795, usa13509, stefano_8k) and CEIL_2D around 10^6 (pla7397).  Goldens were
produced by the oracle and cross-checked against the compiled reference when the fixture was made."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLD_DIR
from tsp_optimization_b200 import engine as eng

pytestmark = pytest.mark.gpu
FI, BI = eng.FI, eng.BI


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def large():
    z = np.load(os.path.join(GOLD_DIR, "large.npz"))
    with open(os.path.join(GOLD_DIR, "goldens_large.json")) as f:
        g = json.load(f)["instances"]
    return {nm: (z[nm + "__xy"], int(z[nm + "__wt"]), g[nm]) for nm in g}


@pytest.mark.parametrize("nm", ["fl3795", "pla7397", "stefano_8k", "usa13509"])
def test_large_instance_matrix_nn_fi_bi(engine, large, nm):
    xy, wt, g = large[nm]
    engine.set_instance(xy, wt)
    expect_exact32 = 1 if nm == "pla7397" else 0
    assert engine.info("exact32") == expect_exact32 and engine.info("fp32_ok") == 1
    m = engine.dist_matrix()
    assert int(m.astype(np.int64).sum()) == g["matrix_sum"] and sha(m) == g["matrix_sha256"]
    del m
    engine.dist_matrix_free()
    succ, cost = engine.nn_tour(0)
    assert cost == g["nn_cost"] and sha(succ) == g["nn_sha256"]
    for route in ((0, 1) if len(xy) <= 4096 else (0,)):
        engine.set_option("single_block", route)
        s, obj, st, _ = engine.two_opt(FI, succ, cost)
        assert obj == g["fi_cost"] and sha(s) == g["fi_sha256"] and st.moves == g["fi_moves"] and st.passes == g["fi_sweeps"], route
    engine.set_option("single_block", -1)
    s, obj, st, log = engine.two_opt(BI, succ, 0.0, max_iters=g["bi_passes"], log_cap=g["bi_passes"] + 4)
    assert log.tolist() == g["bi_log"] and obj == g["bi_cost_after"] and sha(s) == g["bi_sha256_after"]
