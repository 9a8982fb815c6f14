"""tests/golden/large.npz + goldens_large.json: four bigger reference instances that exercise the numeric regimes the
18-instance benchmark set does not — decimal coordinates that are not FP32-representable (fl3795, usa13509, art/stefano_8k:
the all-FP64 matrix kernel and the FP32 filter with a coordinate-rounding term) and CEIL_2D with coordinates around 10^6
(pla7397).  Coordinates come from the reference's own TSPLIB parser (compiled reference, oracle/_ref); goldens from the
oracle restatement (pinned to the compiled reference by tests/test_oracle.py), cross-checked here against the compiled
reference for the NN tour and the first-improvement result."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, RefLib  # noqa: E402

DATA = "/root/reference/data"
FILES = {"fl3795": "all/fl3795.tsp", "pla7397": "all/pla7397.tsp", "usa13509": "all/usa13509.tsp", "stefano_8k": "art/stefano_8k.tsp"}
BI_PASSES = 24


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


orc, ref = Oracle(), RefLib()
arrays, gold = {}, {}
for nm, rel in FILES.items():
    t0 = time.time()
    xy, wt = ref.parse("./" + os.path.relpath(os.path.join(DATA, rel)))  # "./" side-steps the unterminated path copy
    n = len(xy)
    arrays[nm + "__xy"] = xy
    arrays[nm + "__wt"] = np.int32(wt)
    g = {"n": n, "wt": int(wt)}
    m = orc.dist_matrix(xy, wt)
    g["matrix_sha256"], g["matrix_sum"] = sha(m), int(m.astype(np.int64).sum())
    del m
    succ, cost = orc.nn_tour(xy, wt, 0)
    rsucc, rcost = ref.nn_tour(xy, wt, 0)
    assert (succ == rsucc).all() and cost == rcost
    g["nn_cost"], g["nn_sha256"] = cost, sha(succ)
    fs, fobj, fst, _ = orc.two_opt_fi(xy, wt, succ, cost)
    rfs, rfobj = ref.two_opt_fi(xy, wt, succ, cost)
    assert (fs == rfs).all() and fobj == rfobj
    g.update(fi_cost=fobj, fi_sha256=sha(fs), fi_moves=int(fst.moves), fi_sweeps=int(fst.passes))
    bs, bobj, bst, blog = orc.two_opt_bi(xy, wt, succ, max_passes=BI_PASSES, log_cap=BI_PASSES + 4)
    g.update(bi_passes=BI_PASSES, bi_cost_after=bobj, bi_sha256_after=sha(bs), bi_log=blog.tolist())
    gold[nm] = g
    print(nm, n, wt, {k: v for k, v in g.items() if k != "bi_log"}, f"{time.time() - t0:.0f} s", flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "large.npz"), **arrays)
json.dump({"instances": gold, "source": "tests/golden/make_goldens_large.py (oracle + compiled reference, /root/reference/data)"},
          open(os.path.join(ROOT, "tests", "golden", "goldens_large.json"), "w"), indent=1, sort_keys=True)
