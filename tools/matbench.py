"""Dev tool: distance-matrix kernel timing (CUDA events inside tspb200_dist_matrix_build) vs the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
eng = Engine(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for wt in (0, 3, 5):
    for n in (10000, 20000, 32768, 65536):
        if wt != 0 and n != 32768:
            continue
        eng.set_instance(uniform_instance(n), wt)
        eng.dist_matrix_build()
        ms = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            ms.append(eng.dist_matrix_build())
        ld = eng.info("matrix_ld")
        eng.dist_matrix_free()
        m = float(np.median(ms))
        gbs = 4.0 * n * ld / (m * 1e-3) / 1e9
        print(json.dumps({"wt": wt, "n": n, "ms": m, "GBs": gbs, "frac_of_measured_hbm": gbs / peak}), flush=True)
