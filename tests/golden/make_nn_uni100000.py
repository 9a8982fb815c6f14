"""Generates tests/golden/nn_uni100000.npz: the nearest-neighbour start tour (reference greedy(), restated in
oracle/tsp_oracle.c:orc_nn_tour) of the headline benchmark instance uni100000.  ~2 minutes of CPU; committed so
that the CPU reference arm of bench.py and the full-size GPU NN parity test do not have to redo 10^10 calc_dist calls."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402
from tsp_optimization_b200.instances import uniform_instance  # noqa: E402

n = 100000
xy = uniform_instance(n)
t = time.time()
succ, cost = Oracle().nn_tour(xy, 0, 0)
print("NN cost", cost, "in", time.time() - t, "s")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nn_uni100000.npz"), succ=succ.astype(np.int32), cost=np.float64(cost))
