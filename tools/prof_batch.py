"""Short, fixed run for ncu (dev tool): one wave of the position-space one-block BI kernel (GA-style population, uni1000),
then the exact tile pruning path (boxes / filter / pruned scan) at n = 100 000.  python tools/prof_batch.py [tours] [passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tsp_optimization_b200 import BI, Engine  # noqa: E402
from tsp_optimization_b200.instances import order_to_succ, reference_random_population, uniform_instance  # noqa: E402

tours = int(sys.argv[1]) if len(sys.argv) > 1 else 592
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 40
eng = Engine(0)
xy = uniform_instance(1000)
eng.set_instance(xy, 0)
pop = reference_random_population(1000, tours, 123)
succ = np.stack([order_to_succ(o) for o in pop])
sb, ob, st = eng.two_opt_batch(BI, succ)
print("batch bi", st)
n = 100000
eng.set_instance(uniform_instance(n), 0)
s0 = np.load("tests/golden/nn_uni100000.npz")["succ"]
eng.set_option("prune", 1)
eng.tour_upload(s0)
print("pruned", eng.bi_run(passes))
