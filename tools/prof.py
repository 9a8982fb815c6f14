"""Short, fixed run for ncu (dev tool): uni<n> NN start, a few BI passes, one matrix build, a few FI moves."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys
import numpy as np
from tsp_optimization_b200 import Engine, BI, FI
from tsp_optimization_b200.instances import uniform_instance
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 6
eng = Engine(0)
eng.set_option("prune", 0 if "exhaustive" in sys.argv else -1)  # the headline (roofline) kernel is the exhaustive scan
xy = uniform_instance(n)
eng.set_instance(xy, 0)
z = np.load("tests/golden/nn_uni100000.npz") if n == 100000 else None
succ = z["succ"] if z is not None else eng.nn_tour(0)[0]
eng.tour_upload(succ)
st = eng.bi_run(passes)
print("bi", st)
if len(sys.argv) > 3 and sys.argv[3] == "matrix":
    m = uniform_instance(20000)
    eng.set_instance(m, 0)
    print("matrix ms", eng.dist_matrix_build())
