"""Short first-improvement run for ncu launch lists (dev tool): uni100000 NN start, <moves> FI moves.  python tools/fi_short.py [moves]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 100
eng = Engine(0)
eng.set_instance(uniform_instance(100000), 0)
succ = np.load("tests/golden/nn_uni100000.npz")["succ"]
eng.tour_upload(succ)
print("fi", eng.fi_run(moves))
eng.close()
