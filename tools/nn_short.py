"""Short fixed run for ncu (dev tool): one nearest-neighbour tour on the bucket grid.  python tools/nn_short.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
eng = Engine(0)
eng.set_instance(uniform_instance(n), 0)
eng.set_option("nn_grid", 1)
print("nn", eng.nn_tour(0)[1])
eng.close()
