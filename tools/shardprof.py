import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
eng = Engine(0)
n = 100000
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
eng.set_option("debug_shard", (8 << 8) | 0)
eng.tour_upload(succ)
st = eng.bi_run(12)
print(st.gpu_ms * 1e3 / st.passes)
