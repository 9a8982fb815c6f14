import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_optimization_b200 import Engine
from tsp_optimization_b200.instances import uniform_instance
n = int(sys.argv[1]); seg = int(sys.argv[2])
eng = Engine(0)
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
for pdl in (0, 1):
    eng.set_option("pdl", pdl)
    eng.tour_upload(succ)
    out = []; prev = 0; tot = 0.0
    while True:
        st = eng.bi_run(seg)
        c = eng.info("cold_calls")
        out.append(f"{st.gpu_ms*1e3/max(1,st.passes):.0f}/{(c-prev)//max(1,st.passes)}")
        tot += st.gpu_ms
        prev = c
        if st.status == 0: break
    print("pdl", pdl, "total_ms", tot, " ".join(out), flush=True)
