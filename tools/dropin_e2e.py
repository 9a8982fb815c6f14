"""The reference's own drivers, unmodified, timed twice on the same box: linked against its own CPU functions
(oracle/_ref/libtspref.so) and against the drop-in (oracle/_ref/libtspref_gpu.so -> libtspb200.so -> B200)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import RefLib
from tsp_optimization_b200.instances import uniform_instance

cpu, gpu = RefLib(), RefLib(gpu_link=True)
z = np.load("tests/golden/instances.npz")
cases = [(nm, z[nm + "__xy"], int(z[nm + "__wt"])) for nm in ("pr299", "att532", "gr666", "pr1002", "dsj1000")]
cases.append(("uni4000", uniform_instance(4000), 0))
gpu.run_method("HEU_2opt_greedy", cases[0][1], cases[0][2])  # context creation / module load outside the timings
for nm, xy, wt in cases:
    for method in ("HEU_2opt_greedy", "HEU_2opt_extramileage", "HEU_2opt_greedy_iter"):
        if method != "HEU_2opt_greedy" and len(xy) > 1100:
            continue  # O(n^3) on the CPU side
        t0 = time.perf_counter(); sg = gpu.run_method(method, xy, wt); tg = time.perf_counter() - t0
        t0 = time.perf_counter(); sc = cpu.run_method(method, xy, wt); tc = time.perf_counter() - t0
        same = bool((sg[1] == sc[1]).all() and sg[2] == sc[2])
        print(json.dumps({"instance": nm, "n": len(xy), "method": method, "objective": sc[2], "reference_cpu_s": tc, "dropin_b200_s": tg,
                          "speedup": tc / tg, "identical_tour_and_cost": same}), flush=True)
