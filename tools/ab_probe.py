"""A/B timing of library builds of the same ABI (TSPB200_LIB): best-improvement passes at n = 100 000 (exhaustive), one
process per library.  python tools/ab_probe.py <lib.so> [n] [passes]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1]
if lib != "default":
    os.environ["TSPB200_LIB"] = os.path.abspath(lib)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 100

from tsp_optimization_b200 import Engine  # noqa: E402
from tsp_optimization_b200.instances import uniform_instance  # noqa: E402

eng = Engine(0)
try:
    eng.set_option("prune", 0)
except Exception:
    pass
eng.set_instance(uniform_instance(n), 0)
succ, _ = eng.nn_tour(0)
out = {"lib": os.path.basename(lib), "n": n}
for world in (1, 8):
    try:
        eng.set_option("debug_shard", (world << 8) | (world - 1) if world > 1 else 0)
    except Exception:
        continue
    best = 1e9
    for rep in range(3):
        eng.tour_upload(succ)
        eng.bi_run(5)
        st = eng.bi_run(passes)
        best = min(best, 1e3 * st.gpu_ms / st.passes)
    out[f"us_per_pass_w{world}"] = round(best, 2)
    out[f"shape_w{world}"] = [eng.info("block_threads"), eng.info("rows_per_thread"), eng.info("tile_cols")]
print(json.dumps(out), flush=True)
eng.close()
