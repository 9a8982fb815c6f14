"""Every repository path that DESIGN.md / INTEGRATION.md / README.md / profiles/README.md / tools/README.md cite exists."""
import os
import re

import pytest

from conftest import ROOT

DOCS = ["DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md"), os.path.join("tools", "README.md")]
PREFIXES = ("profiles/", "tools/", "tests/", "include/", "examples/", "oracle/", "tsp_optimization_b200/", "csrc/")
BUILT = ("oracle/_ref/", "oracle/_build/", "tsp_optimization_b200/lib/")  # build products, not tracked
REFERENCE_FILES = {"utility.c", "plot.c", "cplex.h", "distutil.c", "heuristics.c", "tabusearch.c", "vns.c", "genetic.c", "solver.c",
                   "callback.c", "main.c", "convexhull.c", "utility.h", "distutil.h", "heuristics.h", "tabusearch.h"}  # deno750/TSP_Optimization


@pytest.mark.parametrize("doc", DOCS)
def test_cited_paths_exist(doc):
    base = os.path.dirname(os.path.join(ROOT, doc))
    txt = open(os.path.join(ROOT, doc)).read()
    missing = []
    for tok in set(re.findall(r"`([A-Za-z0-9_./\-]+)`", txt)):
        tok = tok.split("::")[0]
        if tok in REFERENCE_FILES:
            continue
        cands = [tok]
        if tok.startswith("csrc/"):
            cands = ["tsp_optimization_b200/" + tok]
        elif "/" not in tok and re.search(r"\.(py|cu|cuh|cpp|h|md|json|jsonl|txt|c|npz)$", tok):
            cands = [tok, os.path.relpath(os.path.join(base, tok), ROOT)] + [os.path.join(d, tok) for d in
                     ("tools", "tests", "profiles", "include", "oracle", "tests/golden", "tsp_optimization_b200", "tsp_optimization_b200/csrc", "examples")]
        elif not tok.startswith(PREFIXES):
            continue
        if tok.startswith(BUILT) or "*" in tok or tok.endswith("/") and os.path.isdir(os.path.join(ROOT, tok)):
            continue
        if not any(os.path.exists(os.path.join(ROOT, c)) for c in cands):
            missing.append(tok)
    assert not missing, (doc, sorted(missing))
