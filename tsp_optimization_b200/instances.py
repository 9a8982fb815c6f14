"""Instance helpers for tests and benchmarks: the synthetic generator of SURVEY.md §8(c), a small TSPLIB
NODE_COORD_SECTION reader that mirrors the reference parser's tokenisation (reference
src/utility.c:351-453: separators " :\\n\\t\\r", atoi/atof, 1-based ids), and tour conversions."""
from __future__ import annotations

import random

import numpy as np

WEIGHT_TYPES = {"EUC_2D": 0, "MAX_2D": 1, "MAN_2D": 2, "CEIL_2D": 3, "GEO": 4, "ATT": 5}


def uniform_instance(n: int, seed: int | None = None, hi: int = 10000) -> np.ndarray:
    """uni<n>: rng = random.Random(n); x then y drawn with randint(0, hi) per node; EUC_2D.
    The range mirrors the reference's generator cap (other_codes/generate_tsp_istances.py:3-4,26-27)."""
    rng = random.Random(n if seed is None else seed)
    xy = np.empty((n, 2), dtype=np.float64)
    for i in range(n):
        xy[i, 0] = rng.randint(0, hi)
        xy[i, 1] = rng.randint(0, hi)
    return xy


def read_tsplib(path: str):
    """Returns (xy, weight_type). Unknown / missing EDGE_WEIGHT_TYPE -> -1 (the reference then falls back
    to EUC_2D in calc_dist, src/distutil.c:91)."""
    import re
    n, wt, xy, section = -1, -1, None, 0
    with open(path) as f:
        for line in f:
            if len(line) <= 1:
                continue
            tok = [t for t in re.split(r"[ :\n\t\r]+", line) if t]
            if not tok:
                continue
            key = tok[0]
            if key.startswith("NAME") or key.startswith("COMMENT") or key.startswith("TYPE"):
                section = 0
            elif key.startswith("DIMENSION"):
                n = int(tok[1]); xy = np.zeros((n, 2), dtype=np.float64); section = 0
            elif key.startswith("EOF"):
                break
            elif key.startswith("EDGE_WEIGHT_TYPE"):
                for name, val in WEIGHT_TYPES.items():
                    if tok[1].startswith(name):
                        wt = val
                if tok[1].startswith("EXPLICIT"):
                    raise ValueError("EXPLICIT instances are rejected by the reference parser (utility.c:418)")
                section = 0
            elif key.startswith("NODE_COORD_SECTION"):
                section = 1
            elif key.startswith("EDGE_WEIGHT_SECTION"):
                section = 2
            elif section == 1:
                i = int(key) - 1
                xy[i] = (float(tok[1]), float(tok[2]))
    return xy, wt


def order_to_succ(order) -> np.ndarray:
    order = np.asarray(order, dtype=np.int32)
    succ = np.empty_like(order)
    succ[order] = np.roll(order, -1)
    return succ


def succ_to_order(succ, start: int = 0) -> np.ndarray:
    succ = np.asarray(succ)
    out = np.empty(len(succ), dtype=np.int32)
    at = start
    for p in range(len(succ)):
        out[p] = at
        at = succ[at]
    return out


def random_tours(n: int, batch: int, seed: int) -> np.ndarray:
    """`batch` random Hamiltonian cycles as successor arrays (GA-style random population)."""
    rng = np.random.default_rng(seed)
    out = np.empty((batch, n), dtype=np.int32)
    for b in range(batch):
        out[b] = order_to_succ(rng.permutation(n).astype(np.int32))
    return out


def is_tour(succ) -> bool:
    succ = np.asarray(succ)
    n = len(succ)
    seen = np.zeros(n, dtype=bool)
    at = 0
    for _ in range(n):
        if at < 0 or at >= n or seen[at]:
            return False
        seen[at] = True
        at = succ[at]
    return at == 0
