"""tsp_optimization_b200 — B200 (sm_100a) engine for the TSP_Optimization distance / 2-opt hot path.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/tspb200.h), engine.py (ctypes
mirror of that ABI), dist.py (torch.distributed bootstrap for the multi-GPU paths) and instances.py.
"""
from .engine import (ATT, BI, CEIL_2D, EUC_2D, FI, GEO, MAN_2D, MAX_2D, Engine, Stats, TspB200Error,
                     load_library)

__all__ = ["Engine", "Stats", "TspB200Error", "load_library", "FI", "BI", "EUC_2D", "MAX_2D", "MAN_2D",
           "CEIL_2D", "GEO", "ATT"]
