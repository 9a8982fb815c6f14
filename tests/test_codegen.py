"""Codegen canaries on the built objects (cuobjdump works without a GPU): properties of the SASS that the measured
performance depends on and that an innocent source change can silently lose."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "tsp_optimization_b200", "lib")


def _obj_of(shape: str) -> str:
    """the scan kernel's instantiations are spread over four objects (csrc/kernels_bi_scan.cuh)"""
    if shape.startswith("ILi128E"):
        return os.path.join(LIB, "kernels_bi_128.o")
    if shape.startswith("ILi256E"):
        return os.path.join(LIB, "kernels_bi_256.o")
    return os.path.join(LIB, "kernels_bi_s64.o" if shape.endswith("ELb1") else "kernels_bi_p64.o")


def _sass(mangled: str, obj: str) -> str:
    if not shutil.which("cuobjdump") or not os.path.exists(obj):
        pytest.skip("cuobjdump or the built object is not available")
    r = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, obj], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Function" in r.stdout, r.stderr[-500:]
    return r.stdout


@pytest.mark.parametrize("shape", ["ILi64ELi8ELb0ELb1ELb0ELb0", "ILi64ELi8ELb0ELb0ELb0ELb0", "ILi64ELi8ELb1ELb1ELb0ELb0", "ILi128ELi8ELb0ELb1ELb0ELb0",
                                   "ILi128ELi16ELb0ELb1ELb0ELb0", "ILi64ELi8ELb0ELb1ELb0ELb1", "ILi64ELi8ELb0ELb0ELb0ELb1", "ILi64ELi8ELb1ELb1ELb0ELb1"])
def test_exhaustive_scan_hot_loop_stays_in_the_uniform_datapath(shape):
    """bi_scan_kernel<T, R, ATT, EXACT32, PRUNED=false, SHUF>: the column records are read with LDS.128 [UR + imm] (uniform
    address register), the packed FP32x2 pipe and MUFU.SQRT are in use, and the TMA bulk copy is there.  With vector
    addressing (LDS.128 [R + imm] only) the same kernel measured 5 % slower on a B200 (1470 vs 1389 us per pass at
    n = 100 000): see the s_tile comment in csrc/kernels_bi_scan.cuh."""
    sass = _sass(f"_ZN4tspb14bi_scan_kernel{shape}EEEvNS_6BiArgsE", _obj_of(shape))
    assert len(re.findall(r"LDS\.128 R\d+, \[UR", sass)) >= 4
    assert "MUFU.SQRT" in sass and "FFMA2" in sass and "FADD2" in sass and "UBLKCP" in sass
    if shape.endswith("ELb1"):  # the shuffle variant: R square roots per column (no scalar (R+1)-th one), one SHFL.DOWN instead
        assert "SHFL.DOWN" in sass
