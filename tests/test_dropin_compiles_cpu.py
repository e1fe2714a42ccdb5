"""CPU test (build container only, where /root/reference exists): the reference's UNMODIFIED main.cpp, with its own
inputParam.hpp / solverSetup.hpp edited only in their constants, compiles and links against include/reference_compat and
libpps_b200.so for every solver stack the adapters claim to accept (level 1).  Compile + link only: running
needs a GPU (tests/test_gpu_driver.py)."""
import os

import pytest

from oracle import build_ref as br

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parallelpoissonsolver_b200", "csrc", "libpps_b200.so")

pytestmark = pytest.mark.skipif(not br.available() or not os.path.exists(LIB),
                                reason="needs /root/reference (build container) and the built library")

# config -> what it exercises in the adapters
STACKS = {
    "nb24": "BiCGSTAB + nested local BiCGSTAB (T_Preconditioner, inputParam.hpp:31)",
    "nbg24_i8": "BiCGSTAB + nested GLOBAL BiCGSTAB (BiCGSTAB<.., false, communicationON, NoneSolver>; alpaka T_PreconditionerBiCGStabGlobal)",
    "nc24": "BiCGSTAB + nested local CG with Chebyshev inside (T_Preconditioner3, inputParam.hpp:29)",
    "chm24": "ChebyshevIteration as T_Solver (isMainLoop, communicationON)",
    "o1cgm24": "BaseCG, orderNeumanBcs = 1",
    "q24_cheb": "DIM = 2, BiCGSTAB + Chebyshev",
    "l48": "DIM = 1, BiCGSTAB",
    "m24_chebg": "BiCGSTAB + ChebyshevIteration<.., communicationON, ..> in the preconditioner slot (global polynomial preconditioner)",
}


@pytest.mark.parametrize("name", sorted(STACKS))
def test_unmodified_reference_main_builds_against_the_adapters(name):
    exe = br.build_dropin(name, force=True)
    assert exe and os.path.exists(exe) and os.access(exe, os.X_OK), STACKS[name]
