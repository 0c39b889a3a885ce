"""CPU-side checks of the drop-in boundary: the shared library builds, loads, and exports every symbol
include/swb200.h and include/algoGPU.h declare (no compute calls: there is no GPU here)."""
import ctypes as C
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")


@pytest.fixture(scope="module")
def lib():
    from concurrentproject_b200 import _lib
    _lib.build()
    return _lib.load()


def declared_symbols():
    names = set()
    for hdr in ("swb200.h", "algoGPU.h"):
        txt = (ROOT / "include" / hdr).read_text()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"SWB200_API\s+[\w\s\*]+?\b(\w+)\s*\(", txt))
    return names


def test_every_declared_symbol_is_exported(lib):
    from concurrentproject_b200 import _lib
    decl = declared_symbols()
    assert {"SequentialSmithWatermanScoreGPU", "SmithWatermanLazyGPU", "SmithWatermanScoreCUDA", "SmithDiagonalGPU",
            "swb200_score", "swb200_score_device"} <= decl
    assert decl == set(_lib.EXPORTS), decl ^ set(_lib.EXPORTS)
    for name in decl:
        assert hasattr(lib, name), name


def test_library_is_sm100a_only_and_uses_dpx():
    so = ROOT / "concurrentproject_b200" / "lib" / "libswb200.so"
    out = subprocess.run(["cuobjdump", "-lelf", str(so)], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    for fun in ("_ZN3swb16sw_engine_kernelILi4ELi0ELi1ELi4ELi0ELb0EEEvNS_12EngineLaunchE",      # pair engine, affine, 4 rows
                "_ZN3swb15sw_chain_kernelILi3ELi1EEEvNS_11ChainLaunchE"):                       # CTA-chained engine, linear, 3 rows (cfg2)
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, str(so)], capture_output=True, text=True).stdout
        for mnemonic in ("VIADDMNMX.S16x2", "VIMNMX.S16x2", "VIADD.16x2", "PRMT", "SHFL.IDX"):
            assert mnemonic in sass, (fun, mnemonic)


def test_argument_errors_need_no_gpu(lib):
    from concurrentproject_b200 import _lib
    out = C.c_int(-7)
    # empty input scores 0 without touching the device (both reference oracles return 0)
    assert lib.swb200_score(None, 0, None, 0, None, C.byref(out)) == 0 and out.value == 0
    p = _lib.Params(1, 1, 1, 1)  # positive mismatch: outside the documented limits
    assert lib.swb200_score(None, 0, None, 0, C.byref(p), C.byref(out)) == -2
    assert b"limits" in lib.swb200_last_error()


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from concurrentproject_b200 import api
    with pytest.raises(api.SwbError):
        api.score(b"ACGT", b"ACGT")
