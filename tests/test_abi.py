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


def test_host_packers_write_the_resident_format(lib):
    """swb200_pack_batch_host / swb200_pack_banded_host are format conversion on the host (no GPU): symbol k of a sequence
    lands at bits 2*(k%32) of word k/32 as (c >> 1) & 3; the batch packer puts the shorter sequence first, the banded
    packer keeps seq1 / seq2; a byte outside A,C,G,T is an error."""
    import numpy as np
    from concurrentproject_b200 import api, rng

    def decode(words, stride, k, n):
        w = words[k * stride:(k + 1) * stride]
        return np.array([(int(w[i >> 5]) >> (2 * (i & 31))) & 3 for i in range(n)], dtype=np.uint8)

    r = np.random.default_rng(3)
    s1 = [bytes(rng.random_acgt(3, k, int(r.integers(0, 200)))) for k in range(9)] + [b"", b"ACGT" * 16]
    s2 = [bytes(rng.random_acgt(3, 100 + k, int(r.integers(0, 200)))) for k in range(9)] + [b"ACG", b""]
    f1, o1, l1 = api._flatten(s1); f2, o2, l2 = api._flatten(s2)
    code = lambda b: (np.frombuffer(b, dtype=np.uint8) >> 1) & 3

    w1, st1, w2, st2 = api.pack_banded_host(f1, o1, l1, f2, o2, l2)
    assert st1 == (max(map(len, s1)) + 31) // 32 + 2 and st2 == (max(map(len, s2)) + 31) // 32 + 2
    for k, (a, b) in enumerate(zip(s1, s2)):
        assert np.array_equal(decode(w1, st1, k, len(a)), code(a)) and np.array_equal(decode(w2, st2, k, len(b)), code(b))
        assert not w1[k * st1 + (len(a) + 31) // 32:(k + 1) * st1].any()          # words beyond the end are zero

    qw, qs, tw, ts, ql, tl = api.pack_batch_host(f1, o1, l1, f2, o2, l2)
    for k, (a, b) in enumerate(zip(s1, s2)):
        short, long_ = (a, b) if len(a) <= len(b) else (b, a)
        assert (ql[k], tl[k]) == (len(short), len(long_))
        assert np.array_equal(decode(qw, qs, k, len(short)), code(short)) and np.array_equal(decode(tw, ts, k, len(long_)), code(long_))

    bad = np.frombuffer(b"ACGN", dtype=np.uint8)
    ok = np.frombuffer(b"ACGT", dtype=np.uint8)
    z, four = np.zeros(1, np.int64), np.array([4], np.int32)
    with pytest.raises(api.SwbError):
        api.pack_banded_host(ok, z, four, bad, z, four)
    with pytest.raises(api.SwbError):
        api.pack_batch_host(bad, z, four, ok, z, four)


def test_packed_batch_calls_validate_their_arguments_before_touching_a_gpu(lib):
    """Lengths that break the packed layout are refused by the branch-free length scan, on any machine: a q longer than its
    t, a negative length, a pair longer than the strides allow; the banded call also wants strides of at least 3 words."""
    import numpy as np
    from concurrentproject_b200 import api
    qs, ts = api.batch_strides(40, 90)
    qw, tw = np.zeros(3 * qs, dtype=np.uint64), np.zeros(3 * ts, dtype=np.uint64)
    i32 = lambda *v: np.array(v, dtype=np.int32)
    for ql, tl in ((i32(10, 50, 10), i32(20, 40, 30)),            # q longer than t
                   (i32(10, -1, 10), i32(20, 40, 30)),            # negative
                   (i32(10, 20, 41 + 32), i32(20, 40, 90)),       # q beyond its stride
                   (i32(10, 20, 30), i32(20, 40, 90 + 64))):      # t beyond its stride
        with pytest.raises(api.SwbError) as e:
            api.score_batch_packed(qw, qs, tw, ts, ql, tl)
        assert e.value.code == -2, (ql, tl)                         # SWB200_ERR_ARG
    w = np.zeros(12, dtype=np.uint64)
    with pytest.raises(api.SwbError) as e:
        api.score_banded_batch_packed(w, 2, w, 4, i32(10, 10, 10), i32(10, 10, 10))
    assert e.value.code == -2
    with pytest.raises(api.SwbError) as e:
        api.score_banded_batch_packed(w, 4, w, 4, i32(10, 70, 10), i32(10, 10, 10))      # 70 symbols need 3 + 2 words
    assert e.value.code == -2
