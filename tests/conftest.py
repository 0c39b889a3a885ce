import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device (or without the built library) the `gpu` tests are skipped, not run: the legacy entry
    points abort() on any error (the reference signature has no error channel), which would take pytest down."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        from concurrentproject_b200 import _lib
        have = _lib.load().swb200_device_count()
        why = "no CUDA device"
    except Exception as e:          # library not built
        have, why = 0, f"libswb200.so not loadable: {e}"
    if have == 0:
        skip = pytest.mark.skip(reason=why)
        for it in gpu_items:
            it.add_marker(skip)


def load_json(name):
    return json.loads((GOLDEN / name).read_text())


def fixture_pair(case):
    """Rebuild the byte sequences of a ref_scores_*.json case from its (seed, stream) recipe."""
    from concurrentproject_b200 import rng
    a = rng.random_acgt(case["seed"], case["stream1"], case["n"])
    if case.get("kind") == "planted" or case.get("planted"):
        b = rng.mutate(a, case["seed"], case["mut_stream"], case.get("sub", 0.08), case.get("indel", 0.04))
    else:
        b = rng.random_acgt(case["seed"], case["stream2"], case["m"])
    assert len(b) == case["m"]
    return a, b


def mt_pairs(L):
    rows = []
    for line in (GOLDEN / f"mt12345_L{L}.txt").read_text().split("\n"):
        if line:
            a, b, s = line.split()
            rows.append((a.encode(), b.encode(), int(s)))
    return rows
