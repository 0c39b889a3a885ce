#!/usr/bin/env python
"""Pins the scores of the pairs the reference itself cannot reach in reasonable time:
the 1 000 000 x 1 000 000 pair (seed 6) and BASELINE config 3 (4 000 000 x 4 000 000, seed 3; ~46 h for the
reference's LazySmith).  Computed on the CPU by oracle/gotoh_fast.c -- the restated main.cpp recurrence with one
tile per SIMD lane, proven equal to oracle_gotoh_rolling (and through it to the unmodified reference) by
tests/test_oracle.py -- and written to tests/golden/large_scores.json.

  python tests/golden/make_golden_large.py [name ...]      (default: all; cfg3 takes ~30-45 min on 8 cores)
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O                          # noqa: E402
from concurrentproject_b200 import rng          # noqa: E402

CASES = {  # name: (n, m, seed) -- sequences are rng.random_acgt(seed, 0, n) and rng.random_acgt(seed, 1, m)
    "cfg2": (100000, 100000, 2),
    "ring400k": (400000, 400000, 7),
    "n1m": (1000000, 1000000, 6),
    "cfg3": (4000000, 4000000, 3),
}

out_path = ROOT / "tests" / "golden" / "large_scores.json"
done = json.loads(out_path.read_text()) if out_path.exists() else {}
for name in (sys.argv[1:] or list(CASES)):
    n, m, seed = CASES[name]
    a, b = rng.random_acgt(seed, 0, n), rng.random_acgt(seed, 1, m)
    t0 = time.time()
    s = O.gotoh_fast(a, b)
    dt = time.time() - t0
    done[name] = {"n": n, "m": m, "seed": seed, "params": [1, -1, 1, 1], "score": int(s),
                  "by": "oracle_gotoh_fast (oracle/gotoh_fast.c)", "seconds": round(dt, 1)}
    out_path.write_text(json.dumps(done, indent=1, sort_keys=True) + "\n")
    print(name, s, f"{dt:.1f}s", f"{n * m / dt / 1e9:.1f} GCUPS", flush=True)
