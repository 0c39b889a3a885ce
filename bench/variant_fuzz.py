#!/usr/bin/env python
"""Consistency sweep on one GPU: every (launch config, rows, re-based or not, two-sided or not, linear or affine) variant of
the pair engines must return the score of the 32-bit engine, on sizes chosen to hit block / chunk / ring boundaries
(multiples of 256 and 32, sizes just around them) and on planted pairs whose score level rises and collapses.
usage: python bench/variant_fuzz.py [seed]   -- prints one line per size and a final BAD count"""
import os, sys, json, itertools
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api, rng
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
api.configure("spin_limit", 4000000)
ctx = api.Context(0)
CONFIGS = tuple(int(x) for x in os.environ.get("FUZZ_CONFIGS", "1,2,3,4,5,6,7").split(","))   # FUZZ_CONFIGS=7: the chained engine only
EVERY = os.environ.get("FUZZ_EVERY", "0") == "1"                                                  # all combinations instead of a third per seed
bad = 0; runs = 0
def pair(kind, n, m, s):
    a = rng.random_acgt(s, 0, n)
    if kind == "random":
        b = rng.random_acgt(s, 1, m)
    elif kind == "same":
        b = a[:m].copy() if m <= n else np.concatenate([a, rng.random_acgt(s, 2, m - n)])
    else:   # level rises, collapses, rises again
        b = a.copy()[:m] if m <= n else np.concatenate([a, rng.random_acgt(s, 2, m - n)])
        lo, hi = len(b) // 3, len(b) // 3 + len(b) // 6
        b[lo:hi] = rng.random_acgt(s, 3, hi - lo)
    return a, b
sizes = [(65536, 65536), (65536 + 256, 65536), (65536, 65536 - 1), (40000, 40192), (131072, 131072), (100000, 99840), (30000, 120064), (120064, 30000), (262144, 262144), (200000, 262400)]
for (n, m), kind in itertools.product(sizes, ("random", "same", "collapse")):
    if kind != "random" and max(n, m) > 140000:
        continue
    a, b = pair(kind, n, m, seed * 1000 + n % 997)
    ta, tb = torch.from_numpy(np.ascontiguousarray(a)).cuda(), torch.from_numpy(np.ascontiguousarray(b)).cuda()
    for p in ((1, -1, 1, 1), (2, -3, 5, 1)):
        if kind != "random" and p[0] * min(n, m) > 400000:
            continue
        want = ctx.score_device(ta.data_ptr(), len(a), tb.data_ptr(), len(b), params=p, lanes=32)
        wrong = []
        for cfg, rows, rb, ts, nl in itertools.product(CONFIGS, (2, 3, 4, 6, 8, 10, 14, 16), (-1, 1), (-1, 1), (False, True)):
            if p[2] != p[3] and not nl: continue
            if cfg == 6 and rows > 10: continue
            if cfg == 7 and (rb == 1 or rows > 8): continue
            if rb == -1 and want > 32000: continue
            if rb == 1 and (p[0] + max(p[2], p[3])) * (64 * rows + 416) > 10000: continue
            if not EVERY and (seed + cfg + rows + (rb > 0) + (ts > 0) + nl + n // 256) % 3: continue          # a third of the combinations per seed
            try:
                got = ctx.score_device(ta.data_ptr(), len(a), tb.data_ptr(), len(b), params=p, lanes=16, config=cfg, rows=rows, rebase=rb, two_sided=ts, no_linear=nl)
            except Exception as e:
                msg = str(e)
                if "launch config 7 needs" in msg or "RANGE" in msg or "no kernel" in msg: continue
                got = "ERR " + msg[:80]
            runs += 1
            if got != want: wrong.append((cfg, rows, rb, ts, nl, got))
        bad += len(wrong)
        print(json.dumps({"n": n, "m": m, "kind": kind, "p": p, "want": want, "wrong": wrong[:6], "n_wrong": len(wrong)}), flush=True)
print("RUNS", runs, "BAD", bad)
