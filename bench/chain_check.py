#!/usr/bin/env python
"""Workload of bench/chain_check.sh: launch config 7 (the CTA-chained engine) over rows x {linear, affine} x {one-, two-sided}
on cfg2 and on ragged / boundary-aligned sizes; every score is compared with the 32-bit engine.  With a -DSWB_CHAIN_CHECK
build of swb_chain.cu the library prints one "chaincheck:" line per run (reads checked, mismatches)."""
import sys, json, itertools
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api, rng
api.configure("spin_limit", 3000000)
ctx = api.Context(0)
runs = bad = 0
for n, m in ((100000, 100000), (65536, 65536 + 256), (40000, 40192), (30000, 120064), (99999, 50001), (8191, 70000)):
    a = torch.from_numpy(rng.random_acgt(7, 0, n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(7, 1, m).copy()).cuda()
    for p in ((1, -1, 1, 1), (2, -3, 5, 1)):
        want = ctx.score_device(a.data_ptr(), n, b.data_ptr(), m, params=p, lanes=32)
        for rows, ts, nl in itertools.product((1, 2, 3, 4, 6, 8), (-1, 1), (False, True)):
            if p[2] != p[3] and not nl: continue
            try:
                got = ctx.score_device(a.data_ptr(), n, b.data_ptr(), m, params=p, lanes=16, config=7, rows=rows, two_sided=ts, no_linear=nl)
            except Exception as e:
                if "launch config 7 needs" in str(e) or "no kernel" in str(e): continue
                got = "ERR " + str(e)[:80]
            runs += 1; bad += got != want
print(json.dumps({"runs": runs, "wrong_scores": bad}))
