#!/usr/bin/env python
"""Scores one seeded pair a few times on GPU 0 (the command line ncu captures): python bench/one.py N SEED REPS '{"rows": 4}' [dbg]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api, rng  # noqa: E402

n, seed, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kw = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
if len(sys.argv) > 5:
    api.configure("dbg", sys.argv[5])
ctx = api.Context(0)
a = torch.from_numpy(rng.random_acgt(seed, 0, n).copy()).cuda()
b = torch.from_numpy(rng.random_acgt(seed, 1, n).copy()).cuda()
for _ in range(reps):
    s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, **kw)
    print(s, ctx.last_run(), flush=True)
