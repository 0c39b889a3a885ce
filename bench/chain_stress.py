#!/usr/bin/env python
"""cfg2 through launch config 7 many times: every score must be the golden 11446 (hand-off races show up as rare wrong scores)."""
import sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api, rng
api.configure("spin_limit", 3000000)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = api.Context(0)
n = 100000
a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(2, 1, n).copy()).cuda()
for kw in ({"config": 7, "rows": 3}, {"config": 7, "rows": 3, "two_sided": -1}, {"config": 7, "rows": 3, "no_linear": True}, {"config": 7, "rows": 4}):
    bad = {}; ms = []
    for _ in range(reps):
        s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, **kw); ms.append(ctx.last_run()["engine_ms"])
        if s != 11446: bad[s] = bad.get(s, 0) + 1
    print(json.dumps({"kw": kw, "reps": reps, "bad": bad, "ms_min": round(min(ms), 4)}), flush=True)
