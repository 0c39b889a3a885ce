#!/usr/bin/env python
"""Step cost and per-band lag of launch configs 1 and 7: one-sided sweeps of a 100 000-row Q against T of several lengths
(time = (bands * lag + LT) * cycles per step): python bench/chain_fit.py [rows]"""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from concurrentproject_b200 import api, rng
api.configure("spin_limit", 3000000)
ctx = api.Context(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = 100000
a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(2, 1, 400000).copy()).cuda()
for nl in (False, True):
    for cfg in (1, 7):
        pts = []
        for m in (25000, 50000, 100000, 200000, 400000):
            ms = []
            for _ in range(3):
                s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), m, config=cfg, rows=R, two_sided=-1, orient=1, no_linear=nl); i = ctx.last_run(); ms.append(i["engine_ms"])
            pts.append((m, min(ms)))
        (m0, t0), (m1, t1) = pts[1], pts[-1]
        slope = (t1 - t0) / (m1 - m0)                 # ms per step
        cyc = slope * 1e-3 * 1.965e9
        fill_steps = pts[2][1] / slope - pts[2][0]
        print(json.dumps({"config": cfg, "R": R, "affine": nl, "bands": i["bands"], "ms": [round(t, 3) for _, t in pts], "cyc_per_step": round(cyc, 2), "fill_steps": round(fill_steps), "lag_per_band": round(fill_steps / i["bands"], 1)}), flush=True)
