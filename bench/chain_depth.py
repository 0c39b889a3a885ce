#!/usr/bin/env python
"""Does a deep chain of bands run slower than a shallow one?  One GPU stands in for one rank of the 8-GPU ring: B bands of
R = 14 re-based rows, one warp per scheduler (launch config 3 or 1), one-sided, T of two lengths -> cycles per step
(difference quotient) and per-band lag; plus first / median / last band duration from the %globaltimer stamps.
usage: python bench/chain_depth.py [configs=3,1]"""
import sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from concurrentproject_b200 import api, rng
MHZ = 1965.0
ctx = api.Context(0)
configs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3,1").split(",")]
pf = ROOT / "gpurun_out" / "prof_depth.jsonl"
pf.parent.mkdir(exist_ok=True)
for config in configs:
    for B in (148, 296, 592):
        n = B * 896
        a = torch.from_numpy(rng.random_acgt(11, 0, n).copy()).cuda()
        res = {}
        for m in (200000, 400000):
            b = torch.from_numpy(rng.random_acgt(11, 1, m).copy()).cuda()
            kw = dict(rows=14, config=config, rebase=1, two_sided=-1, orient=1)
            best = None
            for _ in range(3):
                s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), m, **kw)
                info = ctx.last_run()
                best = info["engine_ms"] if best is None else min(best, info["engine_ms"])
            res[m] = best
            if m == 400000:
                if pf.exists(): pf.unlink()
                api.configure("prof", str(pf))
                ctx.score_device(a.data_ptr(), n, b.data_ptr(), m, **kw)
                api.configure("prof", "")
                line = json.loads(pf.read_text().strip().split("\n")[-1])
                w = np.array(line["warps"], dtype=np.float64); w = w[w[:, 3] > 0]
                w = w[np.argsort(w[:, 4])]
                dur = (w[:, 5] - w[:, 4]) / 1e3 if len(w) else np.zeros(1)      # stamps need a -DSWB_ENABLE_PROF build
        c = (res[400000] - res[200000]) * 1e-3 * MHZ * 1e6 / 200000
        lag = (res[400000] * 1e-3 * MHZ * 1e6 / c - 400000) / info["bands"]
        print(json.dumps({"config": config, "bands": info["bands"], "warps": info["warps"], "ms_200k": round(res[200000], 3), "ms_400k": round(res[400000], 3),
                          "cyc_per_step": round(c, 1), "lag_steps": round(lag, 1), "band_us_first": round(float(dur[0]), 1),
                          "band_us_median": round(float(np.median(dur)), 1), "band_us_last": round(float(dur[-1]), 1),
                          "first_band_cyc_per_step": round(float(dur[0]) * MHZ / 400000, 1), "last_band_cyc_per_step": round(float(dur[-1]) * MHZ / 400000, 1)}), flush=True)
