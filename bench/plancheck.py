"""Planner check on one GPU: the kernel variant swb200.cu's make_plan() picks for a pair of size N against every
(rows, launch config, two-sided) variant forced by hand.  usage: python bench/plancheck.py   (prints one line per N)"""
import sys, json, torch
sys.path.insert(0, '.')
from concurrentproject_b200 import api, rng
ctx = api.Context(0)
for n in (5000, 20000, 50000, 100000, 150000, 200000, 250000, 500000):
    a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(2, 1, n).copy()).cuda()
    def run(**kw):
        best = None
        for _ in range(2):
            s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, **kw); i = ctx.last_run()
            best = i if best is None or i["engine_ms"] < best["engine_ms"] else best
        return s, best
    s0, auto = run()
    res = []
    for cfg in (1, 2, 3, 7):
        for R in ((1, 2, 3, 4, 6, 8) if cfg == 7 else (2, 3, 4, 6, 8, 10, 12, 14, 16)):
            for ts in ((1, -1) if n <= 250000 else (-1,)):
                try:
                    s, i = run(rows=R, config=cfg, two_sided=ts)
                except Exception as e:
                    continue
                assert s == s0
                res.append((i["engine_ms"], R, cfg, i["two_sided"], i["rebased"]))
    res.sort()
    print(n, "auto", round(auto["engine_ms"], 3), (auto["rows"], auto["config"], auto["two_sided"], auto["rebased"]), "best3", [(round(m, 3), R, c, t, rb) for m, R, c, t, rb in res[:3]], flush=True)
