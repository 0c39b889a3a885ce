#!/usr/bin/env python
"""Launch config 7 (CTA-chained engine) against the CPU oracle on small shapes, then timed on cfg2: python bench/chain_test.py [quick]"""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O
from concurrentproject_b200 import api, rng
api.configure("spin_limit", 3000000)
ctx = api.Context(0)
bad = 0
def check(a, b, p, want, **kw):
    global bad
    ta = torch.from_numpy(np.ascontiguousarray(a)).cuda(); tb = torch.from_numpy(np.ascontiguousarray(b)).cuda()
    try:
        got = ctx.score_device(ta.data_ptr(), len(a), tb.data_ptr(), len(b), params=p, config=7, lanes=16, rebase=-1, **kw)
        i = ctx.last_run()
    except Exception as e:
        got, i = "ERR " + str(e)[:120], {}
    ok = got == want or (want > 32000 and isinstance(got, str) and "ERR_RANGE" in got)    # plain 16-bit lanes were forced
    bad += not ok
    print(json.dumps({"n": len(a), "m": len(b), "p": p, "kw": kw, "want": want, "got": got, "ok": ok, "cfg": i.get("config"), "ts": i.get("two_sided"), "lin": i.get("linear"), "bands": i.get("bands"), "ctas": i.get("ctas")}), flush=True)
cases = [(300, 280, 1), (700, 900, 1), (2000, 2100, 2), (5000, 4800, 3), (9000, 9100, 4), (20000, 3000, 3), (3000, 20000, 2), (30000, 30000, 3), (1000, 50, 1), (64, 64, 1), (40000, 41000, 6)]
for k, (n, m, R) in enumerate(cases):
    a = rng.random_acgt(900 + k, 0, n)
    b = rng.mutate(a, 900 + k, 1, 0.08, 0.03)
    b = (np.concatenate([b, rng.random_acgt(900 + k, 2, max(0, m - len(b)))]))[:m]
    for p in ((1, -1, 1, 1), (2, -3, 5, 1)):
        want = O.gotoh_mt(a, b, p)
        for ts in (-1, 1):
            if ts == 1 and max(n, m) < 16 * 64 * R: continue
            check(a, b, p, want, rows=R, two_sided=ts)
            if p[2] == p[3]: check(a, b, p, want, rows=R, two_sided=ts, no_linear=True)
print("BAD", bad, flush=True)
if bad == 0 and len(sys.argv) < 2:
    n = 100000
    a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(2, 1, n).copy()).cuda()
    for kw in ({}, {"config": 7}, {"config": 7, "rows": 3}, {"config": 7, "rows": 4}, {"config": 7, "rows": 2}, {"config": 7, "rows": 3, "two_sided": -1}, {"config": 7, "no_linear": True}, {"no_linear": True},
               {"config": 7, "rows": 3, "no_linear": True}, {"config": 7, "rows": 4, "no_linear": True}):
        ms = []
        for _ in range(5):
            s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, **kw); i = ctx.last_run(); ms.append(i["engine_ms"])
        print(json.dumps({"cfg2": kw, "score": s, "ms_min": round(min(ms), 4), "ms_med": round(sorted(ms)[2], 4), "R": i["rows"], "config": i["config"], "ts": i["two_sided"], "ctas": i["ctas"], "bands": i["bands"]}), flush=True)
