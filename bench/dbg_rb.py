import sys, json, torch
sys.path.insert(0, '.')
from concurrentproject_b200 import api, rng
ctx = api.Context(0)
cases = [(20000, 1, 1), (20000, 1, 2), (60000, 2, 2), (100000, 1, 2), (100000, 2, 4), (200000, 2, 8), (400000, 4, 16), (1000000, 8, 0), (2000000, 8, 0)]
for n, R, ctas in cases:
    a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(2, 1, n).copy()).cuda()
    out = {}
    for cfg in (3, 1):
        try:
            s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, rows=R, config=cfg, ctas=ctas, rebase=1, two_sided=-1, lanes=16)
            out[cfg] = (s, ctx.last_run()["bands"], ctx.last_run()["warps"])
        except Exception as e:
            out[cfg] = str(e)[:80]
    print(json.dumps({"n": n, "R": R, "ctas": ctas, "cfg3": out[3], "cfg1": out[1], "same": out[3][0] == out[1][0] if isinstance(out[3], tuple) and isinstance(out[1], tuple) else None}), flush=True)
