import sys, json, torch
sys.path.insert(0, '.')
from concurrentproject_b200 import api, rng
ctx = api.Context(0)
N = 4000512
A = torch.from_numpy(rng.random_acgt(2, 0, N).copy()).cuda(); B = torch.from_numpy(rng.random_acgt(2, 1, N).copy()).cuda()
def run(n, m, **kw):
    out = {}
    for name, k2 in (("c1", dict(config=1)), ("c3", dict(config=3))):
        try:
            s = ctx.score_device(A.data_ptr(), n, B.data_ptr(), m, two_sided=-1, rebase=1, lanes=16, **kw, **k2)
            i = ctx.last_run(); out[name] = s; out["bands"] = i["bands"]
        except Exception as e:
            out[name] = str(e)[:100]
    print(json.dumps({"n": n, "m": m, "kw": kw, **out, "same": out["c1"] == out["c3"]}), flush=True)
for n, m in [(3600000, 3600000), (3800000, 3800000), (3900000, 3900000), (3999744, 3999744), (3999900, 3999900), (4000000, 3999900), (3999900, 4000000),
             (4000000, 4000000), (4000256, 4000256), (4000100, 4000100), (4000000, 3000000), (3000000, 4000000)]:
    run(n, m, rows=8)
run(4000000, 4000000, rows=6)
run(3999900, 3999900, rows=6)
run(4000000, 4000000, rows=4)
run(4000000, 4000000, rows=10)
