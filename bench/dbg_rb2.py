import sys, json, torch
sys.path.insert(0, '.')
from concurrentproject_b200 import api, rng
ctx = api.Context(0)
def run(n, **kw):
    a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(2, 1, n).copy()).cuda()
    s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, **kw); i = ctx.last_run()
    print(json.dumps({"n": n, "kw": kw, "score": s, "lanes": i["lanes"], "rb": i["rebased"], "lin": i["linear"], "bands": i["bands"], "warps": i["warps"], "ms": round(i["engine_ms"], 1)}), flush=True)
ref = {}
for n in (3000000, 3500000, 4000000):
    run(n, rows=14, config=2, two_sided=-1)
    run(n, rows=8, config=1, two_sided=-1)
run(4000000, rows=8, config=1, two_sided=-1, no_linear=True)
run(4000000, rows=8, config=1, two_sided=1)
run(4000000, rows=8, config=1, two_sided=-1, ctas=74)
run(4000000, rows=10, config=1, two_sided=-1)
run(4000000, rows=12, config=1, two_sided=-1)
run(4000000, rows=8, config=4, two_sided=-1)
