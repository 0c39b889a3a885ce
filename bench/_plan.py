import sys; sys.path.insert(0,'.')
import torch
from concurrentproject_b200 import api, rng
ctx = api.Context(0)
for n in (1000, 3000, 10000, 30000, 100000, 300000, 1000000):
    a = torch.from_numpy(rng.random_acgt(2,0,n).copy()).cuda(); b = torch.from_numpy(rng.random_acgt(2,1,n).copy()).cuda()
    for nl in (False, True):
        best=None
        for rep in range(3):
            s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, no_linear=nl)
            i = ctx.last_run(); best = i['engine_ms'] if best is None else min(best, i['engine_ms'])
        print(f"n={n} no_linear={nl} score={s} lanes={i['lanes']} R={i['rows']} cfg={i['config']} ctas={i['ctas']} launches={i['engine_launches']} ms={best:.3f} gcups={n*n/best/1e6:.0f}", flush=True)
