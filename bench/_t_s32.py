from concurrentproject_b200 import api, rng
a=rng.random_acgt(2,0,100000); b=rng.random_acgt(2,1,100000)
for name,f in (("s32",lambda: api.score(a,b,lanes=32)), ("end",lambda: api.score_end(a,b)), ("span",lambda: api.score_span(a,b))):
    f(); r=f(); print(name, r, round(api.last_run()["engine_ms"],3))
