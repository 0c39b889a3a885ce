import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from concurrentproject_b200 import api, rng
LT = 60000
t = rng.random_acgt(7,1,LT)
for lanes,rows,config,nb in [(16,2,1,1),(16,2,1,2),(16,2,1,4),(16,8,1,2)]:
    rpb = (64 if lanes==16 else 32)*rows
    q = rng.random_acgt(7,0,rpb*nb)
    s = api.score(q,t,lanes=lanes,rows=rows,config=config,no_linear=True,orient=1)
    info = api.last_run()
    print(f"dbg={os.environ.get('SWB200_DBG')} lanes={lanes} rows={rows} config={config} bands={info['bands']} ms={info['engine_ms']:.3f} ns/step={info['engine_ms']*1e6/LT:.1f}", flush=True)
