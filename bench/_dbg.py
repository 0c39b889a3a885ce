import sys, os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, oracle_lib as O
from concurrentproject_b200 import api, rng
found=[]
for p in ((3,-2,2,2),(1,-1,1,1),(2,-3,5,1)):
  for n in (520, 600, 700, 1000, 1500, 2000, 3000):
    for seed in range(12):
        a = rng.random_acgt(2000+seed, 0, n); b = rng.random_acgt(2000+seed, 1, n)
        want = O.gotoh_rolling(a,b,p)
        for cfg in (1,3):
            got = api.score(a,b,p,rows=1,config=cfg,two_sided=1)
            if got != want: found.append((p,n,seed,cfg,got,want)); print('BAD',p,n,seed,cfg,got,want,flush=True)
print('total bad',len(found))
