import ctypes as C, sys, os, numpy as np, time, glob
sys.path.insert(0,'.')
from concurrentproject_b200 import rng
n=100000
a=rng.random_acgt(2,0,n); b=rng.random_acgt(2,1,n)
U8P=C.POINTER(C.c_ubyte)
path = sys.argv[1]
lib=C.CDLL(path); out=C.c_int(0)
ts=[]
for _ in range(5):
    t0=time.perf_counter()
    rc=lib.swb200_score(a.ctypes.data_as(U8P), n, b.ctypes.data_as(U8P), n, None, C.byref(out))
    ts.append(round((time.perf_counter()-t0)*1e3,2))
print(path.split('/')[1], rc, out.value, ts[1:], flush=True)
