import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import pnumpy_b200 as pn
pn.init(); pn.enable()
pn.recycler_enable(kind="managed")
def med(fn, k=50):
    fn(); ts=[]
    for _ in range(k):
        t0=time.perf_counter(); fn(); ts.append(time.perf_counter()-t0)
    return np.median(ts)*1e6
n=1_000_000
for dt in (np.int8, np.uint8, np.int16):
    a=(np.arange(n,dtype=np.int64)%253+1).astype(dt); b=a[5]
    c=np.true_divide(a,b)
    pn.cuda_reset_stats()
    t=med(lambda: np.true_divide(a,b,out=c))
    i=pn.cuda_info()
    print(dt.__name__, "a/5 us", round(t,1), "calls", i["calls"], "declined", i["declined"])
    pn.atop_disable(); t0=med(lambda: np.true_divide(a,b,out=c)); pn.atop_enable()
    print("   atop off us", round(t0,1))
    s=np.sin(a)
    pn.cuda_reset_stats()
    t=med(lambda: np.sin(a,out=s), 10)
    i=pn.cuda_info()
    print(dt.__name__, "sin us", round(t,1), "calls", i["calls"], "declined", i["declined"], s.dtype)
    pn.atop_disable(); t0=med(lambda: np.sin(a,out=s), 10); pn.atop_enable()
    print("   atop off us", round(t0,1))
pn.recycler_disable()
# what reaches the hooked loops for np.sin(int16_array): sizes and where they ran
pn.recycler_enable(kind="managed")
a=(np.arange(n,dtype=np.int64)%253+1).astype(np.int16); s=np.sin(a)
pn.ledger_enable(); pn.ledger_clear()
np.sin(a,out=s)
pn.ledger_disable()
for r in pn.ledger_records(): print({k:r[k] for k in ("category","gpu","elements","bytes","ticks")})
print("threshold", pn.cuda_getthreshold(), pn.recycler_info())
pn.recycler_disable()
