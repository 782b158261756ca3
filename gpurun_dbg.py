import sys, time
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from util import load_model
import test_gpu_stat as T
from spatialpy_b200.engine import Engine, FLAG_NO_VTK, FLAG_SKIP_STATIC_FORCES, FLAG_LEAP_DIFFUSION
F=FLAG_SKIP_STATIC_FORCES|FLAG_LEAP_DIFFUSION
for name in ('cdc42','birth_death'):
    fm=load_model(name)
    eng=Engine(fm, flags=F|FLAG_NO_VTK)
    eng.reset(3); t=time.perf_counter(); eng.step(fm.nt); w=time.perf_counter()-t
    xx=eng.get('xx'); print(name,'leap traj ms',round(w*1e3,1), eng.counters(), 'total', xx.sum(axis=0)[:4])
    eng.close()
for name,n in (('cylinder',600),('birth_death',300)):
    try:
        T.check_against_reference(name, n, flags=F); print(name,'LEAP PARITY OK')
    except AssertionError as e:
        print(name,'LEAP PARITY FAIL', str(e)[:300])
