import sys, time, threading, os
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from util import load_model
from spatialpy_b200.ensemble import run_ensemble
fm=load_model('cdc42')
if len(sys.argv)>2: import torch; torch.cuda.init(); torch.cuda.synchronize()
lanes=int(sys.argv[1])
run_ensemble(fm, lanes, 1, devices=[0], lanes=lanes)
for T in (lanes*2, lanes*6):
    t=time.perf_counter(); r=run_ensemble(fm, T, 1000, devices=[0], lanes=lanes); w=time.perf_counter()-t
    print('lanes',lanes,'traj',T,'wall',round(w,2),'traj/s',round(T/w,1), 'torch' if len(sys.argv)>2 else '')
