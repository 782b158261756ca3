import sys, time, threading
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from util import load_model
from spatialpy_b200.engine import Engine, FLAG_NO_VTK, FLAG_SKIP_STATIC_FORCES
fm=load_model('cdc42')
def worker(k, out, ntraj, mode):
    eng=Engine(fm, flags=FLAG_NO_VTK|FLAG_SKIP_STATIC_FORCES)
    eng.run_no_files(1+k,1)
    bar.wait()
    t=time.perf_counter(); tr=0.0
    for q in range(ntraj):
        if mode==0:
            t1=time.perf_counter(); eng.reset(100+k*10+q); tr+=time.perf_counter()-t1; eng.step(fm.nt)
        else:
            eng.run_no_files(100+k*10+q,1)
    w=time.perf_counter()-t
    out[k]=(w/ntraj*1e3, tr/ntraj*1e3)
    eng.close()
for mode in (0,1):
  for T in (1,8,16):
    out={}; bar=threading.Barrier(T)
    th=[threading.Thread(target=worker,args=(k,out,4,mode)) for k in range(T)]
    [t.start() for t in th]; [t.join() for t in th]
    print('mode',mode,'threads',T, 'wall/traj ms', round(sum(v[0] for v in out.values())/T,1), 'reset ms', round(sum(v[1] for v in out.values())/T,2))
