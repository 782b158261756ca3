import sys, time, os
sys.path.insert(0,'.')
from spatialpy_b200 import configs, Solver
fm=configs.cylinder_rdme(nt=1000, output_every=100)
sol=Solver(fm); sol.compile()
t=time.perf_counter(); res=sol.run(seed=1); w=time.perf_counter()-t
files=os.listdir(res.result_dir); size=sum(os.path.getsize(os.path.join(res.result_dir,f)) for f in files)
print('Solver.run 1M particles, 1000 steps,', len(files),'files', round(size/1e6),'MB, wall', round(w,2),'s')
