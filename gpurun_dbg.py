import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from util import *
from spatialpy_b200.engine import Engine
name='diffusion3d'
fm=load_model(name); ref=load_ref(name)
eng=Engine(fm, flags=0)
eng.reset(1000); eng.step(1)
for f in ('Q','C','F','Fbp','Frho','rho'):
    a=eng.get(f); b=ref[f's1_{f}']
    print(f, rel_err(a,b), a.shape)
a=eng.get('Q'); b=ref['s1_Q']
bad=np.where(np.abs(a-b).max(axis=1)>1e-9*np.abs(b).max())[0]
print(len(bad), bad[:20])
for i in bad[:5]:
    print(i, fm.x[i], fm.type[i], a[i], b[i], eng.fm.u0[i])
print('C0', ref['s0_C'][:3], 'Q0', ref['s0_Q'][:3])
