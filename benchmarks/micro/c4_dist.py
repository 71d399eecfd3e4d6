import sys, numpy as np
sys.path.insert(0,'coordinatedescent.jl_b200'); sys.path.insert(0,'.')
import cdgpu
from cdgpu import *
from bench import c4_data
gpu=cdgpu.default()
X,Z,Y=c4_data(); m=4096; zg=np.linspace(0.01,0.99,m)
for tol in (1e-7,1e-9):
    out,_=gpu.locpolyl1(X,Z,Y,zg,2,GaussianKernel(0.2),0.01,False,CDOptions(randomize=False,optTol=tol))
    st=gpu.last_vc_stats
    v=np.array([s['visits'] for s in st]); ps=np.array([s['passes'] for s in st]); ac=np.array([s['accepted'] for s in st])
    print('tol',tol,'device_ms',st[0]['device_ms'],'visits sum',v.sum(),'accepted sum',ac.sum())
    print(' passes pct', np.percentile(ps,[0,50,90,99,100]), 'accepted pct', np.percentile(ac,[0,50,90,99,100]))
    top=np.argsort(-ac)[:12]; print(' top idx',top.tolist(), ac[top].tolist())
    # accepted by decile of index
    print(' accepted by index decile', [int(ac[i*410:(i+1)*410].sum()/1e3) for i in range(10)])
    nn=(out!=0).sum(0); print(' nnz pct', np.percentile(nn,[0,50,90,100]))
