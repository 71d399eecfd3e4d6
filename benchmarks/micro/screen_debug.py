import os, sys, numpy as np, subprocess, json
sys.path.insert(0,'coordinatedescent.jl_b200'); sys.path.insert(0,'tests')
import cdgpu
from cdgpu import *
from helpers import gauss_problem
gpu=cdgpu.default()
def run(n,p,s,lam,weighted,rand,screen,tol=1e-12):
    os.environ["CDGPU_COV_SCREEN"]=str(screen)
    X,y,_=gauss_problem(n,p,s,seed=11)
    A=np.asfortranarray(X.T@X/n); b=-X.T@y/n
    om=np.sqrt(np.diag(A)) if weighted else None
    f=gpu.CDQuadraticLoss(A,b); x=SparseIterate(p)
    gpu.coordinateDescent_(x,f,ProxL1(lam,om),CDOptions(maxIter=20000,optTol=tol,randomize=bool(rand),seed=3))
    st=f.last_stats
    return x.toarray(), st, f.Ax, A, b
for (n,p,s,lam) in [(200,50,10,0.2),(500,300,20,0.05),(300,1000,15,0.1),(300,1000,15,0.3)]:
  for weighted in (False,True):
    for rand in (0,1):
        b0,s0,ax0,A,b=run(n,p,s,lam,weighted,rand,0)
        b1,s1,ax1,_,_=run(n,p,s,lam,weighted,rand,1)
        ok=np.array_equal(b0!=0,b1!=0) and np.max(np.abs(b0-b1))<1e-9
        print((n,p,s,lam,weighted,rand),'OK' if ok else 'MISMATCH', 'passes',s0['passes'],s1['passes'],'visits',s0['visits'],s1['visits'],'nnz',np.count_nonzero(b0),np.count_nonzero(b1),'maxdiff',np.max(np.abs(b0-b1)), 'Ax err', np.max(np.abs(ax1-A@b1)), np.max(np.abs(ax0-A@b0)))
print("---- detail")
for mode in (1,2):
    b1,s1,ax1,A,b=run(300,1000,15,0.1,True,1,mode)
    err=np.abs(ax1-A@b1); print("mode",mode,"bad rows",np.flatnonzero(err>1e-9).tolist(), err.max())
    b1,s1,ax1,A,b=run(300,1000,15,0.1,False,1,mode)
    err=np.abs(ax1-A@b1); print("mode",mode,"unweighted bad rows",np.flatnonzero(err>1e-9).tolist(), err.max())
b1,s1,ax1,A,b=run(300,1000,15,0.1,True,1,1)
err=np.abs(ax1-A@b1); bad=np.flatnonzero(err>1e-9)
print("bad rows",len(bad), bad[:40].tolist()); print("in support:",[int(b1[j]!=0) for j in bad[:40]]); print("err",err[bad[:10]])
S=np.flatnonzero(b1); print("support",S.tolist())
# which single column k explains the error: err_j ~ A[j,k]*delta ?
r=(ax1-A@b1)[bad]
for k in range(1000):
    col=A[bad,k]
    if np.linalg.norm(col)>0:
        d=(col@r)/(col@col)
        if np.linalg.norm(r-d*col)<1e-6*np.linalg.norm(r): print("explained by column",k,"delta",d,"beta_k",b1[k])
