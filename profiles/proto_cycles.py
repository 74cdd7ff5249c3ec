"""Research prototype (NumPy, CPU; uses tests/mg_reference.py and the oracle - not product code): iteration counts of
the multigrid-preconditioned solve for other cycles and outer iterations than the engine's V(1,1) / BiCGSTAB:
V(1,0), V(0,1), other dampings, zebra line Gauss-Seidel, GCR with one V-cycle per iteration.
    python profiles/proto_cycles.py
Results (256^2 / 512^2, tol 1e-14): see DESIGN.md section 9."""
import sys, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ('tests', 'oracle', ''):
    sys.path.insert(0, os.path.join(ROOT, d))
import mg_reference as MG
from test_mg_reference import scaled_system

def cycle(levels, r, k, pre, post, omega, coarse_sweeps=2):
    L=levels[k]
    if k==len(levels)-1:
        z=omega*L.line(r)
        for _ in range(coarse_sweeps-1): z=z+omega*L.line(r-L.apply(z))
        return z
    if pre:
        z=omega*L.line(r); res=L.om*(r-L.apply(z))
    else:
        z=np.zeros_like(r); res=L.om*r
    rc=(res[:,0::2]+res[:,1::2])/levels[k+1].om
    z=z+np.repeat(cycle(levels,rc,k+1,pre,post,omega,coarse_sweeps),2,axis=1)
    if post: z=z+omega*L.line(r-L.apply(z))
    return z

def iters(levels, rhs, prec, tol=1e-14, maxit=200):
    A=levels[0]; x=np.zeros_like(rhs); r=rhs.copy(); rho=alpha=omega=1.0; p=v=None
    for it in range(1,maxit+1):
        rho_new=np.vdot(rhs,r)
        p=r.copy() if it==1 else r+(rho_new/rho)*(alpha/omega)*(p-omega*v)
        ph=prec(p); v=A.apply(ph); alpha=rho_new/np.vdot(rhs,v); s=r-alpha*v
        if np.max(np.abs(s))<=tol: return it-0.5
        sh=prec(s); t=A.apply(sh); omega=np.vdot(t,s)/np.vdot(t,t)
        x+=alpha*ph+omega*sh; r=s-omega*t; rho=rho_new
        if np.max(np.abs(r))<=tol: return it
    return maxit

for n in (256,512):
    w,om,rhs0,c,fn=scaled_system(n,n)
    nlev=MG.level_count(n,n)
    lv=MG.hierarchy(*w, np.ones_like(om)*1.0 if False else om, nlev)
    A=lv[0]
    rhs=rhs0-A.apply(np.ones_like(rhs0))
    for name,pre,post,om_ in (("V(1,1) w=0.7",1,1,0.7),("V(1,0) w=0.7",1,0,0.7),("V(0,1) w=0.7",0,1,0.7),("V(0,1) w=0.8",0,1,0.8),("V(0,1) w=1.0",0,1,1.0),("V(1,0) w=1.0",1,0,1.0),("V(1,1) w=0.8",1,1,0.8),("V(1,1) w=1.0",1,1,1.0)):
        t0=time.time()
        it=iters(lv,rhs,lambda b: cycle(lv,b,0,pre,post,om_))
        print(n,nlev,name,"iterations",it,f"{time.time()-t0:.1f}s",flush=True)

print("zebra variants")
def zebra(L, z, r, order, omega=1.0):
    # line Gauss-Seidel over columns of one colour at a time: z_c += omega T_c^-1 (r - A z)_c
    for colour in order:
        res=r-L.apply(z)
        corr=L.line(res)   # solves all columns; only the colour's columns are used
        z=z.copy(); z[:,colour::2]+=omega*corr[:,colour::2]
    return z
def cycle_z(levels, r, k, omega=1.0, coarse_sweeps=1):
    L=levels[k]
    z=zebra(L,np.zeros_like(r),r,(0,1),omega)
    if k==len(levels)-1:
        for _ in range(coarse_sweeps-1): z=zebra(L,z,r,(0,1),omega)
        return z
    res=L.om*(r-L.apply(z))
    rc=(res[:,0::2]+res[:,1::2])/levels[k+1].om
    z=z+np.repeat(cycle_z(levels,rc,k+1,omega,coarse_sweeps),2,axis=1)
    return zebra(L,z,r,(1,0),omega)
for n in (256,512):
    w,om,rhs0,c,fn=scaled_system(n,n)
    nlev=MG.level_count(n,n)
    lv=MG.hierarchy(*w, om, nlev)
    A=lv[0]
    rhs=rhs0-A.apply(np.ones_like(rhs0))
    for om_ in (1.0,0.9,1.1):
        it=iters(lv,rhs,lambda b: cycle_z(lv,b,0,om_))
        print(n,nlev,"zebra V(1,1) w=%.1f"%om_,"iterations",it,flush=True)

print("GCR (1 V-cycle per iteration)")
def gcr(levels, rhs, prec, tol=1e-14, maxit=100, trunc=None):
    A=levels[0]; x=np.zeros_like(rhs); r=rhs.copy(); P=[]; AP=[]
    for it in range(1,maxit+1):
        z=prec(r); q=A.apply(z)
        lo=0 if trunc is None else max(0,len(P)-trunc)
        for pj,qj in zip(P[lo:],AP[lo:]):
            b=np.vdot(qj,q); z=z-b*pj; q=q-b*qj
        nq=np.sqrt(np.vdot(q,q)); z/=nq; q/=nq
        P.append(z); AP.append(q)
        a=np.vdot(q,r); x+=a*z; r-=a*q
        if np.max(np.abs(r))<=tol: return it
    return maxit
for n in (256,512):
    w,om,rhs0,c,fn=scaled_system(n,n)
    nlev=MG.level_count(n,n)
    lv=MG.hierarchy(*w, om, nlev)
    A=lv[0]
    rhs=rhs0-A.apply(np.ones_like(rhs0))
    for tr in (None,8,4,2):
        it=gcr(lv,rhs,lambda b: cycle(lv,b,0,1,1,0.7),trunc=tr)
        print(n,"GCR trunc",tr,"V-cycles",it,"(BiCGSTAB: 2 per iteration)",flush=True)
