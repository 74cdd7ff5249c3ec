"""Research prototype (NumPy, CPU; oracle + tests/ restatements - not product code): iteration counts on ensemble members
(LC fields, D x a_m) of (1) the plain x-line preconditioner against 2-3 level V(1,1) cycles, (2) an fp32 inner solve with
fp64 iterative refinement.  Results in DESIGN.md section 9.
    python profiles/proto_ensemble.py"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ('tests','oracle',''): sys.path.insert(0, os.path.join(ROOT,d))
import ppfv_oracle as O, mg_reference as MG, xline_reference as XL
from conftest import load_golden, bc_for
from sayram2d_b200 import fields
g=load_golden('lc80')
def system(a,b,nsteps_before=3):
    m=O.Mesh(g['x_edges'],g['y_edges'],g['meta']['dt']); eq=O.Equation(m)
    eq.G,eq.Dxx,eq.Dxy,eq.Dyy,eq.inv_tau=g['G'],g['Dxx']*a,g['Dxy']*a,g['Dyy']*a,g['inv_tau']*b
    bct,lines=bc_for('LC',g['x_edges'],g['y_edges'])
    eq.bc=list(bct); eq.dirichlet_lines=lambda t: lines; eq.init_f=lambda: g['f_0']
    s=O.Solver(m,eq)
    for _ in range(nsteps_before): s.update()
    op=s.assemble(); c=s.f; om=op['diag']*c
    w=[np.zeros_like(c) for _ in range(4)]
    w[0][1:]=op['W'][1:]*c[:-1]; w[1][:-1]=op['E'][:-1]*c[1:]; w[2][:,1:]=op['S'][:,1:]*c[:,:-1]; w[3][:,:-1]=op['N'][:,:-1]*c[:,1:]
    w=[x/om for x in w]
    rhs=op['R']/om - XL.apply_A(w,np.ones_like(c))
    return w,om,rhs
def bicg(A,rhs,prec,tol=1e-14,maxit=300):
    x=np.zeros_like(rhs); r=rhs.copy(); rho=alpha=omega=1.0; p=v=None
    for it in range(1,maxit+1):
        rho_new=np.vdot(rhs,r)
        p=r.copy() if it==1 else r+(rho_new/rho)*(alpha/omega)*(p-omega*v)
        ph=prec(p); v=A(ph); alpha=rho_new/np.vdot(rhs,v); s=r-alpha*v
        sh=prec(s); t=A(sh); omega=np.vdot(t,s)/np.vdot(t,t)
        x+=alpha*ph+omega*sh; r=s-omega*t; rho=rho_new
        if np.max(np.abs(r))<=tol: return it
    return maxit
for mem in (0,21,42,63, 63+64*63):
    a,b=fields.ensemble_scales(np.array([mem])); a=float(a[0]); b=float(b[0])
    w,om,rhs=system(a,b)
    A=lambda x: XL.apply_A(w,x)
    lv1=MG.hierarchy(*w,om,1)
    it_x=bicg(A,rhs,lambda r: lv1[0].line(r))
    out=[f"member {mem} a={a:.2f} b={b:.2f}: x-line {it_x}"]
    for nlev in (2,3):
        lv=MG.hierarchy(*w,om,nlev)
        it=bicg(A,rhs,lambda r: MG.vcycle(lv,r))
        out.append(f"MG{nlev} V(1,1) {it}")
    print(' | '.join(out),flush=True)

print("mixed precision iterative refinement (fp32 inner x-line BiCGSTAB, fp64 residual)")
def bicg32(w32, lv32, rhs32, rtol, maxit=200):
    A=lambda x: XL.apply_A(w32,x)
    prec=lambda r: lv32.line(r)
    x=np.zeros_like(rhs32); r=rhs32.copy(); rho=alpha=omega=np.float32(1.0); p=v=None
    r0=np.max(np.abs(rhs32))
    for it in range(1,maxit+1):
        rho_new=np.vdot(rhs32,r)
        p=r.copy() if it==1 else r+(rho_new/rho)*(alpha/omega)*(p-omega*v)
        ph=prec(p); v=A(ph); alpha=rho_new/np.vdot(rhs32,v); s=r-alpha*v
        sh=prec(s); t=A(sh); omega=np.vdot(t,s)/np.vdot(t,t)
        x+=alpha*ph+omega*sh; r=s-omega*t; rho=rho_new
        if np.max(np.abs(r))<=rtol*r0: return x,it
    return x,maxit
for mem in (0,21,42,63, 63+64*63):
    a,b=fields.ensemble_scales(np.array([mem])); a=float(a[0]); b=float(b[0])
    w,om,rhs=system(a,b)
    w32=[x.astype(np.float32) for x in w]
    lv32=MG.Level(*w32, om.astype(np.float32))
    for rtol in (1e-3,1e-4,1e-5):
        x=np.zeros_like(rhs); r=rhs.copy(); total=0; outer=0
        while np.max(np.abs(r))>1e-14 and outer<12:
            scale=np.max(np.abs(r))
            dx,it=bicg32(w32,lv32,(r/scale).astype(np.float32),np.float32(rtol))
            x+=scale*dx.astype(np.float64); r=rhs-XL.apply_A(w,x); total+=it; outer+=1
        print(f"member {mem} a={a:.2f}: inner rtol {rtol:g}: outer {outer}, total inner iterations {total}, final resid {np.max(np.abs(r)):.1e}",flush=True)
