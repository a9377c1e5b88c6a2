"""Accuracy of the square-root-free (scaled) Schur recursion of gpf_schur.cuh next to the normalised one and LAPACK Cholesky,
against 60-digit factorisations of the slice covariance K + mean(K) 1e-6 I on the uniform grid (numpy restatement of both).

    python tools/schur_scaled_probe.py
"""
import numpy as np, mpmath as mp
def lag(n, sigma, ls):
    x = np.linspace(-1,1,n); d2=(x-x[0])**2
    t = sigma*np.exp(-0.5*d2/ls**2)
    part = sum((n if k==0 else 2*(n-k))*t[k] for k in range(n))
    jit = part/(n*n)*1e-6
    t = t.copy(); t[0]+=jit
    return t
def schur_mixed(t, R):
    n=len(t); uinv=1/np.sqrt(t[0]); col=t*uinv; v=col.copy(); v[0]=0
    R=R.copy(); quad=np.zeros(R.shape[0]); ld=0.0
    for k in range(n):
        w = R[:,k]*uinv; quad+=w*w
        R[:,k:] -= np.outer(w, col[:n-k])
        ld += -2*np.log(uinv)
        if k==n-1: break
        rho=v[k+1]*uinv; e=1-rho*rho; cinv=1/np.sqrt(e); cc=e*cinv; uinv*=cinv
        m=n-1-k
        un=(col[:m]-rho*v[k+1:])*cinv
        v[k+1:]=cc*v[k+1:]-rho*un
        col[:m]=un
    return ld, quad
def schur_scaled(t, R):
    n=len(t); ginv=1/t[0]; col=t.copy(); v=col.copy(); v[0]=0
    R=R.copy(); quad=np.zeros(R.shape[0]); ld=0.0
    for k in range(n):
        wt = R[:,k]*ginv; quad+=R[:,k]*wt
        R[:,k:] -= np.outer(wt, col[:n-k])
        ld += -np.log(ginv)
        if k==n-1: break
        rho=v[k+1]*ginv; e=1-rho*rho; ginv*= 1/e
        m=n-1-k
        un=col[:m]-rho*v[k+1:]
        v[k+1:]=e*v[k+1:]-rho*un
        col[:m]=un
    return ld, quad
def truth(t, R):
    mp.mp.dps=60
    n=len(t)
    T=mp.matrix(n,n)
    for i in range(n):
        for j in range(n): T[i,j]=mp.mpf(t[abs(i-j)])
    L=mp.cholesky(T)
    ld=2*sum(mp.log(L[i,i]) for i in range(n))
    qs=[]
    for r in R:
        y=mp.lu_solve(T, mp.matrix(r)); qs.append(sum(mp.mpf(r[i])*y[i] for i in range(n)))
    return float(ld), np.array([float(q) for q in qs])
rng=np.random.default_rng(0)
for n in (64,128):
  for ls in (0.1,0.3,1.0,3.0,10.0):
    t=lag(n,1.0,ls)
    # functions: smooth-ish draws
    x=np.linspace(-1,1,n)
    R=np.stack([np.sin(3*x)+0.1*rng.standard_normal(n)*0.01, rng.standard_normal(n)*0.1])
    a=schur_mixed(t,R); b=schur_scaled(t,R); tr=truth(t,R)
    Tm=np.array([[t[abs(i-j)] for j in range(n)] for i in range(n)]); Lc=np.linalg.cholesky(Tm)
    ldc=2*np.log(np.diag(Lc)).sum(); qc=((np.linalg.solve(Lc,R.T))**2).sum(0)
    print(n, ls, "ld err mixed %.2e scaled %.2e chol %.2e"%(abs(a[0]-tr[0]),abs(b[0]-tr[0]),abs(ldc-tr[0])),
          "quad rel mixed", np.abs(a[1]-tr[1])/np.abs(tr[1]), "scaled", np.abs(b[1]-tr[1])/np.abs(tr[1]), "chol", np.abs(qc-tr[1])/np.abs(tr[1]))
