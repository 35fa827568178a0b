"""numpy model of the block one-sided Jacobi iteration of csrc/svd.cu (svd_pair_kernel): sweep counts for different rotation
schedules inside a block pair.  python tools/jacobi_model.py <n> <block> full|cross0|every4 ...   (CPU only, minutes for n = 1024)
Measured with it: n=1024, block 32: full 15 sweeps, cross-block-only except round 0: 16, full every 4th round: 15;
n=256, block 16: none / QR / QR twice preconditioning: 12 / 10 / 9 sweeps."""
import numpy as np, sys, time
def rr_pairs(nb, r):
    m = nb-1; out=[]
    for p in range(nb//2):
        if p==0: i,j=m, r%m
        else: i,j=(r+p)%m,(r-p+m)%m
        if i>j: i,j=j,i
        out.append((i,j))
    return out
def rot_params(gpq, app, aqq):
    dd=aqq-app; m2=2*gpq
    with np.errstate(all='ignore'):
        t = np.where(np.abs(gpq)>1e-300, np.where(dd>=0,1,-1)*m2/(np.abs(dd)+np.sqrt(dd*dd+m2*m2)), 0.0)
    c=1/np.sqrt(1+t*t); return c, t*c
def apply(G,J,p,q,c,s):
    n=G.shape[0]
    R=np.eye(n); R[p,p]=c; R[q,q]=c; R[p,q]=s; R[q,p]=-s
    return R.T@G@R, J@R
def inner(G, mode, B):
    n=G.shape[0]; J=np.eye(n)
    if mode=='full':
        for step in range(n-1):
            prs=rr_pairs(n,step); p=np.array([a for a,b in prs]); q=np.array([b for a,b in prs])
            c,s=rot_params(G[p,q],G[p,p],G[q,q]); G,J=apply(G,J,p,q,c,s)
    else:  # cross only: pairs (k, B + (k+step)%B)
        for step in range(B):
            p=np.arange(B); q=B+(p+step)%B
            c,s=rot_params(G[p,q],G[p,p],G[q,q]); G,J=apply(G,J,p,q,c,s)
    return J
def run(A,B,variant):
    m,n=A.shape; W=A.copy(); nb=n//B; tol=4*np.sqrt(m)*2.2e-16
    for sw in range(60):
        offmax=0
        for r in range(nb-1):
            for (i,j) in rr_pairs(nb,r):
                cols=np.r_[i*B:(i+1)*B, j*B:(j+1)*B]; P=W[:,cols]; G=P.T@P
                d=np.sqrt(np.diag(G)); off=np.abs(G)/np.outer(d,d); np.fill_diagonal(off,0); offmax=max(offmax,off.max())
                if off.max()<=tol: continue
                mode='full' if (variant=='full' or (variant=='cross0' and r==0) or (variant.startswith('every') and r%int(variant[5:])==0)) else 'cross'
                W[:,cols]=P@inner(G,mode,B)
        if offmax<=tol: return sw+1
    return -1
n=int(sys.argv[1]); B=int(sys.argv[2])
rng=np.random.default_rng(4); A=rng.standard_normal((n,n))/np.sqrt(n)
for v in sys.argv[3:]:
    t=time.time(); print(n,B,v,"sweeps",run(A,B,v),"%.0fs"%(time.time()-t),flush=True)
