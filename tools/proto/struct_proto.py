import numpy as np, sys
sys.path.insert(0,'/root/repo')
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200 import synthetic

def ldl_nopiv(K):
    K=K.copy(); N=K.shape[0]; L=np.eye(N); d=np.zeros(N)
    for j in range(N):
        d[j]=K[j,j]
        L[j+1:,j]=K[j+1:,j]/d[j]
        K[j+1:,j+1:]-=np.outer(L[j+1:,j],L[j+1:,j])*d[j]
    return L,d
def ldl_solve(L,d,r):
    y=np.linalg.solve(L,r) if False else __import__('scipy.linalg',fromlist=['x']).solve_triangular(L,r,lower=True,unit_diagonal=True)
    y=y/d
    return __import__('scipy.linalg',fromlist=['x']).solve_triangular(L.T,y,lower=False,unit_diagonal=True)

def project_cone(f,mu):
    r=np.hypot(f[0],f[1])
    if mu*f[2]>=r: return f.copy()
    if f[2]<-mu*r: return np.zeros(3)
    ratio=mu*f[2]/r
    ray=np.array([ratio*f[0],ratio*f[1],f[2]])
    nr=np.linalg.norm(ray)
    if nr>0: ray=ray/nr
    return ray.dot(f)*ray

class Struct:
    """reduced KKT: order [R, y, D0]; D+ eliminated"""
    def __init__(self,Q,A,nr,shift,tau):
        n=Q.shape[0]; m=A.shape[0]
        qd=np.diag(Q)[nr:]+shift
        D=np.arange(nr,n)
        self.d0=D[qd<=tau]; self.dp=D[qd>tau]
        self.hp=qd[qd>tau]; h0=qd[qd<=tau]
        self.nr=nr; self.n=n; self.m=m
        AR=A[:,:nr]; Ap=A[:,self.dp]; A0=A[:,self.d0]
        self.Ap=Ap
        C=(Ap/self.hp)@Ap.T
        n0=len(self.d0)
        K=np.zeros((nr+m+n0,)*2)
        K[:nr,:nr]=Q[:nr,:nr]+shift*np.eye(nr)
        K[nr:nr+m,:nr]=AR; K[:nr,nr:nr+m]=AR.T
        K[nr:nr+m,nr:nr+m]=-C
        K[nr+m:,nr:nr+m]=A0.T; K[nr:nr+m,nr+m:]=A0
        K[nr+m:,nr+m:]=np.diag(h0)
        self.K=K
        self.L,self.d=ldl_nopiv(K)
    def solve(self,rx,ry):
        nr,m=self.nr,self.m
        rr=np.concatenate([rx[:nr], ry-self.Ap@(rx[self.dp]/self.hp), rx[self.d0]])
        s=ldl_solve(self.L,self.d,rr)
        x=np.zeros(self.n)
        x[:nr]=s[:nr]; y=s[nr:nr+m]; x[self.d0]=s[nr+m:]
        x[self.dp]=(rx[self.dp]-self.Ap.T@y)/self.hp
        return x

def solve_qp(Q,b,A,beq,mu,lb,ub,nc,lcs,nr,rho=5e-5,eps=1e-6,max_iter=100,tau=0.0):
    n=len(b)
    S0=Struct(Q,A,nr,0.0,tau)
    x=S0.solve(-b,beq)
    x0=x.copy()
    xbar=x.copy(); lcbar=x[lcs:lcs+nc].copy()
    mux=np.zeros(n); muc=np.zeros(nc)
    S1=None
    n_iter=max_iter
    for it in range(max_iter):
        if it>0:
            if S1 is None: S1=Struct(Q,A,nr,rho,tau)
            w=xbar-mux; w[lcs:lcs+nc]=lcbar-muc
            x=S1.solve(-b+rho*w,beq)
        xbar=np.clip(x+mux,lb,ub) if False else np.maximum(np.minimum(x+mux,ub),lb)
        for c in range(nc//3):
            lcbar[3*c:3*c+3]=project_cone(x[lcs+3*c:lcs+3*c+3]+muc[3*c:3*c+3],mu[c])
        rx=x-xbar; rc=x[lcs:lcs+nc]-lcbar
        mux+=rx; muc+=rc
        if np.abs(rc).max()<eps and np.abs(rx).max()<eps:
            n_iter=it;break
    return x,n_iter,x0

if __name__=="__main__":
    qp=load_walking_log()
    g=np.load('/root/repo/tests/golden/walking_cold.npz')
    nr=22
    tau=float(sys.argv[1]) if len(sys.argv)>1 else 0.0
    step=int(sys.argv[2]) if len(sys.argv)>2 else 10
    errs=[];mis=0
    for i in range(0,qp.batch,step):
        x,ni,x0=solve_qp(qp.Q[i],qp.b[i],qp.A_eq[i],qp.b_eq[i],qp.friction_coeffs[i],qp.lb[i],qp.ub[i],12,38,nr,tau=tau)
        e=np.abs(x-g['z'][i]).max()/max(1,np.abs(g['z'][i]).max())
        errs.append(e)
        if ni!=g['n_iter'][i]: mis+=1; print('mismatch',i,ni,g['n_iter'][i])
    errs=np.array(errs)
    print('max rel err',errs.max(),'p50',np.median(errs),'mismatch',mis,'of',len(errs))
