import numpy as np, sys, scipy.linalg as sl
sys.path.insert(0,'/root/repo')
from fcc_qp_b200.logdata import load_walking_log
from struct_proto import ldl_nopiv, ldl_solve, project_cone

class Struct2:
    """reduced KKT: order [R, y, D0]; D+ eliminated; sets detected per QP"""
    def __init__(self,Q,A,shift,tau):
        n=Q.shape[0]; m=A.shape[0]
        off=Q-np.diag(np.diag(Q))
        sep=(np.abs(off).max(axis=1)==0)
        qd=np.diag(Q)+shift
        self.R=np.where(~sep)[0]
        self.d0=np.where(sep&(qd<=tau))[0]; self.dp=np.where(sep&(qd>tau))[0]
        self.hp=qd[self.dp]; h0=qd[self.d0]
        nr=len(self.R); self.nr=nr; self.n=n; self.m=m
        AR=A[:,self.R]; Ap=A[:,self.dp]; A0=A[:,self.d0]
        self.Ap=Ap
        C=(Ap/self.hp)@Ap.T
        n0=len(self.d0)
        K=np.zeros((nr+m+n0,)*2)
        K[:nr,:nr]=Q[np.ix_(self.R,self.R)]+shift*np.eye(nr)
        K[nr:nr+m,:nr]=AR; K[:nr,nr:nr+m]=AR.T
        K[nr:nr+m,nr:nr+m]=-C
        K[nr+m:,nr:nr+m]=A0.T; K[nr:nr+m,nr+m:]=A0
        K[nr+m:,nr+m:]=np.diag(h0)
        self.K=K
        self.L,self.d=ldl_nopiv(K)
    def solve(self,rx,ry):
        nr,m=self.nr,self.m
        rr=np.concatenate([rx[self.R], ry-self.Ap@(rx[self.dp]/self.hp), rx[self.d0]])
        s=ldl_solve(self.L,self.d,rr)
        x=np.zeros(self.n)
        x[self.R]=s[:nr]; y=s[nr:nr+m]; x[self.d0]=s[nr+m:]
        x[self.dp]=(rx[self.dp]-self.Ap.T@y)/self.hp
        return x

def solve_qp(Q,b,A,beq,mu,lb,ub,nc,lcs,rho=5e-5,eps=1e-6,max_iter=100,tau=0.0,cls=Struct2,tau1=None):
    n=len(b)
    S0=cls(Q,A,0.0,tau)
    x=S0.solve(-b,beq)
    x0=x.copy()
    xbar=x.copy(); lcbar=x[lcs:lcs+nc].copy()
    mux=np.zeros(n); muc=np.zeros(nc)
    S1=None
    n_iter=max_iter
    for it in range(max_iter):
        if it>0:
            if S1 is None: S1=cls(Q,A,rho,tau if tau1 is None else tau1)
            w=xbar-mux; w[lcs:lcs+nc]=lcbar-muc
            x=S1.solve(-b+rho*w,beq)
        xbar=np.maximum(np.minimum(x+mux,ub),lb)
        for c in range(nc//3):
            lcbar[3*c:3*c+3]=project_cone(x[lcs+3*c:lcs+3*c+3]+muc[3*c:3*c+3],mu[c])
        rx=x-xbar; rc=x[lcs:lcs+nc]-lcbar
        mux+=rx; muc+=rc
        if np.abs(rc).max()<eps and np.abs(rx).max()<eps:
            n_iter=it;break
    return x,n_iter,x0,S0

if __name__=="__main__":
    qp=load_walking_log()
    g=np.load('/root/repo/tests/golden/walking_cold.npz')
    tau=float(sys.argv[1]) if len(sys.argv)>1 else 0.0
    step=int(sys.argv[2]) if len(sys.argv)>2 else 10
    errs=[];mis=0;sizes={}
    for i in range(0,qp.batch,step):
        x,ni,x0,S0=solve_qp(qp.Q[i],qp.b[i],qp.A_eq[i],qp.b_eq[i],qp.friction_coeffs[i],qp.lb[i],qp.ub[i],12,38,tau=tau)
        key=(S0.nr,len(S0.dp),len(S0.d0)); sizes[key]=sizes.get(key,0)+1
        e=np.abs(x-g['z'][i]).max()/max(1,np.abs(g['z'][i]).max())
        errs.append(e)
        if ni!=g['n_iter'][i]: mis+=1; print('mismatch',i,ni,g['n_iter'][i],e)
    errs=np.array(errs)
    print('sizes (nr,ndp,nd0):',sizes)
    print('max rel err',errs.max(),'p50',np.median(errs),'mismatch',mis,'of',len(errs))
