import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
from struct2 import *
from fcc_qp_b200 import synthetic
import inspect

class Struct3(Struct2):
    refine=True
    def solve(self,rx,ry):
        nr,m=self.nr,self.m
        rr=np.concatenate([rx[self.R], ry-self.Ap@(rx[self.dp]/self.hp), rx[self.d0]])
        s=ldl_solve(self.L,self.d,rr)
        if self.refine:
            s=s+ldl_solve(self.L,self.d,rr-self.K@s)
        x=np.zeros(self.n)
        x[self.R]=s[:nr]; y=s[nr:nr+m]; x[self.d0]=s[nr+m:]
        x[self.dp]=(rx[self.dp]-self.Ap.T@y)/self.hp
        return x

def run(qp,g,lcs,nc,refine0,refine1,step=1,opts=None):
    rho,eps,mi=5e-5,1e-6,100
    if opts is not None:
        mi=int(opts[0]);rho=float(opts[1]);eps=float(opts[2])
    errs=[];mis=[]
    class P0(Struct3): refine=refine0
    class P1(Struct3): refine=refine1
    for i in range(0,qp.batch,step):
        Q,A,b,beq=qp.Q[i],qp.A_eq[i],qp.b[i],qp.b_eq[i]
        n=len(b)
        S0=P0(Q,A,0.0,0.0)
        x=S0.solve(-b,beq)
        xbar=x.copy(); lcbar=x[lcs:lcs+nc].copy(); mux=np.zeros(n); muc=np.zeros(nc); S1=None; n_iter=mi
        for it in range(mi):
            if it>0:
                if S1 is None: S1=P1(Q,A,rho,0.0)
                w=xbar-mux; w[lcs:lcs+nc]=lcbar-muc
                x=S1.solve(-b+rho*w,beq)
            xbar=np.maximum(np.minimum(x+mux,qp.ub[i]),qp.lb[i])
            for c in range(nc//3):
                lcbar[3*c:3*c+3]=project_cone(x[lcs+3*c:lcs+3*c+3]+muc[3*c:3*c+3],qp.friction_coeffs[i][c])
            rx=x-xbar; rc=x[lcs:lcs+nc]-lcbar
            mux+=rx; muc+=rc
            if (np.abs(rc).max() if nc else 0)<eps and np.abs(rx).max()<eps:
                n_iter=it;break
        e=np.abs(x-g['z'][i]).max()/max(1,np.abs(g['z'][i]).max())
        errs.append(e)
        if n_iter!=g['n_iter'][i]: mis.append((i,n_iter,int(g['n_iter'][i])))
    errs=np.array(errs)
    return errs.max(),np.median(errs),mis,len(errs)

if __name__=="__main__":
    which=sys.argv[1]; r0=int(sys.argv[2]); r1=int(sys.argv[3]); step=int(sys.argv[4])
    if which=='log':
        qp=load_walking_log(); g=np.load('/root/repo/tests/golden/walking_cold.npz'); lcs,nc=38,12
        print(which,r0,r1,run(qp,g,lcs,nc,r0,r1,step))
    else:
        shp={'humanoid':synthetic.HUMANOID,'quadruped':synthetic.QUADRUPED,'multicontact':synthetic.MULTICONTACT}[which]
        B={'humanoid':192,'quadruped':192,'multicontact':96}[which]
        qp=synthetic.make_batch(shp,B); g=np.load(f'/root/repo/tests/golden/synthetic_{which}_cold.npz')
        print(which,r0,r1,run(qp,g,shp.lambda_c_start,shp.nc,r0,r1,step))
