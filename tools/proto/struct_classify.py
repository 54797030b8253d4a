import numpy as np, sys
sys.path.insert(0,'/root/repo')
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200 import synthetic
def classify(Q,A):
    n=Q.shape[0]
    off=Q-np.diag(np.diag(Q))
    sep=np.abs(off).max(axis=1)==0
    q=np.diag(Q)
    nnz=(A!=0).sum(axis=0)
    R=np.where(~sep)[0]
    d0=np.where(sep&(q==0))[0]
    cand1=np.where(sep&(q>0)&(nnz<=1))[0]
    rows={}
    for j in cand1:
        if nnz[j]==1:
            k=int(np.nonzero(A[:,j])[0][0]); rows.setdefault(k,[]).append(j)
    conflict=[j for k,v in rows.items() if len(v)>1 for j in v]
    d1=[j for j in cand1 if j not in conflict]
    dp=[j for j in np.where(sep&(q>0))[0] if j not in d1]
    empty0=[j for j in d0 if nnz[j]==0]
    return len(R),len(dp),len(d1),len(d0),len(conflict),len(empty0),tuple(R)
qp=load_walking_log()
from collections import Counter
c=Counter()
for i in range(qp.batch):
    c[classify(qp.Q[i],qp.A_eq[i])[:6]]+=1
print('log (nr,ndp,nd1,nd0,conflict,empty0):',c)
print('R example', classify(qp.Q[0],qp.A_eq[0])[6], classify(qp.Q[300],qp.A_eq[300])[6])
for shp in (synthetic.HUMANOID,synthetic.QUADRUPED,synthetic.MULTICONTACT):
    b=synthetic.make_batch(shp,64); c=Counter()
    for i in range(64): c[classify(b.Q[i],b.A_eq[i])[:6]]+=1
    print(shp.name,c)
