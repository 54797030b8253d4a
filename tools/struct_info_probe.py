import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
from fcc_qp_b200 import _native as nat
qp = load_walking_log().tile(65536)
dev = torch.device("cuda:0")
args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6)); s.time_kernel = False
for i in range(4):
    s.Solve(*args)
    print(i, s._caps, nat.last_launch_info(), nat.last_struct_info())
torch.cuda.synchronize()
print(nat.last_struct_info())
