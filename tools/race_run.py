"""Small device-resident solves for compute-sanitizer (structure-exploiting kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200 import synthetic as syn, _native as nat
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
G = os.path.join(ROOT, "tests", "golden")
shape = sys.argv[1] if len(sys.argv) > 1 else "log"
if shape == "log":
    qp = load_walking_log().take(np.arange(300, 428)); gold = np.load(os.path.join(G, "walking_cold.npz")); gz = gold["z"][300:428]
else:
    qp = syn.make_batch(syn.SHAPES[shape], 96); gz = np.load(os.path.join(G, f"synthetic_{shape}_cold.npz"))["z"][:96]
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6)); s.structure = "probe"
for rep in range(3):
    s.Solve(*[torch.as_tensor(a, device="cuda:0") for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)])
    torch.cuda.synchronize()
    z = s.GetSolution().z.cpu().numpy()
    e = np.abs(z - gz).max(1) / np.maximum(1.0, np.abs(gz).max(1))
    print(shape, "rep", rep, "max err %.3e" % e.max(), "bad QPs", int((e > 1e-6).sum()), nat.last_struct_info(), flush=True)
