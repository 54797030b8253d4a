"""Convert the reference's pickled walking log into the compact fixture.

Run in the build container only (needs /root/reference):
    python tests/golden/make_walking_log.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from fcc_qp_b200.logdata import stack_reference_log, save_compact, load_compact

src = "/root/reference/test_data/id_qp_log_walking.npz"
dst = os.path.join(os.path.dirname(__file__), "walking_log_compact.npz")
full = stack_reference_log(src)
save_compact(full, dst)
back = load_compact(dst)
for k in ("Q", "b", "A_eq", "b_eq", "friction_coeffs", "lb", "ub"):
    assert np.array_equal(getattr(full, k), getattr(back, k)), k
print("ok", full.batch, os.path.getsize(dst) / 1e6, "MB")
