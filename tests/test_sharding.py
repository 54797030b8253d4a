"""Multi-GPU plumbing on CPU: world_size-2 gloo job.  The path has no data-path collective
(SURVEY.md 8e); ranks own contiguous batch ranges and only time / results are reduced."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
from fcc_qp_b200 import sharding
rank, local_rank, world = sharding.init_process_group("gloo")
assert world == 2
B = 101
lo, hi = sharding.shard_range(B, rank, world)
local = np.arange(lo, hi, dtype=np.float64)[:, None] * np.ones((1, 3))   # stand-in for a solved shard
allrows = sharding.gather_rows(local, B)
assert allrows.shape == (B, 3) and np.array_equal(allrows[:, 0], np.arange(B))
t = sharding.max_over_ranks(1.0 + rank)
assert t == 2.0
s = sharding.sum_over_ranks(float(hi - lo))
assert s == B
sharding.barrier()
print("rank", rank, "ok")
"""


def test_shard_range_partitions():
    from fcc_qp_b200.sharding import shard_range
    for B in (0, 1, 7, 65536, (1 << 20) + 3):
        for W in (1, 2, 4, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
