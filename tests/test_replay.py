"""Headless replay CLI (fcc_qp_b200.replay) -- the stand-in for the reference's plotting script
fcc_qp_test.py:72-91.  CPU tests cover the log handling and the summary arithmetic (fed with the
compiled reference's golden outputs); GPU tests run both modes and compare with the goldens."""
import json
import os

import numpy as np
import pytest

from fcc_qp_b200 import replay
from fcc_qp_b200.logdata import load_compact

G = os.path.join(os.path.dirname(__file__), "golden")
OPTS = dict(max_iter=100, rho=5e-5, eps_fcone=1e-6, eps_bound=1e-6)


def test_describe_and_convert_roundtrip(tmp_path, walking_log, capsys):
    assert replay.main(["--describe"]) == 0
    d = json.loads(capsys.readouterr().out)
    assert (d["qps"], d["num_vars"], d["num_equality_constraints"], d["nc"], d["lambda_c_start"]) == (2019, 60, 38, 12, 38)
    assert d["bounded_variables"] == 10                      # only the torques are bounded
    out = tmp_path / "c.npz"
    assert replay.main(["--convert", str(out)]) == 0
    capsys.readouterr()
    again = load_compact(str(out))
    for k in ("Q", "b", "A_eq", "b_eq", "friction_coeffs", "lb", "ub"):
        assert np.array_equal(getattr(again, k), getattr(walking_log, k)), k
    assert replay.main(["--log", str(out), "--describe"]) == 0


def test_reference_format_log_is_accepted(tmp_path, walking_log, capsys):
    """The reference's on-disk format: pickled object array of dicts under key 'qps' (fcc_qp_test.py:22-24)."""
    qps = np.empty(5, dtype=object)
    for i in range(5):
        qps[i] = walking_log.qp(i)
    p = tmp_path / "id_qp_log_tiny.npz"
    np.savez(p, qps=qps)
    assert replay.main(["--log", str(p), "--describe"]) == 0
    d = json.loads(capsys.readouterr().out)
    assert d["qps"] == 5 and d["num_vars"] == 60
    got = replay.load_log(str(p), 12, 38)
    assert np.array_equal(got.Q, walking_log.Q[:5]) and np.array_equal(got.A_eq, walking_log.A_eq[:5])


def test_summary_of_golden_solutions(walking_log):
    """summarize() on the compiled reference's own warm-sequential outputs: the numbers the reference plots."""
    g = np.load(os.path.join(G, "walking_warm.npz"))
    s = replay.summarize(walking_log, g["z"], g["n_iter"], g["status"], g["res_bounds"],
                         g["res_fcone"], g["bounds_viol"], g["fcone_viol"], 100)
    h = s["iterations"]["histogram"]
    assert sum(h.values()) == 2019 and h["0"] == 1810 and h["100"] == 39          # SURVEY 8c oracle stats
    assert s["iterations"]["hit_max_iter"] == 39 and s["status_counts"]["1"] == 39
    assert s["equality_residual_inf"] < 1e-6                                     # A x = b_eq at every exit
    assert s["bounds_viol"]["max"] == 0.0 and s["friction_cone_viol"]["max"] < 7e-3
    assert set(s["slices"]) == {"vdot", "u", "lambda_h", "lambda_c"}
    assert -300.0 - 1e-6 <= s["slices"]["u"]["min"] and s["slices"]["u"]["max"] <= 300.0 + 1e-6


def test_replay_needs_the_gpu(monkeypatch):
    """No CPU solve path: without a CUDA device the replay raises instead of producing numbers."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(Exception):
        replay.main(["--mode", "batch", "--limit", "4"])


@pytest.mark.gpu
def test_batch_mode_matches_cold_goldens(tmp_path, walking_log):
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    js, sol = tmp_path / "s.json", tmp_path / "z.npz"
    assert replay.main(["--mode", "batch", "--json", str(js), "--save-solutions", str(sol)]) == 0
    s = json.loads(js.read_text())
    z = np.load(sol)
    err = np.abs(z["z"] - gold["z"]).max(1) / np.maximum(1.0, np.abs(gold["z"]).max(1))
    assert err.max() <= 1e-6
    assert np.array_equal(z["n_iter"], gold["n_iter"])
    assert s["iterations"]["histogram"] == {"0": 1978, "6": 14, "9": 1, "100": 26}
    assert s["mode"] == "batch" and s["qps"] == 2019 and s["device_time_s"] > 0


@pytest.mark.gpu
def test_sequential_mode_matches_warm_goldens(tmp_path, walking_log):
    gold = np.load(os.path.join(G, "walking_warm.npz"))
    js, sol = tmp_path / "s.json", tmp_path / "z.npz"
    N = 400
    assert replay.main(["--limit", str(N), "--json", str(js), "--save-solutions", str(sol)]) == 0
    s = json.loads(js.read_text())
    z = np.load(sol)
    err = np.abs(z["z"] - gold["z"][:N]).max(1) / np.maximum(1.0, np.abs(gold["z"][:N]).max(1))
    assert err.max() <= 1e-6
    assert np.array_equal(z["n_iter"], gold["n_iter"][:N])
    assert s["warm_start"] is True and s["qps"] == N and s["solve_time_s"]["p50"] > 0
