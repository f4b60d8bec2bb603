"""Time-parallel rollout of small sample counts (`rollout_tpar`, dynamic-window tick kernels with STASH = 2): when every CTA
owns <= 64 samples of one robot the horizon is split into noise / recurrence / cost / sum phases over all threads of the CTA.
It uses the same device functions as the serial rollout and adds a sample's stage costs in horizon order, so it must agree with
(a) the serial kernel on the same Philox stream (MPPI_TPAR=0 at create), (b) the C oracle fed the generated noise, for the
race-car class (mppi_race_car_obstacle.py:65-274, 200-entry window, footprint collisions), the obstacle-free race car
(mppi_race_car.py) and the diff-drive rules (mppi_differential_drive.py:87-165, _obs.py:228-313) run with a 200-entry window."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import Golden  # noqa: E402
from gpu_util import engine_from_spec  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402

U_ATOL = 2e-5       # updated nominal against the oracle's soft-min of FP64 costs (same bound as test_gpu_parity.py)


def _engine(sp, path, tpar):
    old = os.environ.get("MPPI_TPAR")
    os.environ["MPPI_TPAR"] = "1" if tpar else "0"
    try:
        return engine_from_spec(sp, path)
    finally:
        if old is None:
            del os.environ["MPPI_TPAR"]
        else:
            os.environ["MPPI_TPAR"] = old


def _spec(kind, K, T):
    if kind.startswith("racecar"):
        g = Golden("racecar_alpha0.9")
        sp = g.spec()
        sp.K, sp.T = K, T
        if kind == "racecar_noobs":
            sp.collision, sp.obstacles = "none", np.zeros((0, 3))
        x0 = np.asarray(g.rec["x0"][0], dtype=np.float64)
        return sp, g.path, x0
    g = Golden("diffdrive_pe0.05")
    obs = np.array([[0.9, 1.0, 0.3], [2.5, 1.6, 0.4]]) if kind == "diffdrive_obs" else None
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen", obstacles=obs, margin=1.2)
    sp.temperature = 2.0
    sp.window = 200                  # the race-car's SEARCH_IDX_LEN: the dynamic-window kernels
    return sp, g.path, np.array([0.3, 0.2, 0.4])


@pytest.mark.parametrize("kind,K,T", [("racecar", 16384, 50), ("racecar", 4096, 50), ("racecar", 1000, 33), ("racecar", 37, 50),
                                      ("racecar", 296 * 64, 20), ("racecar_noobs", 8000, 50), ("diffdrive", 16384, 30),
                                      ("diffdrive", 5000, 101), ("diffdrive_obs", 3000, 30)])
def test_time_parallel_rollout_matches_serial_kernel_and_oracle(kind, K, T):
    sp, path, x0 = _spec(kind, K, T)
    U0 = (np.random.default_rng(3).normal(0, 0.2, (T, 2)) * np.asarray(sp.u_max)[None, :] * 0.5).astype(np.float32)
    out = {}
    for tpar in (False, True):
        eng = _engine(sp, path, tpar)
        eng.set_keep_costs(True)
        eng.set_nominal(U0)
        eng.set_waypoint_idx(0)
        u0, u = eng.step(x0, None, seed=11, tick=4)
        n_top = min(K, 512)
        d_traj = torch.zeros(n_top, T, 4 if kind.startswith("racecar") else 3, device="cuda")
        d_cost = torch.zeros(n_top, device="cuda")
        d_idx = torch.zeros(n_top, dtype=torch.int32, device="cuda")
        eng.top_trajectories(x0, d_traj, n_top, d_idx=d_idx, d_cost=d_cost, seed=11, tick=4)
        out[tpar] = dict(u0=u0, u=u, st=eng.stats(), cost=d_cost.cpu().numpy(), idx=d_idx.cpu().numpy(), n_ctas=None)
        eng.set_trace(True)
        eng.set_nominal(U0)
        eng.set_waypoint_idx(0)
        eng.step(x0, None, seed=11, tick=4)
        out[tpar]["n_ctas"] = len(eng.trace()[0])
        if tpar:
            eps = torch.zeros(K, T, 2, device="cuda")
            eng.generate_noise(eps, seed=11, tick=4)
            assert eng.check_guards() == 0
        eng.close()
    a, b = out[False], out[True]
    # the time-parallel grid really ran: CTAs of <= 64 samples (the serial kernel's own at least 128)
    assert (K + b["n_ctas"] - 1) // b["n_ctas"] <= 64 and (K <= 128 or b["n_ctas"] > a["n_ctas"]), (a["n_ctas"], b["n_ctas"])
    # (a) against the serial kernel: same samples, same cost terms, same horizon-order sum
    assert np.array_equal(a["idx"], b["idx"]) or np.mean(a["idx"] != b["idx"]) < 0.02      # equal-cost neighbours may swap
    assert np.allclose(a["cost"], b["cost"], rtol=2e-6, atol=1e-6), np.max(np.abs(a["cost"] - b["cost"]) / np.abs(a["cost"]))
    assert a["st"]["idx"] == b["st"]["idx"] and a["st"]["min_collisions"] == b["st"]["min_collisions"]
    assert abs(a["st"]["rho"] - b["st"]["rho"]) <= 2e-6 * abs(a["st"]["rho"]) + 1e-6
    assert abs(a["st"]["eta"] - b["st"]["eta"]) <= 1e-4 * a["st"]["eta"]
    assert np.max(np.abs(a["u"] - b["u"])) <= 2e-6, np.max(np.abs(a["u"] - b["u"]))
    # (b) against the oracle on the generated noise
    o = co.tick(sp, path, U0.astype(np.float64), 0, x0, eps.cpu().numpy())
    tol = 1e-4 if kind.startswith("racecar") else U_ATOL       # FP32 class, controls of O(1): as in test_gpu_fullsize.py
    assert np.max(np.abs(b["u"] - o["U_after"])) <= tol, np.max(np.abs(b["u"] - o["U_after"]))


def test_time_parallel_rollout_is_bit_reproducible_and_closed_loop_runs_on_it():
    sp, path, x0 = _spec("racecar", 16384, 50)
    eng = _engine(sp, path, True)
    U0 = np.zeros((50, 2), dtype=np.float32)
    ref = None
    for rep in range(20):
        eng.set_nominal(U0)
        eng.set_waypoint_idx(0)
        _, u = eng.step(x0, None, seed=2, tick=1)
        if ref is None:
            ref = u.copy()
        assert np.array_equal(u, ref), rep
    # graph-captured closed loop (plant step in the last CTA) on the same kernel against the serial one
    eng.set_nominal(U0)
    eng.set_waypoint_idx(0)
    st_t, ct_t = eng.run_closed_loop(x0, 12, seed=5, tick0=0, plant=1)
    assert eng.check_guards() == 0
    eng.close()
    eng = _engine(sp, path, False)
    st_s, ct_s = eng.run_closed_loop(x0, 12, seed=5, tick0=0, plant=1)
    eng.close()
    assert np.max(np.abs(ct_t - ct_s)) <= 1e-4 and np.max(np.abs(st_t - st_s)) <= 1e-3, (np.max(np.abs(ct_t - ct_s)), np.max(np.abs(st_t - st_s)))
