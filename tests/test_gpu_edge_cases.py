"""Edge cases of the tick through the C ABI against the C oracle: ragged sizes (K not a multiple of the block,
odd horizons), windows truncated at the end of the path, explore/exploit boundaries, gamma != 0, many / zero
obstacles, tiny K, long horizons that do not fit the shared-memory noise stash."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import Golden  # noqa: E402
from gpu_util import cost_mismatch, engine_from_spec  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402


def _check(sp, path, x0, idx0, ticks=2, seed=17, injected=False, u_atol=2e-5):
    eng = engine_from_spec(sp, path)
    eng.set_waypoint_idx(idx0)
    eps = torch.zeros(sp.K, sp.T, 2, dtype=torch.float32, device="cuda")
    S = torch.zeros(sp.K, dtype=torch.float32, device="cuda")
    U, idx = np.zeros((sp.T, 2)), idx0
    for tick in range(ticks):
        eng.generate_noise(eps, seed=seed, tick=tick)
        e = eps.cpu().numpy()
        o = co.tick(sp, path, U, idx, np.asarray(x0, np.float32).astype(np.float64), e)
        eng.set_nominal(U.astype(np.float32)); eng.set_waypoint_idx(idx)
        eng.rollout_costs(x0, S, eps if injected else None, seed=seed, tick=tick)
        Sg = S.cpu().numpy()
        assert np.array_equal(np.round(Sg / 1e10), np.round(o["S"] / 1e10))
        frac, worst = cost_mismatch(Sg, o["S"], rtol=2e-5)
        assert frac == 0.0, (frac, worst)
        eng.set_waypoint_idx(idx)
        u0, useq = eng.step(x0, eps if injected else None, seed=seed, tick=tick)
        assert np.max(np.abs(useq - o["U_after"])) <= u_atol, np.max(np.abs(useq - o["U_after"]))
        assert eng.get_waypoint_idx() == o["idx_after"]
        U, idx = useq.astype(np.float64), o["idx_after"]
    eng.close()


@pytest.mark.parametrize("K,T", [(1, 10), (7, 11), (255, 13), (257, 30), (1000, 31), (3001, 50), (640, 77), (512, 128)])
@pytest.mark.parametrize("cost_mode", ["sum", "last"])
def test_ragged_sizes_diffdrive(K, T, cost_mode):
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode=cost_mode, waypoint_mode="frozen")
    sp.temperature = 1.0
    _check(sp, g.path, np.array([0.3, 0.2, 0.4]), 2)
    _check(sp, g.path, np.array([0.3, 0.2, 0.4]), 2, injected=True)


@pytest.mark.parametrize("idx0", [150, 160, 166, 167])
def test_window_truncated_at_path_end(idx0):
    """The slice path[idx:idx+20] runs off the end of the 168-point path (Q4, Q10)."""
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=300, T=20, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 1.0
    x0 = np.append(g.path[idx0, :2] + 0.05, g.path[idx0, 2])
    _check(sp, g.path, x0, idx0)


@pytest.mark.parametrize("pe", [0.0, 1e-4, 0.25, 0.999, 1.0])
def test_explore_exploit_split_boundaries(pe):
    """Q6: k < (1 - param_exploration) * K in Python floats, including all-exploit and all-explore."""
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=400, T=15, param_exploration=max(pe, 1e-9), cost_mode="sum", waypoint_mode="frozen")
    sp.param_exploration = pe
    sp.temperature = 0.7
    eng_U = np.random.default_rng(0).normal(0, 0.3, (15, 2))
    eng = engine_from_spec(sp, g.path)
    eps = torch.zeros(sp.K, sp.T, 2, dtype=torch.float32, device="cuda")
    eng.generate_noise(eps, seed=2, tick=0)
    eng.set_nominal(eng_U.astype(np.float32))
    u0, useq = eng.step(np.zeros(3), None, seed=2, tick=0)
    o = co.tick(sp, g.path, eng_U.astype(np.float32).astype(np.float64), 0, np.zeros(3), eps.cpu().numpy())
    assert np.max(np.abs(useq - o["U_after"])) <= 2e-5
    eng.close()


@pytest.mark.parametrize("alpha", [0.0, 0.5, 0.98])
def test_gamma_term_with_correlated_noise(alpha):
    """gamma = lambda (1 - alpha) != 0 with a full (non-diagonal) Sigma: u^T Sigma^-1 v and the Cholesky factor."""
    g = Golden("diffdrive_pe0.05")
    sigma = np.array([[0.2, 0.03], [0.03, 0.02]])
    sp = orc.diffdrive_spec(K=500, T=20, param_exploration=0.1, param_lambda=3.0, param_alpha=alpha, sigma=sigma,
                            cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    U0 = np.random.default_rng(1).normal(0, 0.5, (20, 2)).astype(np.float32)
    eps = torch.zeros(sp.K, sp.T, 2, dtype=torch.float32, device="cuda")
    eng.generate_noise(eps, seed=8, tick=4)
    e = eps.cpu().numpy().astype(np.float64).reshape(-1, 2)
    assert np.max(np.abs(np.cov(e.T) - sigma)) < 0.02
    eng.set_nominal(U0)
    u0, useq = eng.step(np.array([0.1, 0.0, 0.2]), None, seed=8, tick=4)
    o = co.tick(sp, g.path, U0.astype(np.float64), 0, np.array([0.1, 0.0, 0.2]), eps.cpu().numpy())
    assert np.max(np.abs(useq - o["U_after"])) <= 2e-5
    eng.close()


@pytest.mark.parametrize("n_obs", [0, 1, 5, 16])
def test_racecar_obstacle_counts(n_obs):
    g = Golden("racecar_default")
    rng = np.random.default_rng(n_obs)
    obs = np.column_stack([rng.uniform(-12, 12, n_obs), rng.uniform(-6, 6, n_obs), rng.uniform(0.3, 1.5, n_obs)]) if n_obs else None
    sp = orc.racecar_spec(K=700, T=25, obstacles=obs, dtype=np.float64)
    x0 = g.path[3].astype(np.float64) + np.array([0.2, -0.1, 0.02, 0.3])
    _check(sp, g.path, x0, 1, u_atol=5e-5)


def test_diffdrive_circle_obstacles_sum_mode():
    g = Golden("diffdrive_obs")
    sp = g.spec(cost_mode="sum", waypoint_mode="frozen")
    sp.K, sp.T = 900, 25
    _check(sp, g.path, np.array([1.6, 1.5, 0.6]), 20)


def test_bad_configurations_are_rejected():
    from mppi_b200 import MppiError
    g = Golden("diffdrive_pe0.05")
    for kw in (dict(T=4), dict(T=129), dict(K=0)):
        sp = orc.diffdrive_spec(K=kw.get("K", 64), T=kw.get("T", 12), cost_mode="sum", waypoint_mode="frozen")
        with pytest.raises(MppiError):
            engine_from_spec(sp, g.path)
    sp = orc.diffdrive_spec(K=64, T=12, cost_mode="sum", waypoint_mode="frozen")
    from mppi_b200.engine import MPPIEngine
    eng = MPPIEngine(model="diffdrive", K=64, T=12, dt=0.1, u_max=(5, 3), sigma=np.diag([0.1, 0.01]), stage_w=[5, 5, 10],
                     term_w=[5, 5, 10], param_exploration=0.05, param_lambda=1.0, param_alpha=0.2, temperature=1.0,
                     window=20, cost_mode="sum", waypoint_mode="frozen", filter_kind="diffdrive", yaw_wrap=False)
    with pytest.raises(MppiError):                      # stepping before a reference path is set
        eng.step(np.zeros(3))
    eng.close()
