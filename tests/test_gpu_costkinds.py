"""SURVEY.md 8f row 3 on the GPU: the goal-point diff-drive MPPI (test/mppi_differential_drive_obs.py) and the
moving-soft-obstacle running cost (test/test_mppi_diff_obs.py), both through the C ABI.  The per-tick golden
checks of the goal class run with every other class in test_gpu_parity.py (it is in ALL_CASES); here: the
drop-in class surface, the Philox / sum / sharded-shape variants against the C oracle, and the soft-obstacle cost
against what the reference's own functions produced."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import TARGET_SOFT_CASE, Golden, rel_err  # noqa: E402
from gpu_util import cost_mismatch, engine_from_spec  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402

COST_RTOL = 1e-5
U_ATOL = 2e-5


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def test_goal_dropin_class_matches_reference_ticks():
    """Same constructor kwargs and 4-tuple as test/mppi_differential_drive_obs.py:MPPIAlgorithms; every recorded
    tick of the reference class is reproduced (nominal after the shift, returned u0 = post-shift row 0)."""
    from mppi_b200.mppi_differential_drive_goal import MPPIAlgorithms
    g = Golden("diffdrive_goal")
    m = g.meta
    ctrl = MPPIAlgorithms(delta_t=m["delta_t"], goal_point=np.array(m["goal"]), max_speed=m["max_speed"],
                          max_omega=m["max_omega"], num_samples_K=m["num_samples_K"], num_horizons_T=m["num_horizons_T"],
                          param_exploration=m["param_exploration"], param_lambda=m["param_lambda"],
                          param_alpha=m["param_alpha"], sigma=np.array([[0.1, 0.0], [0.0, 0.01]]),
                          stage_cost_weight=10 * np.array([5.0, 9.0]), terminal_cost_weight=10 * np.array([5.0, 9.0]),
                          obstacle_circles=g.obstacles, safety_margin_rate=m["safety_margin_rate"],
                          visualize_optimal_traj=False, visualze_sampled_trajs=False)
    assert ctrl.ref_path is None and np.array_equal(ctrl.goal_point, m["goal"])
    for i in range(g.n_ticks):
        ctrl.u_prev = g.rec["U0"][i]
        u0, useq, opt, samp = ctrl._calc_input_control(g.rec["x0"][i], noise=g.eps[i])
        assert useq.dtype == np.float64 and opt.shape == (ctrl.T, 3) and samp.shape == (ctrl.K, ctrl.T, 3)
        assert np.max(np.abs(useq - g.rec["U_after"][i])) <= U_ATOL, i
        assert np.max(np.abs(u0 - g.rec["u0"][i])) <= U_ATOL
        assert useq is ctrl.u_prev
    # moving the goal re-targets the same handle
    ctrl.goal_point = [1.0, -2.0]
    sp = g.spec()
    sp.goal = np.array([1.0, -2.0, 0.0])
    o = co.tick(sp, None, g.rec["U0"][1], 0, g.rec["x0"][1], g.eps[1])
    ctrl.u_prev = g.rec["U0"][1]
    _, useq, _, _ = ctrl._calc_input_control(g.rec["x0"][1], noise=g.eps[1])
    assert np.max(np.abs(useq - o["U_after"])) <= U_ATOL


@pytest.mark.parametrize("cost_mode", ["last", "sum"])
def test_goal_philox_large_K_matches_oracle(cost_mode):
    """K = 65 536 with in-kernel Philox noise: export the exact noise, feed it to the FP64 C oracle."""
    K, T = 65536, 30
    sp = orc.goal_spec(K, T, [5.0, 5.0], cost_mode=cost_mode)
    sp.temperature = 5.0
    eng = engine_from_spec(sp, None)
    U = (np.random.default_rng(3).normal(0, 0.4, (T, 2))).astype(np.float32)
    x0 = np.array([2.6, 1.4, 0.5])
    eps = torch.zeros(K, T, 2, dtype=torch.float32, device="cuda")
    eng.generate_noise(eps, seed=21, tick=4)
    S = torch.zeros(K, dtype=torch.float32, device="cuda")
    eng.set_nominal(U)
    eng.rollout_costs(x0, S, None, seed=21, tick=4)
    e = eps.cpu().numpy()
    So, _, _ = co.costs(sp, None, U, 0, x0, e)
    Sg = S.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.round(Sg / 1e10) > 0, np.round(So / 1e10) > 0) or \
        np.mean((np.round(Sg / 1e10)) != np.round(So / 1e10)) < 2e-4      # circle-boundary near ties
    same = np.round(Sg / 1e10) == np.round(So / 1e10)
    frac, worst = cost_mismatch(Sg[same], So[same], rtol=COST_RTOL)
    assert frac == 0.0, (frac, worst)
    assert (So >= 1e10).any() and (So < 1e10).any()                       # the fixture exercises both branches
    eng.set_nominal(U)
    u0, useq = eng.step(x0, None, seed=21, tick=4)
    o = co.update(sp, None, U, np.where(same, So, Sg), e)
    assert np.max(np.abs(useq - o["U_after"])) <= 5e-5
    eng.close()


@pytest.mark.parametrize("cost_mode", ["last", "sum"])
def test_goal_closed_loop_reaches_goal(cost_mode):
    """On-device closed loop (plant = DifferentialDrive.update_state, test/mppi_differential_drive_obs.py:33-40)
    with the script's own parameters: the robot ends next to the goal.  In the literal `last` mode only the final
    state of a rollout is tested for collision (the stage cost is assigned, not accumulated, :121), so the plant may
    clip an inflated obstacle on the way; in `sum` mode every step counts and it never enters one."""
    sp = orc.goal_spec(4096, 20, [5.0, 5.0], cost_mode=cost_mode)
    if cost_mode == "sum":
        sp.temperature = 2.0
    eng = engine_from_spec(sp, None)
    states, controls = eng.run_closed_loop(np.array([0.0, 0.0, 0.0]), 250, seed=5, tick0=0, plant=0)
    d_goal = np.hypot(states[:, 0] - 5.0, states[:, 1] - 5.0)
    assert d_goal.min() < 0.3, d_goal.min()
    if cost_mode == "sum":
        rr = sp.robot_radius * sp.margin
        for ox, oy, orad in sp.obstacles:
            assert np.all(np.hypot(states[:, 0] - ox, states[:, 1] - oy) >= rr + orad - 1e-3)
    assert np.all(np.abs(controls[:, 0]) <= 10.0 + 1e-5) and np.all(np.isfinite(states))
    eng.close()


def test_target_soft_costs_match_reference_functions():
    """Per-sample rollout costs vs the reference's `dynamics` + `running_cost` executed by its own
    `_compute_rollout_costs` loop (float32 torch; golden made by tests/golden/make_golden.py extras)."""
    g = Golden(TARGET_SOFT_CASE)
    sp = g.spec()
    eng = engine_from_spec(sp, None)
    S = torch.zeros(sp.K, dtype=torch.float32, device="cuda")
    for i in range(g.n_ticks):
        eng.set_nominal(g.rec["U0"][i])
        eng.rollout_costs(g.rec["x0"][i], S, _dev(g.eps[i]))
        frac, worst = cost_mismatch(S.cpu().numpy(), g.rec["S"][i], rtol=COST_RTOL)
        assert frac == 0.0, (i, frac, worst)
    eng.close()


def test_target_soft_full_tick_and_moving_obstacles():
    """Full tick vs the C oracle; then the obstacle motion matters: freezing the obstacles changes the costs."""
    g = Golden(TARGET_SOFT_CASE)
    sp = g.spec()
    eng = engine_from_spec(sp, None)
    for i in range(g.n_ticks):
        o = co.tick(sp, None, g.rec["U0"][i], 0, g.rec["x0"][i], g.eps[i])
        eng.set_nominal(g.rec["U0"][i])
        u0, useq = eng.step(g.rec["x0"][i], _dev(g.eps[i]))
        assert np.max(np.abs(useq - o["U_after"])) <= U_ATOL, (i, np.max(np.abs(useq - o["U_after"])))
    S = torch.zeros(sp.K, dtype=torch.float32, device="cuda")
    eng.set_nominal(g.rec["U0"][1])
    eng.rollout_costs(g.rec["x0"][1], S, _dev(g.eps[1]))
    moving = S.cpu().numpy().copy()
    eng.set_moving_obstacles(sp.obstacles, 20.0 * sp.obs_vel)
    eng.rollout_costs(g.rec["x0"][1], S, _dev(g.eps[1]))
    sp_fast = g.spec()
    sp_fast.obs_vel = 20.0 * sp.obs_vel
    So, _, _ = co.costs(sp_fast, None, g.rec["U0"][1], 0, g.rec["x0"][1], g.eps[1])
    assert rel_err(S.cpu().numpy(), So) <= 2e-5
    assert np.max(np.abs(S.cpu().numpy() - moving)) > 1.0
    eng.close()


def test_cost_kind_argument_checks():
    """Error behaviour of the new entry points: wrong kind / wrong model fail loudly, never fall back."""
    from mppi_b200 import MppiError
    g = Golden("diffdrive_goal")
    eng = engine_from_spec(g.spec(), None)
    with pytest.raises(MppiError):
        eng.set_ref_path(np.zeros((4, 3)))                 # a goal handle takes no path
    with pytest.raises(MppiError):
        eng.set_moving_obstacles(np.zeros((1, 2)), np.zeros((1, 2)))
    eng.close()
    path_eng = engine_from_spec(Golden("diffdrive_pe0.05").spec(), Golden("diffdrive_pe0.05").path)
    with pytest.raises(MppiError):
        path_eng.set_goal([1.0, 2.0])
    path_eng.close()
    sp = orc.racecar_spec(64, 20)
    sp.cost_kind = "goal"
    with pytest.raises(MppiError):
        engine_from_spec(sp, None)                         # goal cost is a diff-drive cost


@pytest.mark.parametrize("name", ["diffdrive_viz", "racecar_viz"])
def test_top_n_trajectories_match_reference_replays(name):
    """the n lowest-cost samples in np.argsort(S) order.  The golden holds the reference class's
    own (K,T,nx) replay array (t-1 indexing) and its costs, so row i of the top-N output must be the reference's
    replay of sample argsort(S)[i]."""
    from golden_util import Golden as G
    g = G(name)
    sp = g.spec()
    eng = engine_from_spec(sp, g.path, clamp_nominal=True)
    eng.set_keep_costs(True)
    n_top = 7
    nx = sp.nx
    for i in range(g.n_ticks):
        eng.set_nominal(g.rec["U0"][i])
        eng.set_waypoint_idx(int(g.rec["idx0"][i]))
        d_eps = _dev(g.eps[i])
        eng.step(g.rec["x0"][i], d_eps)
        traj = torch.zeros(n_top, sp.T, nx, dtype=torch.float32, device="cuda")
        idx = torch.zeros(n_top, dtype=torch.int32, device="cuda")
        cost = torch.zeros(n_top, dtype=torch.float32, device="cuda")
        opt = eng.top_trajectories(g.rec["x0"][i], traj, n_top, idx, cost, want_optimal=True, index_shift=1, d_eps=d_eps)
        Sref = g.rec["S"][i].astype(np.float64)
        order = np.argsort(Sref, kind="stable")[:n_top]
        got = idx.cpu().numpy()
        c = cost.cpu().numpy()
        assert np.all(np.diff(c) >= 0)
        # same samples unless two costs are closer than the FP32 cost tolerance
        gap_ok = np.abs(Sref[got] - Sref[order]) <= 2e-5 * np.abs(Sref[order]) + 1e-6
        assert np.all(gap_ok), (name, i, got, order)
        assert np.max(np.abs(c - Sref[got])) <= 2e-5 * np.max(np.abs(Sref[order])) + 1e-6
        assert np.max(np.abs(traj.cpu().numpy() - g.rec["sampled_traj"][i][got])) <= 2e-5, (name, i)
        assert np.max(np.abs(opt - g.rec["optimal_traj"][i])) <= 5e-5
    eng.close()


def test_dynamic_obstacle_controller_surface():
    """MPPIDynamicObstacles: command() = row 0 of the updated nominal before the shift; get_trajectories() = optimal
    rollout + the top max(10, K/10) samples by cost, controls indexed t (test/test_mppi_diff_obs.py:88-111)."""
    from mppi_b200.mppi_diff_obs_dynamic import MPPIDynamicObstacles
    g = Golden(TARGET_SOFT_CASE)
    sp = g.spec()
    m = g.meta
    ctrl = MPPIDynamicObstacles(noise_sigma=np.array(m["sigma"]), num_samples=m["K"], horizon=m["T"], lambda_=sp.temperature,
                                u_min=[-1.0, -1.0], u_max=[1.0, 1.0], delta_t=m["delta_t"], target=m["target"], Q=m["Q"],
                                R=m["R"], obstacle_positions=m["obs_pos"], obstacle_velocities=m["obs_vel"],
                                safety_distance=m["soft_sd"], obstacle_weight=m["soft_w"])
    i = 1
    ctrl.u_prev = g.rec["U0"][i]
    u = ctrl.command(g.rec["x0"][i], noise=g.eps[i])
    o = co.tick(sp, None, g.rec["U0"][i], 0, g.rec["x0"][i], g.eps[i])
    Upre = np.concatenate([o["U_after"][:1] * 0, o["U_after"][:-1]])        # U_after is U_pre shifted up by one row
    U_pre0 = g.rec["U0"][i][0] + (orc.filter_matrix(sp.T, "racecar") @ o["w_eps"])[0]
    assert np.max(np.abs(u - U_pre0)) <= U_ATOL
    assert np.max(np.abs(ctrl.u_prev - o["U_after"])) <= U_ATOL
    opt, samp = ctrl.get_trajectories()
    n_top = max(10, sp.K // 10)
    assert samp.shape == (n_top, sp.T, 3) and opt.shape == (sp.T, 3)
    order = np.argsort(o["S"], kind="stable")[:n_top]
    got = ctrl.last_top_idx
    assert np.all(np.abs(o["S"][got] - o["S"][order]) <= 2e-5 * np.abs(o["S"][order]) + 1e-6)
    V, X = orc.rollout_states(g.spec(dtype=np.float64), g.rec["U0"][i].astype(np.float64), g.rec["x0"][i], g.eps[i].astype(np.float64))
    assert np.max(np.abs(samp - X[got])) <= 2e-5
    # optimal rollout: the clamped updated nominal applied with index t
    Upre_full = np.vstack([U_pre0[None], o["U_after"][:-1]])
    sp1 = g.spec(dtype=np.float64, K=1)
    _, Xo = orc.rollout_states(sp1, np.zeros((sp.T, 2)), g.rec["x0"][i], np.clip(Upre_full, -1, 1)[None])
    assert np.max(np.abs(opt - Xo[0])) <= 5e-5
    with pytest.raises(ValueError):
        MPPIDynamicObstacles(noise_sigma=np.eye(2), num_samples=64, horizon=20, lambda_=1.0, u_min=[-1, -2], u_max=[1, 1])
