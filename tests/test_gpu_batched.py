"""Batched multi-robot mode (BASELINE config 4) vs R independent oracle ticks."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import Golden  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402


@pytest.mark.parametrize("model", ["diffdrive", "bicycle"])
def test_batched_robots_match_independent_oracle_ticks(model):
    from mppi_b200.batched import BatchedMPPI
    R, K = 12, 1024
    rng = np.random.default_rng(11)
    if model == "diffdrive":
        T = 30
        path = Golden("diffdrive_pe0.05").path
        sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
        sp.temperature = 2.0
        b = BatchedMPPI(R, path, num_samples_K=K, num_horizons_T=T, temperature=2.0, seed=5)
        x0 = np.stack([np.append(path[(7 * r) % 140, :2] + rng.normal(0, 0.1, 2), path[(7 * r) % 140, 2] + rng.normal(0, 0.1))
                       for r in range(R)])
    else:
        T = 20
        path = Golden("racecar_noobs").path
        sp = orc.racecar_spec(K=K, T=T, obstacles=None, dtype=np.float64)
        b = BatchedMPPI(R, path, model="bicycle", delta_t=0.05, max_u=(0.523, 2.0), num_samples_K=K, num_horizons_T=T,
                        param_exploration=0.01, param_lambda=50.0, param_alpha=1.0, sigma=((0.5, 0.0), (0.0, 0.1)),
                        stage_cost_weight=(50.0, 50.0, 1.0, 20.0), terminal_cost_weight=(50.0, 50.0, 1.0, 20.0),
                        window=200, seed=5)
        x0 = np.stack([path[(5 * r) % 90] + rng.normal(0, [0.3, 0.3, 0.05, 0.5]) for r in range(R)])
    x0_d = torch.from_numpy(x0.astype(np.float32)).cuda().contiguous()
    eps = torch.zeros(K, T, 2, dtype=torch.float32, device="cuda")
    U = np.zeros((R, T, 2))
    idx = np.zeros(R, dtype=int)
    for tick in range(2):
        u0 = b.step(x0_d).cpu().numpy()
        Unew = b.nominal()
        inew = b.waypoint_idx()
        for r in range(R):
            b.engine.generate_noise(eps, seed=5, tick=tick, robot=r)
            o = co.tick(sp, path, U[r], int(idx[r]), x0_d[r].cpu().numpy().astype(np.float64), eps.cpu().numpy())
            assert np.max(np.abs(Unew[r] - o["U_after"])) <= 2e-5, (model, tick, r)
            assert np.max(np.abs(u0[r] - o["u0"])) <= 2e-5
            assert inew[r] == o["idx_after"]
        U, idx = Unew.astype(np.float64), inew
    b.engine.close()


def test_batched_4096_robots_run_in_one_launch():
    """Full BASELINE config 4 shape: 4096 robots x K=1024 x H=30; sanity (finite, bounded) + launch count."""
    from mppi_b200.batched import BatchedMPPI
    path = Golden("diffdrive_pe0.05").path
    R = 4096
    b = BatchedMPPI(R, path, num_samples_K=1024, num_horizons_T=30, temperature=2.0, seed=1)
    rng = np.random.default_rng(0)
    x0 = np.stack([np.append(path[r % 168, :2] + rng.normal(0, 0.1, 2), path[r % 168, 2]) for r in range(R)])
    x0_d = torch.from_numpy(x0.astype(np.float32)).cuda().contiguous()
    l0 = b.engine.timings()["launches"]
    u0 = b.step(x0_d)
    b.engine.synchronize()
    assert b.engine.timings()["launches"] - l0 == 1
    u0 = u0.cpu().numpy()
    assert np.all(np.isfinite(u0)) and np.all(np.abs(u0[:, 0]) <= 5.0 + 1e-6) and np.all(np.abs(u0[:, 1]) <= 3.14 + 1e-6)
    idx = b.waypoint_idx()
    assert np.all(idx >= 0) and np.all(idx < 168)
    # robots that share the same state but not the same stream give different controls; same robot is reproducible
    b2 = BatchedMPPI(R, path, num_samples_K=1024, num_horizons_T=30, temperature=2.0, seed=1)
    u0b = b2.step(x0_d).cpu().numpy()
    assert np.array_equal(u0, u0b)
    b.engine.close(); b2.engine.close()


def _angle_diff(a, b):
    d = a - b
    return np.abs(np.arctan2(np.sin(d), np.cos(d)))


def test_device_spline_matches_reference_courses():
    """per-robot courses generated on the device vs the reference's own calc_spline_course output
    (tests/golden/spline_courses.npz) and its restatement; FP64 on the device, rounded once to float32."""
    import json
    import os
    from golden_util import GOLDEN_DIR
    from mppi_b200.batched import BatchedMPPI
    from oracle.spline_oracle import spline_course
    z = np.load(os.path.join(GOLDEN_DIR, "spline_courses.npz"))
    for i, case in enumerate(json.loads(str(z["meta"]))["cases"]):
        wx, wy = z["wx%d" % i].astype(np.float32), z["wy%d" % i].astype(np.float32)
        WX = np.stack([wx, wx + 3.0, wx[::-1].copy()])
        WY = np.stack([wy, wy - 1.5, wy[::-1].copy()])
        b = BatchedMPPI(3, None, num_samples_K=256, num_horizons_T=12, temperature=2.0)
        b.set_waypoints(WX, WY, ds=case["ds"], max_points=2048)
        for r in range(3):
            ref = spline_course(WX[r].astype(np.float64), WY[r].astype(np.float64), case["ds"])
            got = b.ref_path(r)
            assert got.shape == ref.shape, (i, r, got.shape, ref.shape)
            assert np.max(np.abs(got[:, :2] - ref[:, :2])) <= 2e-6 * max(1.0, np.max(np.abs(ref[:, :2])))
            assert np.max(_angle_diff(got[:, 2], ref[:, 2])) <= 2e-6
        # robot 0's waypoints are the golden's (rounded to float32): the reference's own course within that rounding
        gold = z["course%d" % i]
        got = b.ref_path(0)
        assert got.shape == gold.shape and np.max(np.abs(got[:, :2] - gold[:, :2])) <= 1e-5
        b.engine.close()
    # capacity and argument errors are reported, not truncated
    from mppi_b200 import MppiError
    b = BatchedMPPI(2, None, num_samples_K=256, num_horizons_T=12, temperature=2.0)
    with pytest.raises(MppiError):
        b.set_waypoints(np.array([[0, 5, 10.0], [0, 5, 10.0]]), np.zeros((2, 3)), ds=0.1, max_points=50)
    with pytest.raises(MppiError):
        b.step(torch.zeros(2, 3, device="cuda"))          # no path installed yet
    b.engine.close()


def test_batched_per_robot_paths_match_independent_oracle_ticks():
    """Every robot tracks ITS OWN spline course: one launch vs R independent oracle ticks on those courses."""
    from mppi_b200.batched import BatchedMPPI
    R, K, T = 10, 1024, 30
    rng = np.random.default_rng(5)
    n_wp = 6
    ang = np.cumsum(rng.normal(0, 0.5, (R, n_wp)), axis=1)
    step = rng.uniform(1.0, 2.0, (R, n_wp))
    WX = (np.cumsum(step * np.cos(ang), axis=1) + rng.normal(0, 3.0, (R, 1))).astype(np.float32)
    WY = (np.cumsum(step * np.sin(ang), axis=1) + rng.normal(0, 3.0, (R, 1))).astype(np.float32)
    b = BatchedMPPI(R, None, num_samples_K=K, num_horizons_T=T, temperature=2.0, seed=9)
    b.set_waypoints(WX, WY, ds=0.1)
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    paths = [b.ref_path(r).astype(np.float64) for r in range(R)]
    assert len({p.shape[0] for p in paths}) > 1                        # ragged: the robots' courses differ in length
    x0 = np.stack([np.append(p[5 + r, :2] + rng.normal(0, 0.1, 2), p[5 + r, 2] + rng.normal(0, 0.1)) for r, p in enumerate(paths)])
    x0_d = torch.from_numpy(x0.astype(np.float32)).cuda().contiguous()
    eps = torch.zeros(K, T, 2, dtype=torch.float32, device="cuda")
    U = np.zeros((R, T, 2))
    idx = np.zeros(R, dtype=int)
    for tick in range(2):
        u0 = b.step(x0_d).cpu().numpy()
        Unew, inew = b.nominal(), b.waypoint_idx()
        for r in range(R):
            b.engine.generate_noise(eps, seed=9, tick=tick, robot=r)
            o = co.tick(sp, paths[r], U[r], int(idx[r]), x0_d[r].cpu().numpy().astype(np.float64), eps.cpu().numpy())
            assert np.max(np.abs(Unew[r] - o["U_after"])) <= 2e-5, (tick, r, np.max(np.abs(Unew[r] - o["U_after"])))
            assert inew[r] == o["idx_after"]
        U, idx = Unew.astype(np.float64), inew
    # a shared path can be installed again afterwards
    b.engine.set_ref_path(Golden("diffdrive_pe0.05").path)
    assert b.ref_path(3).shape[0] == 168
    b.engine.close()


# ---- SURVEY 8f row 1 remainder: fleets in closed loop on the device, the n-tick loop captured in a CUDA graph ---------------
def _fleet_oracle_loop(b, sp, path, x0, r, n, plant, tick_base=0, U=None, idx=0):
    """Robot r of the fleet stepped from the host: oracle tick fed the robot's exported Philox noise + the reference plant."""
    K, T = sp.K, sp.T
    eps = torch.zeros(K, T, 2, dtype=torch.float32, device="cuda")
    x = x0.astype(np.float32).astype(np.float64)
    U = np.zeros((T, 2)) if U is None else U
    xs, us = [x.copy()], []
    for i in range(n):
        b.engine.generate_noise(eps, seed=b.seed, tick=tick_base + i, robot=r)
        o = co.tick(sp, path, U, idx, x.astype(np.float32).astype(np.float64), eps.cpu().numpy())
        U, idx = o["U_after"], o["idx_after"]
        x = plant(x, o["u0"])
        xs.append(x.copy()); us.append(o["u0"].copy())
    return np.array(xs), np.array(us), U, idx


@pytest.mark.parametrize("model", ["diffdrive", "bicycle"])
def test_fleet_closed_loop_on_device_matches_independent_oracle_loops(model):
    from mppi_b200.batched import BatchedMPPI
    R, K, n = 6, 1024, 12
    rng = np.random.default_rng(3)
    if model == "diffdrive":
        T = 20
        path = Golden("diffdrive_pe0.05").path
        sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
        sp.temperature = 2.0
        b = BatchedMPPI(R, path, num_samples_K=K, num_horizons_T=T, temperature=2.0, seed=9)
        x0 = np.stack([np.append(path[(11 * r) % 120, :2] + rng.normal(0, 0.05, 2), path[(11 * r) % 120, 2]) for r in range(R)])
        plant = lambda x, u: orc.plant_diffdrive(x, u, sp.dt)          # noqa: E731
    else:
        T = 20
        path = Golden("racecar_noobs").path
        sp = orc.racecar_spec(K=K, T=T, obstacles=None, dtype=np.float64)
        b = BatchedMPPI(R, path, model="bicycle", delta_t=0.05, max_u=(0.523, 2.0), num_samples_K=K, num_horizons_T=T,
                        param_exploration=0.01, param_lambda=50.0, param_alpha=1.0, sigma=((0.5, 0.0), (0.0, 0.1)),
                        stage_cost_weight=(50.0, 50.0, 1.0, 20.0), terminal_cost_weight=(50.0, 50.0, 1.0, 20.0),
                        window=200, seed=9)
        x0 = np.stack([path[(9 * r) % 80] + rng.normal(0, [0.2, 0.2, 0.03, 0.3]) for r in range(R)])
        plant = lambda x, u: orc.plant_bicycle(x, u, sp.dt)            # noqa: E731
    states, controls = b.run_closed_loop(x0, n)
    assert states.shape == (n + 1, R, b.nx) and controls.shape == (n, R, 2)
    # a SECOND call continues every robot's loop (same instantiated graph, ticks n .. 2n-1)
    states2, controls2 = b.run_closed_loop(states[-1], n)
    launches = b.engine.timings()["launches"]
    assert launches == 2 * n, launches                                  # one launch per tick for the whole fleet
    idx_dev, U_dev = b.waypoint_idx(), b.nominal()
    for r in range(R):
        xs, us, U, idx = _fleet_oracle_loop(b, sp, path, x0[r], r, n, plant)
        assert np.max(np.abs(states[:, r] - xs)) <= 5e-4, (model, r, np.max(np.abs(states[:, r] - xs)))
        assert np.max(np.abs(controls[:, r] - us)) <= 2e-4, (model, r)
        xs2, us2, U, idx = _fleet_oracle_loop(b, sp, path, states[-1, r].astype(np.float64), r, n, plant, tick_base=n, U=U, idx=idx)
        assert np.max(np.abs(states2[:, r] - xs2)) <= 2e-3, (model, r, np.max(np.abs(states2[:, r] - xs2)))
        assert idx_dev[r] == idx
        assert np.max(np.abs(U_dev[r] - U)) <= 5e-3
    b.engine.close()


def test_fleet_of_4096_robots_50_ticks_closed_loop_subset_vs_oracle():
    """BASELINE config 4 in closed loop: 4096 robots x K=1024 x H=30, 50 ticks, no host round trip; a random subset of
    robots against independent oracle loops (the loops are compared over the first 15 ticks: afterwards FP32/FP64
    differences are amplified by the closed loop itself)."""
    from mppi_b200.batched import BatchedMPPI
    path = Golden("diffdrive_pe0.05").path
    R, K, T, n = 4096, 1024, 30, 50
    b = BatchedMPPI(R, path, num_samples_K=K, num_horizons_T=T, temperature=2.0, seed=4)
    rng = np.random.default_rng(0)
    x0 = np.stack([np.append(path[r % 100, :2] + rng.normal(0, 0.1, 2), path[r % 100, 2] + rng.normal(0, 0.1)) for r in range(R)])
    b.engine.set_waypoint_idx(np.arange(R) % 100)                       # every robot starts next to its own waypoint
    states, controls = b.run_closed_loop(x0, n)
    assert np.all(np.isfinite(states)) and np.all(np.isfinite(controls))
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    plant = lambda x, u: orc.plant_diffdrive(x, u, sp.dt)              # noqa: E731
    for r in rng.choice(R, 4, replace=False):
        xs, us, _, _ = _fleet_oracle_loop(b, sp, path, x0[r], int(r), 15, plant, idx=int(r % 100))
        assert np.max(np.abs(states[:16, r] - xs)) <= 1e-3, (r, np.max(np.abs(states[:16, r] - xs)))
        assert np.max(np.abs(controls[:15, r] - us)) <= 5e-4, r
    # the fleet really ran closed loop (not 50 copies of tick 0): the robots stay on the path they started ~0.08 m off and move
    def cross_track(xy):
        return np.sqrt(((xy[:, None, :] - path[None, :, :2]) ** 2).sum(-1)).min(axis=1).mean()
    sub = rng.choice(R, 256, replace=False)
    assert cross_track(states[-1, sub, :2]) < 1.5 * cross_track(states[0, sub, :2]) + 0.02
    assert np.median(np.linalg.norm(states[-1, :, :2] - states[0, :, :2], axis=1)) > 0.05
    assert np.median(np.abs(controls[1:] - controls[:-1]).max(axis=(0, 2))) > 1e-3          # controls change from tick to tick
    b.engine.close()
