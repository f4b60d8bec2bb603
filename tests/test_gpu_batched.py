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
