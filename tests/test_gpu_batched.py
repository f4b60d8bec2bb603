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
