"""Race evidence without a sanitizer (compute-sanitizer is closed on this pool): a data race between warps, CTAs or
clusters shows up as run-to-run differences, so every hand-off the tick relies on is exercised repeatedly and must give
BIT-IDENTICAL results -- the stash reused as merge scratch by the last CTA, the atomic-ticket election, the float4 partial
merge, the hand-off of split quads between clusters in the learned-dynamics kernel, the grid-wide ticket of the
graph-captured closed loop -- and no kernel may write past any device buffer (guard zones, mppi_debug_check_guards)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import Golden  # noqa: E402
from gpu_util import engine_from_spec  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402


@pytest.mark.parametrize("K,T,mode", [(1 << 20, 50, "philox"), (300000, 50, "philox"), (20000, 30, "injected"), (5000, 61, "philox")])
def test_tick_is_bit_reproducible_and_stays_inside_its_buffers(K, T, mode):
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    eps = None
    if mode == "injected":
        eps = torch.zeros(K, T, 2, device="cuda")
        eng.generate_noise(eps, seed=5, tick=3)
    U0 = np.random.default_rng(1).normal(0, 0.3, (T, 2)).astype(np.float32)
    x0 = np.array([0.3, 0.2, 0.4])
    ref = None
    for rep in range(25 if K <= 300000 else 12):
        eng.set_nominal(U0)
        eng.set_waypoint_idx(0)
        u0, u = eng.step(x0, eps, seed=5, tick=3)
        if ref is None:
            ref = u.copy()
        assert np.array_equal(u, ref), (K, T, mode, rep, np.max(np.abs(u - ref)))
    assert eng.check_guards() == 0
    eng.close()


def test_learned_dynamics_schedules_are_bit_reproducible():
    """Balanced ping-pong schedule (state records handed between clusters through global flags): 10 launches, same bits."""
    g = Golden("diffdrive_pe0.05")
    for K, T, n_in, n_hidden in ((50000, 12, 3, 2), (50000, 11, 5, 2), (30000, 11, 5, 3)):       # pair MMAs, tcgen05 layer 1, two GEMMs
        mlp = orc.make_mlp(seed=2, out_scale=0.05, n_in=n_in, scalers=(n_in == 5), scaler_gain=1.0, n_hidden=n_hidden)
        sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen", model="diffdrive_mlp", mlp=mlp)
        sp.temperature = 2.0
        eng = engine_from_spec(sp, g.path)
        sc = [mlp[k] for k in ("in_mean", "in_scale", "out_mean", "out_scale")] if n_in == 5 else []
        eng.set_mlp([mlp["W%d" % i] for i in range(n_hidden + 2)], [mlp["b%d" % i] for i in range(n_hidden + 2)], *sc)
        S = torch.zeros(K, device="cuda")
        ref = None
        for rep in range(10):
            S.zero_()
            eng.set_waypoint_idx(0)
            eng.rollout_costs(np.array([0.4, 0.3, 0.5]), S, None, seed=9, tick=2)
            if ref is None:
                ref = S.clone()
            assert torch.equal(S, ref), (K, rep)
        assert eng.check_guards() == 0
        eng.close()


def test_fleet_closed_loop_graph_is_bit_reproducible():
    from mppi_b200.batched import BatchedMPPI
    path = Golden("diffdrive_pe0.05").path
    R = 300
    x0 = np.stack([np.append(path[r % 100, :2], path[r % 100, 2]) for r in range(R)])
    runs = []
    for rep in range(3):
        b = BatchedMPPI(R, path, num_samples_K=512, num_horizons_T=20, temperature=2.0, seed=4)
        st, ct = b.run_closed_loop(x0, 12)
        runs.append((st.copy(), ct.copy()))
        assert b.engine.check_guards() == 0
        b.engine.close()
    for st, ct in runs[1:]:
        assert np.array_equal(st, runs[0][0]) and np.array_equal(ct, runs[0][1])


def test_non_finite_state_fails_the_tick_loudly_and_leaves_the_nominal_alone():
    """A NaN observed state must not poison the nominal for good: MPPI_E_NUMERIC, nothing applied, next tick fine."""
    from mppi_b200 import MppiError
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=4096, T=30, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    U0 = np.random.default_rng(1).normal(0, 0.3, (30, 2)).astype(np.float32)
    eng.set_nominal(U0)
    with pytest.raises(MppiError, match="non-finite"):
        eng.step(np.array([np.nan, 0.0, 0.0]), None, seed=1, tick=0)
    assert np.array_equal(eng.get_nominal(), U0) and eng.get_waypoint_idx() == 0
    u0, u = eng.step(np.array([0.1, 0.0, 0.0]), None, seed=1, tick=1)
    assert np.all(np.isfinite(u)) and not np.array_equal(u, U0)
    eng.close()


def test_waypoint_index_outside_the_path_is_refused():
    """an index the reference would fail on (empty window slice) is refused instead of read out of bounds, and a
    shorter path re-assigned later clamps the carried index."""
    from mppi_b200 import MppiError
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=512, T=30, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    eng = engine_from_spec(sp, g.path)
    for bad in (-1, len(g.path), 10 ** 6):
        with pytest.raises(MppiError):
            eng.set_waypoint_idx(bad)
    eng.set_waypoint_idx(len(g.path) - 1)
    eng.set_ref_path(g.path[:50])
    assert eng.get_waypoint_idx() == 49
    u0, u = eng.step(np.array([0.0, 0.0, 0.0]), None, seed=1, tick=0)
    assert np.all(np.isfinite(u)) and eng.check_guards() == 0
    eng.close()
