"""BASELINE config 1 at its literal size: the reference's own diff-drive class,
K = 1000 samples, T = 30, seeds 0-4, param_exploration in {1e-4 (the literal default), 0.05}, 200-tick
closed loop through DifferentialDrive.update_state on the 168-point spline course
(controllers/mppi_differential_drive.py:33-40,87-165,400-410).  tests/golden/c1_K1000_T30_*.npz hold
what the UNMODIFIED class produced (make_golden.py c1); the noise is regenerated from the seed.

CPU part: the oracle restatements reproduce the class at this size (more samples = more waypoint-index
breakpoints per tick than the K = 256 fixtures exercise).  GPU part: the strict multi-pass driver
behind the drop-in class, literal modes (cost_mode='last', waypoint_mode='strict')."""
import numpy as np
import pytest

from golden_util import C1_PES, C1_SEEDS, C1Golden
from oracle import c_oracle as co
from oracle import mppi_oracle as orc

COST_RTOL = 1e-5       # FP32 device arithmetic vs the FP64 reference class
U_ATOL = 2e-5


def _nominal_before(g, i):
    return np.zeros((g.meta["num_horizons_T"], 2)) if i == 0 else g.z["U_after"][i - 1]


@pytest.mark.parametrize("pe", C1_PES)
@pytest.mark.parametrize("seed", [0, 3])
def test_oracles_reproduce_the_reference_class_at_K1000(seed, pe):
    g = C1Golden(seed, pe)
    sp = g.spec()
    es = g.eps_stream()
    ticks = list(g.z["S_ticks"])
    for i in range(12):
        eps = next(es).astype(np.float64)
        if i not in ticks:
            continue
        j = ticks.index(i)
        inp = dict(path=g.path, U=_nominal_before(g, i), idx=int(g.z["idx0"][i]), x0=g.z["x0"][i], eps=eps)
        for name, o in (("vec", orc.tick_vec(sp, **inp)), ("c", co.tick(sp, **inp))):
            assert np.max(np.abs(o["S"] - g.z["S"][j]) / np.abs(g.z["S"][j])) <= 1e-12, (name, i)
            assert o["idx_after"] == int(g.z["idx_after"][i]), (name, i)
            assert np.max(np.abs(o["U_after"] - g.z["U_after"][i])) <= 1e-10, (name, i)


def test_c1_noise_stream_is_the_documented_one():
    """The regenerated noise is N(0, diag(0.1, 0.01)) rounded to float32 and reproduces the stored weighted-noise sum of
    the reference tick (an independent check that the stream matches the one the fixture was made with)."""
    g = C1Golden(2, "0.05")
    eps = next(g.eps_stream())
    assert eps.dtype == np.float32 and eps.shape == (1000, 30, 2)
    sp = g.spec()
    o = orc.tick_vec(sp, g.path, _nominal_before(g, 0), int(g.z["idx0"][0]), g.z["x0"][0], eps.astype(np.float64))
    assert np.max(np.abs(o["w_eps"] - g.z["w_eps"][0])) <= 1e-12


# ---------------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _strict_waypoint_margin(sp, path, U, idx0, idx_after, x0, eps, rows):
    """Smallest relative gap between the best and the second-best waypoint distance over the rollout of the samples
    `rows`, over every window start the strict rule can have used during this tick (the index only moves forward, from
    the carried index to the index after the tick)."""
    path = np.asarray(path, np.float64)
    _, X = orc.rollout_states(sp, np.asarray(U, np.float64), np.asarray(x0, np.float64), eps)
    X = X[rows]
    best = np.full(len(rows), np.inf)
    for s0 in range(max(0, idx0), idx_after + 1):
        seg = path[s0:s0 + sp.window]
        if seg.shape[0] < 2:
            continue
        d = np.sort((X[..., 0:1] - seg[:, 0]) ** 2 + (X[..., 1:2] - seg[:, 1]) ** 2, axis=-1)
        best = np.minimum(best, ((d[..., 1] - d[..., 0]) / np.maximum(d[..., 1], 1e-30)).min(axis=1))
    return best


def _ctrl(g, **kw):
    from mppi_b200.mppi_differential_drive import MPPIAlgorithms
    m = g.meta
    return MPPIAlgorithms(delta_t=m["delta_t"], ref_path=g.path, max_speed=m["max_speed"], max_omega=m["max_omega"],
                          num_samples_K=m["num_samples_K"], num_horizons_T=m["num_horizons_T"],
                          param_exploration=m["param_exploration"], param_lambda=m["param_lambda"], param_alpha=m["param_alpha"],
                          sigma=np.array([[0.1, 0.0], [0.0, 0.01]]), stage_cost_weight=np.array([5.0, 5.0, 10.0]),
                          terminal_cost_weight=np.array([5.0, 5.0, 10.0]), visualize_optimal_traj=False,
                          visualze_sampled_trajs=False, **kw)


@gpu
@pytest.mark.parametrize("pe", C1_PES)
@pytest.mark.parametrize("seed", C1_SEEDS)
def test_gpu_strict_costs_and_index_on_every_stored_tick(seed, pe):
    """Teacher-forced ticks (state, nominal and index of the reference run): per-sample costs rtol 1e-5, the index after
    the tick equal, and -- where the soft-min is not decided by a near-tie -- the shifted nominal within 2e-5."""
    torch = pytest.importorskip("torch")
    g = C1Golden(seed, pe)
    ctrl = _ctrl(g)                                  # literal defaults: cost_mode='last', waypoint_mode='strict'
    eng = ctrl.engine
    sp = g.spec()
    S = torch.zeros(1000, dtype=torch.float32, device="cuda")
    es = g.eps_stream()
    ticks = list(g.z["S_ticks"])
    passes = []
    n_u_checked = n_near_tie = 0
    for i in range(g.n_ticks):
        eps = next(es)
        if i not in ticks:
            continue
        j = ticks.index(i)
        d_eps = torch.from_numpy(eps).cuda()
        eng.set_nominal(_nominal_before(g, i))
        eng.set_waypoint_idx(int(g.z["idx0"][i]))
        eng.rollout_costs(g.z["x0"][i], S, d_eps)
        Sg, Sr = S.cpu().numpy().astype(np.float64), g.z["S"][j]
        rel = np.abs(Sg - Sr) / np.abs(Sr)
        bad = np.nonzero(rel > COST_RTOL)[0]
        # a nearest-waypoint near-tie decided differently in FP32 and FP64 changes one sample's cost by a visible amount:
        # at most 2 samples of the 1000, and each one must BE a near-tie for the FP64 restatement
        assert bad.size <= 2, (seed, pe, i, bad.size, rel.max())
        if bad.size:
            n_near_tie += bad.size
            m = _strict_waypoint_margin(sp, g.path, _nominal_before(g, i), int(g.z["idx0"][i]), int(g.z["idx_after"][i]),
                                        g.z["x0"][i], eps.astype(np.float64), bad)
            assert np.all(m < 1e-4), (seed, pe, i, bad, m)
        assert eng.get_waypoint_idx() == int(g.z["idx_after"][i]), (seed, pe, i)
        passes.append(eng.timings()["last_passes"])
        # full tick through the class
        eng.set_nominal(_nominal_before(g, i))
        ctrl.prev_way_point_idx = int(g.z["idx0"][i])
        u0, u, _, _ = ctrl._calc_input_control(g.z["x0"][i], noise=eps)
        assert ctrl.prev_way_point_idx == int(g.z["idx_after"][i])
        if bad.size:
            continue                                  # the update below is compared on ticks without a flipped near-tie
        # K2 given the device's own costs (within 1e-5 of the reference's, checked above): the update must be the exact
        # soft-min of THOSE costs -- at temperature 1e-4 the weights amplify a 1e-6 relative cost difference 10^4-fold, so
        # this is the well-posed form of the end-to-end check ('softmax conditioning') ...
        o = orc.update_vec(sp, _nominal_before(g, i), Sg, eps.astype(np.float64), g.z["idx_after"][i])
        assert np.max(np.abs(u - o["U_after"])) <= U_ATOL, (seed, pe, i, np.max(np.abs(u - o["U_after"])))
        # ... and end to end against the reference class wherever its own soft-min is well conditioned: always at
        # temperature 0.05; at 1e-4 when no runner-up within FP32 cost resolution of the minimum carries weight
        tau = g.meta["param_exploration"]
        gap = np.partition(Sr, 1)[1] - Sr.min()
        if tau >= 0.05 or gap > 20.0 * tau:
            assert np.max(np.abs(u - g.z["U_after"][i])) <= U_ATOL, (seed, pe, i, np.max(np.abs(u - g.z["U_after"][i])))
            n_u_checked += 1
        assert np.array_equal(u0, u[0])                                            # Q8
    assert max(passes) >= 2, "no tick exercised an index breakpoint"          # the multi-pass rule really ran
    assert n_u_checked >= (len(ticks) - n_near_tie if float(pe) >= 0.05 else 1), n_u_checked
    assert n_near_tie <= 3, n_near_tie                # ~1e-4 of the 29 000 sample evaluations of a case


@gpu
@pytest.mark.parametrize("seed", C1_SEEDS)
def test_gpu_closed_loop_200_ticks_within_1cm(seed):
    """pe = 0.05: the drop-in class driven by its OWN outputs through the unicycle plant stays within 1 cm of the
    trajectory the reference class produced with the same noise, over all 200 ticks, and ends on the same waypoint."""
    g = C1Golden(seed, "0.05")
    ctrl = _ctrl(g)
    es = g.eps_stream()
    x = g.z["x0"][0].astype(np.float64).copy()
    dev = 0.0
    for i in range(g.n_ticks):
        dev = max(dev, float(np.max(np.abs(x[:2] - g.z["x0"][i][:2]))))
        u0, _, _, _ = ctrl._calc_input_control(x, noise=next(es))
        x = orc.plant_diffdrive(x, np.asarray(u0, dtype=np.float64), 0.1)
    assert dev < 1e-2, (seed, dev)
    assert ctrl.prev_way_point_idx == int(g.z["idx_after"][-1])
