"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE CLASSES from
/root/reference (build container only).  Usage:  python tests/golden/make_golden.py

Each file holds the inputs of one or more control ticks (path, x0, nominal U, carried
waypoint index, injected noise) and what the reference class produced (S, w, weighted
noise before/after the filter, the shifted nominal, the returned u0, the index after).
Noise is drawn once, rounded to float32 and stored, so the reference, the oracle and
the CUDA path all consume bit-identical values.  numpy's version is recorded (Q12).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_loader  # noqa: E402
sys.path.insert(0, os.path.join(HERE, ".."))
from golden_util import c1_eps  # noqa: E402  (the tests regenerate the C1 noise with the same function)


def spline_path(ref):
    cx, cy, cyaw, _, _ = ref["calc_spline_course"]([0.0, 0.5, 1.0, 3.0, 3.0, 1.0, -4.0],
                                                   [0.0, 1.0, 1.0, 2.0, 5.0, 1.0, -1.0], ds=0.1)
    return np.array([cx, cy, cyaw]).T


def draw_eps(rng, sigma, K, T):
    return rng.multivariate_normal(np.zeros(2), sigma, (K, T)).astype(np.float32)


def run_ticks(ctrl, idx_attr, states, eps_list, plant=None, keep_sampled=False):
    """Steps `ctrl` through len(eps_list) ticks.  `states` is a list of observed states, or
    a single initial state when `plant` (closed loop) is given."""
    cap = ref_loader.instrument(ctrl, [e.astype(ctrl.u_prev.dtype) for e in eps_list])
    rec = {k: [] for k in ("x0", "U0", "idx0", "S", "w", "w_eps", "w_eps_filt", "U_after", "u0", "idx_after")}
    x = np.array(states[0], dtype=float)
    for i in range(len(eps_list)):
        if plant is None:
            x = np.array(states[i], dtype=float)
        rec["x0"].append(x.copy())
        rec["U0"].append(np.array(ctrl.u_prev, copy=True))
        rec["idx0"].append(int(getattr(ctrl, idx_attr)))
        with ref_loader.quiet():
            step = ctrl._calc_input_control if hasattr(ctrl, "_calc_input_control") else ctrl._calc_control_input
            u0, useq, out_traj, out_samp = step(x)
        for k in ("S", "w", "w_eps", "w_eps_filt"):
            rec[k].append(cap[k])
        rec["U_after"].append(np.array(ctrl.u_prev, copy=True))
        rec["u0"].append(np.array(u0, copy=True))
        rec.setdefault("optimal_traj", []).append(np.array(out_traj, copy=True))
        if keep_sampled:
            rec.setdefault("sampled_traj", []).append(np.array(out_samp, copy=True))
        rec["idx_after"].append(int(getattr(ctrl, idx_attr)))
        if plant is not None:
            x = plant(x, np.array(u0, dtype=float))
    return {k: np.array(v) for k, v in rec.items()}



def save(name, meta, path, eps_list, rec, obstacles=None):
    rec = {k: np.array(v) if isinstance(v, list) else v for k, v in rec.items()}
    meta = dict(meta, numpy=np.__version__, generator="tests/golden/make_golden.py",
                source="unmodified reference classes executed in the build container")
    out = dict(meta=json.dumps(meta), path=path, eps=np.array(eps_list, dtype=np.float32), **rec)
    if obstacles is not None:
        out["obstacles"] = np.asarray(obstacles, float)
    fn = os.path.join(HERE, name + ".npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, "%.1f KB" % (os.path.getsize(fn) / 1024))


def make_extras():
    """the goal-point class and the moving-soft-obstacle running cost."""
    import types
    import torch
    ref = ref_loader.load_reference_extras()
    # ---- goal-point diff-drive MPPI (test/mppi_differential_drive_obs.py), the script's own parameters (:396-424)
    K, T = 160, 20
    sigma = np.array([[0.1, 0.0], [0.0, 0.01]])
    w = 10 * np.array([5.0, 9.0])
    obs = np.array([[5.0, 3.0, 0.5], [3.0, 2.5, 0.5]])
    kw = dict(delta_t=0.1, max_speed=10.0, max_omega=5.0, num_samples_K=K, num_horizons_T=T, param_exploration=0.1,
              param_lambda=1.0, param_alpha=0.98, safety_margin_rate=0.8)
    goal = np.array([5.0, 5.0])
    ctrl = ref["GoalMPPI"](goal_point=goal, sigma=sigma, stage_cost_weight=w, terminal_cost_weight=w,
                           obstacle_circles=obs, visualize_optimal_traj=False, visualze_sampled_trajs=False, **kw)
    ctrl.prev_way_point_idx = 0                       # the class carries no waypoint index; run_ticks records one
    rng = np.random.default_rng(11)
    eps_list = [draw_eps(rng, sigma, K, T) for _ in range(6)]
    # ticks 0-2 closed loop from the origin, then states next to the obstacles / the goal (bearing wrap, collisions)
    states = [[0, 0, 0], [0.4, 0.05, 0.2], [0.9, 0.3, 0.5], [2.4, 1.8, 0.7], [4.6, 2.2, 2.9], [5.2, 4.6, -2.8]]
    rec = run_ticks(ctrl, "prev_way_point_idx", states, eps_list)
    save("diffdrive_goal", dict(kind="diffdrive_goal", goal=goal.tolist(), **kw), np.zeros((1, 3)), eps_list, rec,
         obstacles=obs)

    # ---- moving soft obstacles (test/test_mppi_diff_obs.py): the script's dynamics + running_cost, executed by its
    # own MPPIWrapper._compute_rollout_costs loop on clamped controls V = clip(U + eps)
    K, T, dt = 192, 25, 0.05
    sig = np.diag([0.5, 0.5])
    rng = np.random.default_rng(12)
    recs = dict(x0=[], U0=[], V=[], S=[])
    eps_list = []
    u_lim = np.array([1.0, 1.0], dtype=np.float32)
    for x0 in ([3.0, 3.0, 0.3], [4.2, 3.1, 0.9], [4.9, 5.2, 1.4], [0.0, 0.0, 0.0]):
        eps = draw_eps(rng, sig, K, T)
        U = (rng.normal(0, 0.3, (T, 2))).astype(np.float32)
        V = np.clip(U[None] + eps, -u_lim, u_lim).astype(np.float32)
        # pytorch_mppi wraps the user callbacks so they see flat (N, nx) / (N, nu) batches; same glue here
        def dyn(st, u, t):
            return ref["dynobs_dynamics"](st.reshape(-1, 3), u.reshape(-1, 2), t).reshape(st.shape)

        def cost(st, u, t):
            return ref["dynobs_running_cost"](st.reshape(-1, 3), u.reshape(-1, 2), t).reshape(st.shape[:-1])
        fake = types.SimpleNamespace(nu=2, nx=3, d="cpu", dtype=torch.float32, M=1, u_scale=1.0, delta_t=dt,
                                     state=torch.tensor(x0, dtype=torch.float32), _dynamics=dyn,
                                     _running_cost=cost, rollout_var_discount=1.0,
                                     rollout_var_cost=0.0, terminal_state_cost=None)
        cost_total, _, _ = ref["dynobs_wrapper"]._compute_rollout_costs(fake, torch.from_numpy(V))
        eps_list.append(eps)
        recs["x0"].append(np.array(x0)); recs["U0"].append(U); recs["V"].append(V); recs["S"].append(cost_total.numpy())
    meta = dict(kind="diffdrive_target_soft", K=K, T=T, delta_t=dt, u_max=[1.0, 1.0], target=[6.0, 6.0, 1.57],
                Q=[30.0, 5.0, 9.0], R=[0.1, 0.1], soft_w=100.0, soft_sd=2.0, sigma=sig.tolist(),
                obs_pos=ref["dynobs_positions"].tolist(), obs_vel=ref["dynobs_velocities"].tolist(),
                note="S = MPPIWrapper._compute_rollout_costs (test/test_mppi_diff_obs.py:113-151) on V; float32 torch")
    save("diffdrive_target_soft", meta, np.zeros((1, 3)), eps_list, recs)


def make_mlp_golden():
    """forward passes of the reference's own torch modules (3-input dnn/simple_mlp.py and the
    5-input class the trained saved_models belong to), loaded with the seeded weights oracle.make_mlp regenerates on
    any machine, with StandardScaler objects applied the way test/bullet_differential_drive_dnn.py:363-370 does."""
    import torch
    from sklearn.preprocessing import StandardScaler
    from oracle.mppi_oracle import make_mlp
    ref = ref_loader.load_reference_mlps()
    rng = np.random.default_rng(21)
    out = {}
    for tag, n_in, cls, last, n_hidden in (("3", 3, ref["MLP3"], "output_layer", 2), ("5", 5, ref["MLP5"], "out_layer", 2),
                                           ("5_3l", 5, lambda: ref["MLP5_3L"](5), "out_layer", 3)):
        w = make_mlp(seed=5, out_scale=0.01, dtype=np.float32, n_in=n_in, scalers=(n_in == 5), n_hidden=n_hidden)
        net = cls()
        names = ["input_layer"] + ["hidden_layer.%d" % i for i in range(n_hidden)] + [last]
        net.load_state_dict({n + s: torch.from_numpy(w[k + str(i)]) for i, n in enumerate(names)
                             for s, k in ((".weight", "W"), (".bias", "b"))})
        net = net.double()
        X = rng.normal(0, 1.5, (48, n_in))
        if n_in == 5:
            ss, cs, es = StandardScaler(), StandardScaler(), StandardScaler()
            ss.mean_, ss.scale_ = w["in_mean"][:3], w["in_scale"][:3]
            cs.mean_, cs.scale_ = w["in_mean"][3:], w["in_scale"][3:]
            es.mean_, es.scale_ = w["out_mean"], w["out_scale"]
            xin = np.concatenate([ss.transform(X[:, :3]), cs.transform(X[:, 3:])], axis=1)
            with torch.no_grad():
                y = es.inverse_transform(net(torch.from_numpy(xin)).numpy())
        else:
            with torch.no_grad():
                y = net(torch.from_numpy(X)).numpy()
        out["X" + tag], out["Y" + tag] = X, y
    out["meta"] = json.dumps(dict(seed=5, out_scale=0.01, numpy=np.__version__, torch=torch.__version__,
                                  source="dnn/simple_mlp.py:MultiLayerPerception and "
                                         "simulation/bullet_differential_drive_dnn.py:MultiLayerPerceptron, "
                                         "train/train_diff_mlp.py:MultiLayerPerceptron(5) (three hidden layers), float64"))
    fn = os.path.join(HERE, "mlp_forward.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, "%.1f KB" % (os.path.getsize(fn) / 1024))


# ---- BASELINE config 1 at its literal size: K=1000, T=30, seeds 0-4, param_exploration in {1e-4, 0.05} ----
C1_K, C1_T, C1_TICKS = 1000, 30, 200
C1_S_TICKS = list(range(0, 10)) + list(range(10, 200, 10))      # ticks whose per-sample costs are stored (29 of 200)


def _c1_case(args):
    seed, pe = args
    ref = ref_loader.load_reference()
    path = spline_path(ref)
    sigma_dd = np.array([[0.1, 0.0], [0.0, 0.01]])
    w_dd = np.array([5.0, 5.0, 10.0])
    kw = dict(delta_t=0.1, max_speed=5.0, max_omega=3.14, num_samples_K=C1_K, num_horizons_T=C1_T,
              param_exploration=pe, param_lambda=1.0, param_alpha=0.2)
    ctrl = ref["MPPIAlgorithms"](ref_path=path, sigma=sigma_dd, stage_cost_weight=w_dd, terminal_cost_weight=w_dd,
                                 visualize_optimal_traj=False, visualze_sampled_trajs=False, **kw)
    rng = np.random.default_rng(seed)
    eps_list = [c1_eps(rng, C1_K, C1_T) for _ in range(C1_TICKS)]

    def plant(x, u):
        return ref["DifferentialDrive"](x).update_state(0.1, x, u)      # controllers/mppi_differential_drive.py:33-40
    rec = run_ticks(ctrl, "prev_way_point_idx", [[0, 0, 0]], eps_list, plant=plant)
    out = dict(x0=rec["x0"], idx0=rec["idx0"].astype(np.int32), idx_after=rec["idx_after"].astype(np.int32),
               u0=rec["u0"], U_after=rec["U_after"],            # the nominal before tick i is U_after[i-1] (zeros for i = 0)
               S=rec["S"][C1_S_TICKS].astype(np.float64), S_ticks=np.array(C1_S_TICKS, dtype=np.int32),
               w_eps=rec["w_eps"][C1_S_TICKS])
    meta = dict(kind="diffdrive", seed=seed, numpy=np.__version__, generator="tests/golden/make_golden.py c1",
                noise="tests/golden_util.py:c1_eps(np.random.default_rng(seed)), one call per tick",
                source="unmodified controllers/mppi_differential_drive.py:MPPIAlgorithms + DifferentialDrive.update_state, "
                       "200-tick closed loop from (0,0,0) on the spline course", **kw)
    tag = "1e-4" if pe == 1e-4 else "%g" % pe
    fn = os.path.join(HERE, "c1_K1000_T30_seed%d_pe%s.npz" % (seed, tag))
    np.savez_compressed(fn, meta=json.dumps(meta), path=path, **out)
    return fn, os.path.getsize(fn) / 1024


def make_c1_literal(seeds=(0, 1, 2, 3, 4), pes=(1e-4, 0.05)):
    """~1 s per reference tick: 10 cases x 200 ticks, one process per case."""
    import multiprocessing as mp
    cases = [(s, pe) for s in seeds for pe in pes]
    with mp.get_context("fork").Pool(min(len(cases), os.cpu_count() or 1)) as pool:
        for fn, kb in pool.imap_unordered(_c1_case, cases):
            print("wrote", fn, "%.1f KB" % kb, flush=True)


def make_trained_mlp_golden():
    """SURVEY.md 8f row 4 / VERDICT r1 item 1b: the reference's TRAINED residual checkpoints and their StandardScaler
    statistics as plain arrays (category (b) test fixtures -- data the reference ships, not code), each with forward passes
    of the reference's own torch class + sklearn scalers (test/test_diff_dyna_eval.py:54-56 composition) on 64 inputs, so
    the GPU box can pin the oracle to the reference without the reference tree.
      mlp_diff_300x100.pth          + scalers_mlp_diff_300x100_20_l.pth      two hidden layers
                                      (simulation/bullet_differential_drive_dnn.py:37-60)
      mlp_diff_300x100_3l_mppi.pth  + scalers_mlp_diff_300x100_3l_mppi.pth   three hidden layers
                                      (train/train_diff_mlp.py:13-36)"""
    import warnings
    import torch
    ref = ref_loader.load_reference_mlps()
    root = ref_loader.REFERENCE_ROOT
    for tag, ckpt, scalers, cls_key in (("mlp_diff_300x100", "mlp_diff_300x100.pth", "scalers_mlp_diff_300x100_20_l.pth", "MLP5"),
                                        ("mlp_diff_300x100_3l_mppi", "mlp_diff_300x100_3l_mppi.pth",
                                         "scalers_mlp_diff_300x100_3l_mppi.pth", "MLP5_3L")):
        sd = torch.load(os.path.join(root, "saved_models", ckpt), map_location="cpu")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sc = torch.load(os.path.join(root, "saved_models", scalers), weights_only=False)
        n_hidden = 3 if cls_key == "MLP5_3L" else 2
        names = ["input_layer"] + ["hidden_layer.%d" % i for i in range(n_hidden)] + ["out_layer"]
        out = {}
        for i, n in enumerate(names):
            out["W%d" % i] = sd[n + ".weight"].numpy().astype(np.float32)
            out["b%d" % i] = sd[n + ".bias"].numpy().astype(np.float32)
        out["in_mean"] = np.concatenate([sc["state_scaler"].mean_, sc["control_scaler"].mean_])
        out["in_scale"] = np.concatenate([sc["state_scaler"].scale_, sc["control_scaler"].scale_])
        out["out_mean"], out["out_scale"] = sc["error_scaler"].mean_, sc["error_scaler"].scale_
        net = ref[cls_key](5) if cls_key == "MLP5_3L" else ref[cls_key]()
        net.load_state_dict(sd)
        net = net.double()
        X = np.random.default_rng(0).normal(0, 1.0, (64, 5)) * [3.0, 2.0, 1.0, 1.0, 1.5]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            xin = np.concatenate([sc["state_scaler"].transform(X[:, :3]), sc["control_scaler"].transform(X[:, 3:])], axis=1)
            with torch.no_grad():
                Y = sc["error_scaler"].inverse_transform(net(torch.from_numpy(xin)).numpy())
        out["X"], out["Y_ref"] = X, Y
        out["meta"] = json.dumps(dict(checkpoint="saved_models/" + ckpt, scalers="saved_models/" + scalers, n_hidden=n_hidden,
                                      torch=torch.__version__, numpy=np.__version__,
                                      source="reference checkpoint (float32 weights verbatim) + forward passes of the unmodified "
                                             "reference torch class in float64 with the sklearn scalers applied as "
                                             "test/test_diff_dyna_eval.py:54-56 does"))
        fn = os.path.join(HERE, "trained_" + tag + ".npz")
        np.savez_compressed(fn, **out)
        print("wrote", fn, "%.1f KB" % (os.path.getsize(fn) / 1024))


def make_spline_golden():
    """Courses of the reference's own calc_spline_course (path_generator/cubic_spline_planner.py:311-323) for a few
    waypoint sets (4-12 waypoints, spacings 0.05-0.25): the pin of oracle/spline_oracle.py and, through it, of the
    device-side per-robot path generation."""
    ref = ref_loader.load_reference()
    rng = np.random.default_rng(31)
    out = {}
    meta = []
    for i, (n, ds) in enumerate(((4, 0.1), (5, 0.25), (7, 0.1), (9, 0.05), (12, 0.2), (6, 0.1))):
        ang = np.cumsum(rng.normal(0, 0.6, n))
        step = rng.uniform(0.8, 2.5, n)
        wx = np.cumsum(step * np.cos(ang)) + rng.normal(0, 2.0)
        wy = np.cumsum(step * np.sin(ang)) + rng.normal(0, 2.0)
        cx, cy, cyaw, _, _ = ref["calc_spline_course"](list(wx), list(wy), ds=ds)
        out["wx%d" % i], out["wy%d" % i] = wx, wy
        out["course%d" % i] = np.array([cx, cy, cyaw]).T
        meta.append(dict(n_wp=n, ds=ds, n_pts=len(cx)))
    out["meta"] = json.dumps(dict(cases=meta, numpy=np.__version__,
                                  source="path_generator/cubic_spline_planner.py:calc_spline_course, unmodified"))
    fn = os.path.join(HERE, "spline_courses.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, "%.1f KB" % (os.path.getsize(fn) / 1024))


def main():
    ref = ref_loader.load_reference()
    path = spline_path(ref)
    sigma_dd = np.array([[0.1, 0.0], [0.0, 0.01]])
    w_dd = np.array([5.0, 5.0, 10.0])

    def plant_dd(dt):
        def f(x, u):
            d = ref["DifferentialDrive"](x)
            return d.update_state(dt, x, u)
        return f

    # ---- diff-drive literal class: open-loop ticks from several states, 2 temperatures
    for tag, pe in (("pe1e-4", 1e-4), ("pe0.05", 0.05)):
        K, T = 256, 30
        kw = dict(delta_t=0.1, max_speed=5.0, max_omega=3.14, num_samples_K=K, num_horizons_T=T,
                  param_exploration=pe, param_lambda=1.0, param_alpha=0.2)
        ctrl = ref["MPPIAlgorithms"](ref_path=path, sigma=sigma_dd, stage_cost_weight=w_dd,
                                     terminal_cost_weight=w_dd, visualize_optimal_traj=False,
                                     visualze_sampled_trajs=False, **kw)
        rng = np.random.default_rng(0)
        eps_list = [draw_eps(rng, sigma_dd, K, T) for _ in range(4)]
        rec = run_ticks(ctrl, "prev_way_point_idx", [[0, 0, 0]], eps_list, plant=plant_dd(0.1))
        save("diffdrive_" + tag, dict(kind="diffdrive", **kw), path, eps_list, rec)

    # ---- diff-drive closed loop, longer (small K), temperature 0.05 and the ratchet
    K, T = 64, 12
    kw = dict(delta_t=0.1, max_speed=5.0, max_omega=3.14, num_samples_K=K, num_horizons_T=T,
              param_exploration=0.05, param_lambda=1.0, param_alpha=0.2)
    ctrl = ref["MPPIAlgorithms"](ref_path=path, sigma=sigma_dd, stage_cost_weight=w_dd,
                                 terminal_cost_weight=w_dd, visualize_optimal_traj=False,
                                 visualze_sampled_trajs=False, **kw)
    rng = np.random.default_rng(1)
    eps_list = [draw_eps(rng, sigma_dd, K, T) for _ in range(40)]
    rec = run_ticks(ctrl, "prev_way_point_idx", [[0, 0, 0]], eps_list, plant=plant_dd(0.1))
    save("diffdrive_closed_loop", dict(kind="diffdrive", **kw), path, eps_list, rec)

    # ---- diff-drive + circular obstacles (the obs script's own parameters)
    K, T = 200, 20
    obs = np.array([[2.0, 2.0, 0.4], [3.0, 3.5, 0.4]])
    sigma_obs = np.array([[0.1, 0.0], [0.0, 0.01]])
    w_obs = 10 * np.array([5.0, 6.0, 9.0])
    kw = dict(delta_t=0.1, max_speed=5.0, max_omega=3.14, num_samples_K=K, num_horizons_T=T,
              param_exploration=0.05, param_lambda=10.0, param_alpha=0.98, safety_margin_rate=0.8)
    ctrl = ref["MPPIObs"](ref_path=path, sigma=sigma_obs, stage_cost_weight=w_obs,
                          terminal_cost_weight=w_obs, obstacle_circles=obs,
                          visualize_optimal_traj=False, visualze_sampled_trajs=False, **kw)
    rng = np.random.default_rng(2)
    eps_list = [draw_eps(rng, sigma_obs, K, T) for _ in range(4)]
    states = [[0, 0, 0], [1.6, 1.5, 0.6], [2.3, 1.9, 0.9], [2.9, 2.6, 1.4]]
    rec = run_ticks(ctrl, "prev_way_point_idx", states, eps_list)
    save("diffdrive_obs", dict(kind="diffdrive_obs", **kw), path, eps_list, rec, obstacles=obs)

    # ---- race-car + obstacles, constructor defaults, states = ref_path[i] (+ perturbed)
    from oracle.mppi_oracle import lemniscate_path
    sigma_rc = np.array([[0.5, 0.0], [0.0, 0.1]])
    for tag, alpha, nobs in (("default", 1.0, True), ("alpha0.9", 0.9, True), ("noobs", 1.0, False)):
        K, T = 128, 50 if nobs else 20
        cls = ref["MPPIRacecarController"] if nobs else ref["MPPIRacecarNoObs"]
        ctrl = cls(horizon_step_T=T, number_of_samples_K=K, param_alpha=alpha,
                   visualize_optimal_traj=False, visualze_sampled_trajs=False)
        lp = ctrl.generate_lemniscate_trajectory(100, 10.0) if nobs else lemniscate_path()
        assert np.array_equal(lp.astype(np.float32), lemniscate_path()), "lemniscate restatement drifted"
        ctrl.ref_path = lp.astype(np.float32)
        rng = np.random.default_rng(3)
        eps_list = [draw_eps(rng, sigma_rc, K, T) for _ in range(6)]
        pert = rng.normal(0, [0.3, 0.3, 0.05, 0.5], (6, 4))
        # ticks 0-2 teleport along the path like the reference main (:336-339); 3-5 perturbed
        states = [lp[0], lp[1], lp[2], lp[8] + pert[3], lp[9] + pert[4], lp[10] + pert[5]]
        states = [np.asarray(s, dtype=np.float32) for s in states]
        rec = run_ticks(ctrl, "prev_waypoints_idx", states, eps_list)
        meta = dict(kind="racecar" if nobs else "racecar_noobs", horizon_step_T=T,
                    number_of_samples_K=K, param_alpha=alpha)
        save("racecar_" + tag, meta, lp.astype(np.float32), eps_list, rec,
             obstacles=ctrl.obstacle_circles if nobs else None)

    # ---- visualisation outputs on (A16) incl. the in-place clamp of the nominal (Q9); large sigma so clamping bites
    K, T = 48, 12
    sig_big = np.array([[4.0, 0.0], [0.0, 2.0]])
    kw = dict(delta_t=0.1, max_speed=1.0, max_omega=0.8, num_samples_K=K, num_horizons_T=T,
              param_exploration=0.05, param_lambda=1.0, param_alpha=0.2)
    ctrl = ref["MPPIAlgorithms"](ref_path=path, sigma=sig_big, stage_cost_weight=w_dd, terminal_cost_weight=w_dd,
                                 visualize_optimal_traj=True, visualze_sampled_trajs=True, **kw)
    rng = np.random.default_rng(7)
    eps_list = [draw_eps(rng, sig_big, K, T) for _ in range(3)]
    rec = run_ticks(ctrl, "prev_way_point_idx", [[0, 0, 0]], eps_list, plant=plant_dd(0.1), keep_sampled=True)
    save("diffdrive_viz", dict(kind="diffdrive", viz=True, sigma=sig_big.tolist(), **kw), path, eps_list, rec)
    K, T = 48, 12
    ctrl = ref["MPPIRacecarController"](horizon_step_T=T, number_of_samples_K=K, max_steer_abs=0.1, max_accel_abs=0.5,
                                        param_lambda=5.0, visualize_optimal_traj=True, visualze_sampled_trajs=True)
    ctrl.ref_path = lemniscate_path()
    rng = np.random.default_rng(8)
    eps_list = [draw_eps(rng, sigma_rc, K, T) for _ in range(3)]
    lp = lemniscate_path()
    states = [np.asarray(lp[i] + np.array([0.0, 0.0, 0.0, 0.0]), dtype=np.float32) for i in (20, 21, 22)]
    rec = run_ticks(ctrl, "prev_waypoints_idx", states, eps_list, keep_sampled=True)
    save("racecar_viz", dict(kind="racecar", viz=True, horizon_step_T=T, number_of_samples_K=K, param_alpha=1.0,
                             max_steer_abs=0.1, max_accel_abs=0.5, param_lambda=5.0), lp, eps_list, rec,
         obstacles=ctrl.obstacle_circles)

    make_extras()
    make_mlp_golden()
    make_spline_golden()

    # ---- literal filter operators as matrices (Q7)
    Ms = {}
    for T in (10, 12, 20, 30, 50):
        Ms["diffdrive_T%d" % T] = ref["MPPIAlgorithms"]._moving_average_filter(None, np.eye(T), 10)
        Ms["racecar_T%d" % T] = ref["MPPIRacecarController"]._moving_average_filter(
            None, np.eye(T, dtype=np.float32), 10)
    np.savez_compressed(os.path.join(HERE, "filter_matrices.npz"), **Ms)
    np.savez_compressed(os.path.join(HERE, "paths.npz"), spline=path, lemniscate=lemniscate_path())
    print("done")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "extras":
        make_extras()               # only the SURVEY 8f row-3 fixtures (the others stay byte-identical)
    elif len(sys.argv) > 1 and sys.argv[1] == "mlp":
        make_mlp_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "spline":
        make_spline_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "trained":
        make_trained_mlp_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "c1":
        make_c1_literal()
    else:
        main()
