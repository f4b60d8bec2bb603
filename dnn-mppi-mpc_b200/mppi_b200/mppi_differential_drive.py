"""Drop-in for controllers/mppi_differential_drive.py:MPPIAlgorithms (reference :42-289).

Same constructor kwargs, same per-step method `_calc_input_control(observed_x)` returning
`(u0, u_seq, optimal_traj, sampled_traj_list)`; the K x T Python loops are replaced by one
call into libmppi_b200.so.  Defaults reproduce the class literally (quirks Q1-Q8 of
SURVEY.md 8a): stage cost overwritten (`cost_mode='last'`), temperature =
`param_exploration`, waypoint index mutated during the rollouts (`waypoint_mode='strict'`).
`cost_mode='sum'` / `waypoint_mode='frozen'` select the throughput modes."""
import numpy as np

from ._base import ControllerBase


class MPPIAlgorithms(ControllerBase):
    _out_dtype = np.float64
    _collision = "none"

    def __init__(self, delta_t, ref_path, max_speed, max_omega, num_samples_K, num_horizons_T,
                 param_exploration, param_lambda, param_alpha, sigma, stage_cost_weight,
                 terminal_cost_weight, visualize_optimal_traj=True, visualze_sampled_trajs=True,
                 *, seed=0, cost_mode="last", waypoint_mode="strict", temperature=None, dynamics=None,
                 device=0, rank=0, world=1,
                 _obstacles=None, _margin=1.0):
        self.delta_t = delta_t
        self.max_speed = max_speed
        self.max_omega = max_omega
        self.dim_x, self.dim_u = 3, 2
        self.T, self.K = int(num_horizons_T), int(num_samples_K)
        self.param_exploration = param_exploration
        self.param_lambda = param_lambda
        self.param_alpha = param_alpha
        self.param_gamma = param_lambda * (1.0 - param_alpha)                  # :74
        self.Sigma = np.asarray(sigma, dtype=np.float64)
        self.stage_cost_weight = np.asarray(stage_cost_weight, dtype=np.float64)
        self.terminal_cost_weight = np.asarray(terminal_cost_weight, dtype=np.float64)
        self.visualize_optimal_traj = visualize_optimal_traj
        self.visualze_sampled_trajs = visualze_sampled_trajs
        self._init_engine(
            ref_path=ref_path, seed=seed, device=device, rank=rank, world=world,
            model="diffdrive" if dynamics is None else "diffdrive_mlp", K=self.K, T=self.T, dt=delta_t, u_max=(max_speed, max_omega),
            sigma=self.Sigma, stage_w=self.stage_cost_weight, term_w=self.terminal_cost_weight,
            param_exploration=param_exploration, param_lambda=param_lambda, param_alpha=param_alpha,
            temperature=param_exploration if temperature is None else temperature,   # Q2 (:175,:178)
            window=20,                                                            # :204
            cost_mode=cost_mode, waypoint_mode=waypoint_mode, filter_kind="diffdrive",
            yaw_wrap=False, collision=self._collision, obstacles=_obstacles, margin=_margin,
            clamp_nominal=bool(visualze_sampled_trajs))                           # Q9 (:145-148)

        if dynamics is not None:
            self.set_dynamics(dynamics)

    prev_way_point_idx = property(ControllerBase._get_idx, ControllerBase._set_idx)

    def set_dynamics(self, dynamics, scalers=None):
        """Learned dynamics x+ = x + dt*([v cos th, v sin th, w] + residual).  `dynamics` is a torch
        module / state dict shaped like dnn/simple_mlp.py (3 inputs: residual = MLP(x)) or like the trained
        saved_models/mlp_diff*.pth (5 inputs: residual = MLP([x; u])), or a dict {'W0','b0',...,'W3','b3'} of
        nn.Linear-layout arrays.  `scalers`: the {'state_scaler','control_scaler','error_scaler'} dict saved next to the
        trained models (saved_models/scalers_*.pth) or {'in_mean','in_scale','out_mean','out_scale'} arrays:
        residual = out_scale * MLP(([x; u] - in_mean) / in_scale) + out_mean (test/test_diff_dyna_eval.py:54-56).
        Requires waypoint_mode='frozen'."""
        if hasattr(dynamics, "state_dict"):
            dynamics = dynamics.state_dict()
        if any(str(k).endswith(".weight") for k in dynamics):
            # torch state dict: dnn/simple_mlp.py names its last layer `output_layer`, the trained 5-input models
            # (simulation/bullet_differential_drive_dnn.py:37-60, saved_models/mlp_diff*.pth) `out_layer`
            sd = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in dynamics.items()}
            n_hidden = len([k for k in sd if k.startswith("hidden_layer.") and k.endswith(".weight")])
            last = "output_layer" if "output_layer.weight" in sd else "out_layer"
            names = ["input_layer"] + ["hidden_layer.%d" % i for i in range(n_hidden)] + [last]
            W = [sd[n + ".weight"] for n in names]
            b = [sd[n + ".bias"] for n in names]
        else:
            n = len([k for k in dynamics if str(k).startswith("W")])
            W = [np.asarray(dynamics["W%d" % i]) for i in range(n)]
            b = [np.asarray(dynamics["b%d" % i]) for i in range(n)]
            if scalers is None and "in_scale" in dynamics:
                scalers = {k: dynamics[k] for k in ("in_mean", "in_scale", "out_mean", "out_scale") if k in dynamics}
        n_in = W[0].shape[1]
        shapes = [(512, n_in)] + [(512, 512)] * (len(W) - 2) + [(3, 512)]
        if n_in not in (3, 5) or [tuple(w.shape) for w in W] != shapes:
            raise ValueError("MLP weights must be (512,3|5), (512,512)..., (3,512) like dnn/simple_mlp.py / saved_models/mlp_diff*.pth")
        kw = {}
        if scalers is not None:
            if "state_scaler" in scalers:            # the dict torch.load("saved_models/scalers_*.pth") returns
                ss, cs, es = scalers["state_scaler"], scalers["control_scaler"], scalers["error_scaler"]
                kw = dict(in_mean=np.concatenate([ss.mean_, cs.mean_])[:n_in], in_scale=np.concatenate([ss.scale_, cs.scale_])[:n_in],
                          out_mean=es.mean_, out_scale=es.scale_)
            else:
                kw = {k: np.asarray(scalers[k], dtype=np.float64) for k in ("in_mean", "in_scale", "out_mean", "out_scale") if k in scalers}
        self._engine.set_mlp(W, b, **kw)

    def _calc_input_control(self, observed_x, noise=None):
        """One control tick (reference :87-165).  `noise`: optional injected (K,T,2) epsilon."""
        return self._tick_impl(observed_x, noise)
