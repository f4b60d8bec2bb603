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

    def set_dynamics(self, dynamics):
        """Learned dynamics x+ = x + dt*([v cos th, v sin th, w] + MLP(x)).  `dynamics` is a
        dnn/simple_mlp.py-shaped torch module (input_layer, hidden_layer[0..1], output_layer) or a dict
        {'W0','b0',...,'W3','b3'} of nn.Linear-layout arrays.  Requires waypoint_mode='frozen'."""
        if hasattr(dynamics, "state_dict"):
            sd = {k: v.detach().cpu().numpy() for k, v in dynamics.state_dict().items()}
            names = ["input_layer", "hidden_layer.0", "hidden_layer.1", "output_layer"]
            W = [sd[n + ".weight"] for n in names]
            b = [sd[n + ".bias"] for n in names]
        else:
            W = [np.asarray(dynamics["W%d" % i]) for i in range(4)]
            b = [np.asarray(dynamics["b%d" % i]) for i in range(4)]
        shapes = [(512, 3), (512, 512), (512, 512), (3, 512)]
        for w, sh in zip(W, shapes):
            if tuple(w.shape) != sh:
                raise ValueError("MLP weights must have the dnn/simple_mlp.py shapes %s" % (shapes,))
        self._engine.set_mlp(W, b)

    def _calc_input_control(self, observed_x, noise=None):
        """One control tick (reference :87-165).  `noise`: optional injected (K,T,2) epsilon."""
        return self._tick_impl(observed_x, noise)
