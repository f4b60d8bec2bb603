"""Drop-in for test/mppi_differential_drive_obs.py:MPPIAlgorithms -- the goal-point diff-drive MPPI
(reference :42-313).  It is the controllers/mppi_differential_drive.py tick (:90-166, same noise, clamp, Euler
unicycle, stage cost overwritten by `=`, weights with temperature `param_exploration`, weighted noise,
moving-average filter incl. its tail bug, update, shift) with three differences:
  * no reference path and no waypoint index: the state cost is `w0*|xy-goal|^2 + w1*wrap(atan2(dy,dx)-yaw)^2`
    (:202-232);
  * the circle-circle obstacle penalty `1e10` inside both the stage and the terminal cost (:219,:230,:301-313);
  * the weights have two entries (`stage_cost_weight = 10*[5, 9]`, :410-411).
Same constructor kwargs, same `_calc_input_control(observed_x)` 4-tuple."""
import numpy as np

from ._base import ControllerBase


class MPPIAlgorithms(ControllerBase):
    _out_dtype = np.float64

    def __init__(self, delta_t, goal_point, max_speed, max_omega, num_samples_K, num_horizons_T,
                 param_exploration, param_lambda, param_alpha, sigma, stage_cost_weight,
                 terminal_cost_weight, obstacle_circles, safety_margin_rate,
                 visualize_optimal_traj=True, visualze_sampled_trajs=True,
                 *, seed=0, cost_mode="last", temperature=None, device=0, rank=0, world=1):
        self.delta_t = delta_t
        self._goal_point = np.asarray(goal_point, dtype=np.float64).reshape(2)
        self.max_speed = max_speed
        self.max_omega = max_omega
        self.dim_x, self.dim_u = 3, 2
        self.T, self.K = int(num_horizons_T), int(num_samples_K)
        self.param_exploration = param_exploration
        self.param_lambda = param_lambda
        self.param_alpha = param_alpha
        self.param_gamma = param_lambda * (1.0 - param_alpha)                  # :75
        self.Sigma = np.asarray(sigma, dtype=np.float64)
        self.stage_cost_weight = np.asarray(stage_cost_weight, dtype=np.float64)
        self.terminal_cost_weight = np.asarray(terminal_cost_weight, dtype=np.float64)
        self.visualize_optimal_traj = visualize_optimal_traj
        self.visualze_sampled_trajs = visualze_sampled_trajs
        self._obstacle_circles = np.asarray(obstacle_circles, dtype=np.float64).reshape(-1, 3)
        self.safefy_margin_rate = safety_margin_rate          # (sic) attribute name of the reference (:84)
        self._init_engine(
            ref_path=None, seed=seed, device=device, rank=rank, world=world,
            model="diffdrive", K=self.K, T=self.T, dt=delta_t, u_max=(max_speed, max_omega),
            sigma=self.Sigma, stage_w=self.stage_cost_weight, term_w=self.terminal_cost_weight,
            param_exploration=param_exploration, param_lambda=param_lambda, param_alpha=param_alpha,
            temperature=param_exploration if temperature is None else temperature,   # :175,:178
            window=20, cost_mode=cost_mode, waypoint_mode="frozen", filter_kind="diffdrive",
            yaw_wrap=False, collision="circle" if len(self._obstacle_circles) else "none",
            obstacles=self._obstacle_circles, margin=safety_margin_rate,
            clamp_nominal=bool(visualze_sampled_trajs),                              # :145-148
            cost_kind="goal", goal=self._goal_point)

    @property
    def goal_point(self):
        return self._goal_point

    @goal_point.setter
    def goal_point(self, g):
        self._goal_point = np.asarray(g, dtype=np.float64).reshape(2)
        self._engine.set_goal(self._goal_point)

    @property
    def obstacle_circles(self):
        return self._obstacle_circles

    @obstacle_circles.setter
    def obstacle_circles(self, v):
        self._obstacle_circles = np.asarray(v, dtype=np.float64).reshape(-1, 3)
        self._engine.set_obstacles(self._obstacle_circles)

    def _calc_input_control(self, observed_x, noise=None):
        """One control tick (reference :90-166).  `noise`: optional injected (K,T,2) epsilon."""
        return self._tick_impl(observed_x, noise)
