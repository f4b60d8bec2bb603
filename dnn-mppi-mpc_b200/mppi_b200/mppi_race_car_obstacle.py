"""Drop-in for controllers/mppi_race_car_obstacle.py:MPPIRacecarController (reference :10-274):
kinematic-bicycle MPPI with the footprint-vs-circle collision penalty.  Same constructor
kwargs and defaults (:11-30), same `_calc_control_input(observed_x)` 4-tuple."""
import numpy as np

from ._base import ControllerBase


class MPPIRacecarController(ControllerBase):
    _out_dtype = np.float32
    _idx_attr = "prev_waypoints_idx"
    _with_obstacles = True

    def __init__(self, delta_t=0.05, wheel_base=2.5, max_steer_abs=0.523, max_accel_abs=2.000,
                 ref_path=np.array([[0.0, 0.0, 0.0, 1.0], [10.0, 0.0, 0.0, 1.0]]),
                 horizon_step_T=10, number_of_samples_K=100, param_exploration=0.01,
                 param_lambda=50.0, param_alpha=1.0,
                 sigma=np.array([[0.5, 0.0], [0.0, 0.1]]),
                 stage_cost_weight=np.array([50.0, 50.0, 1.0, 20.0]),
                 terminal_cost_weight=np.array([50.0, 50.0, 1.0, 20.0]),
                 obstacle_circles=np.array([[5.0, 5.0, 1.0], [7.0, 7.0, 1.0]]),
                 collision_safety_margin_rat=1.5,
                 visualize_optimal_traj=True, visualze_sampled_trajs=True,
                 *, seed=0, device=0, rank=0, world=1):
        self.dim_x, self.dim_u = 4, 2
        self.T, self.K = int(horizon_step_T), int(number_of_samples_K)
        self.param_exploration = param_exploration
        self.param_lambda = param_lambda
        self.param_alpha = param_alpha
        self.param_gamma = param_lambda * (1.0 - param_alpha)                  # :40
        self.Sigma = np.asarray(sigma).astype(np.float32)
        self.stage_cost_weight = np.asarray(stage_cost_weight).astype(np.float32)
        self.terminal_cost_weight = np.asarray(terminal_cost_weight).astype(np.float32)
        self.visualize_optimal_traj = visualize_optimal_traj
        self.visualze_sampled_trajs = visualze_sampled_trajs
        self.delta_t = delta_t
        self.wheel_base = wheel_base
        self.max_steer_abs = max_steer_abs
        self.max_accel_abs = max_accel_abs
        self.vehicle_w, self.vehicle_l = 3.0, 4.0                              # :53-54
        self._obstacle_circles = (np.asarray(obstacle_circles, dtype=np.float64).reshape(-1, 3)
                                  if self._with_obstacles else np.zeros((0, 3)))
        self.collision_safety_margin_rate = collision_safety_margin_rat
        self._init_engine(
            ref_path=np.asarray(ref_path).astype(np.float32), seed=seed, device=device, rank=rank, world=world,
            model="bicycle", K=self.K, T=self.T, dt=delta_t, u_max=(max_steer_abs, max_accel_abs),
            sigma=self.Sigma, stage_w=self.stage_cost_weight, term_w=self.terminal_cost_weight,
            param_exploration=param_exploration, param_lambda=param_lambda, param_alpha=param_alpha,
            temperature=param_lambda,                                             # Q2 (:222,:224)
            window=200,                                                           # :175
            cost_mode="sum", waypoint_mode="frozen", filter_kind="racecar", yaw_wrap=True,
            collision="footprint" if self._with_obstacles else "none",
            obstacles=self._obstacle_circles, margin=collision_safety_margin_rat,
            wheel_base=wheel_base, vehicle_l=self.vehicle_l, vehicle_w=self.vehicle_w,
            clamp_nominal=bool(visualize_optimal_traj))                           # Q9 (:112-115)

    prev_waypoints_idx = property(ControllerBase._get_idx, ControllerBase._set_idx)

    def _viz_gates(self):
        return bool(self.visualize_optimal_traj), bool(self.visualze_sampled_trajs)

    @property
    def obstacle_circles(self):
        return self._obstacle_circles

    @obstacle_circles.setter
    def obstacle_circles(self, v):
        self._obstacle_circles = np.asarray(v, dtype=np.float64).reshape(-1, 3)
        self._engine.set_obstacles(self._obstacle_circles)

    def _calc_control_input(self, observed_x, noise=None):
        """One control tick (reference :65-131).  `noise`: optional injected (K,T,2) epsilon."""
        return self._tick_impl(observed_x, noise)

    def generate_lemniscate_trajectory(self, num_points, radius):
        """Figure-eight reference course (x, y, yaw, v=5), as the reference helper (:288-299)."""
        t = np.linspace(0, 2 * np.pi, num_points, dtype=np.float32)
        x = radius * np.cos(t) / (1 + np.sin(t) ** 2)
        y = radius * np.sin(t) * np.cos(t) / (1 + np.sin(t) ** 2)
        yaw = np.arctan2(np.gradient(y), np.gradient(x))
        return np.stack([x, y, yaw, np.ones_like(t) * 5.0], axis=1)
