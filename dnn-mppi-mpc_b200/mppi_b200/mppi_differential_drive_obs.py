"""Drop-in for controllers/mppi_differential_drive_obs.py:MPPIAlgorithms -- the diff-drive
controller plus the circle-circle obstacle penalty (reference :58-59,:242,:257,:301-313)."""
import numpy as np

from .mppi_differential_drive import MPPIAlgorithms as _Base


class MPPIAlgorithms(_Base):
    _collision = "circle"

    def __init__(self, delta_t, ref_path, max_speed, max_omega, num_samples_K, num_horizons_T,
                 param_exploration, param_lambda, param_alpha, sigma, stage_cost_weight,
                 terminal_cost_weight, obstacle_circles, safety_margin_rate,
                 visualize_optimal_traj=True, visualze_sampled_trajs=True, **kw):
        self._obstacle_circles = np.asarray(obstacle_circles, dtype=np.float64).reshape(-1, 3)
        self.safefy_margin_rate = safety_margin_rate          # (sic) attribute name of the reference (:85)
        super().__init__(delta_t, ref_path, max_speed, max_omega, num_samples_K, num_horizons_T,
                         param_exploration, param_lambda, param_alpha, sigma, stage_cost_weight,
                         terminal_cost_weight, visualize_optimal_traj, visualze_sampled_trajs,
                         _obstacles=self._obstacle_circles, _margin=safety_margin_rate, **kw)

    @property
    def obstacle_circles(self):
        return self._obstacle_circles

    @obstacle_circles.setter
    def obstacle_circles(self, v):
        self._obstacle_circles = np.asarray(v, dtype=np.float64).reshape(-1, 3)
        self._engine.set_obstacles(self._obstacle_circles)
