"""Drop-in for controllers/mppi_race_car.py:MPPIRacecarController -- the race-car controller
without obstacles.  The reference raises IndexError when the nearest waypoint reaches the
end of the path (:65); so does this class."""
from .mppi_race_car_obstacle import MPPIRacecarController as _WithObstacles


class MPPIRacecarController(_WithObstacles):
    _with_obstacles = False

    def __init__(self, *args, **kw):
        kw.pop("obstacle_circles", None)
        kw.pop("collision_safety_margin_rat", None)
        super().__init__(*args, **kw)

    def _calc_control_input(self, observed_x, noise=None):
        out = self._tick_impl(observed_x, noise)
        if self.prev_waypoints_idx >= self.ref_path.shape[0] - 1:
            raise IndexError("[ERROR] Reached the end of the reference path.")
        return out
