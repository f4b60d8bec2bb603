"""Host mirror of how test/test_mppi_diff_obs.py drives its controller: a diff-drive MPPI whose running cost is a
quadratic pose error to a target, a quadratic control effort and an exponential soft penalty around circular
obstacles that MOVE at constant velocity (reference :14-20 `get_obstacle_positions`, :28-42 `dynamics`, :44-66
`running_cost`), with `command(state)` returning the action to apply and `get_trajectories(state)` returning the
optimal rollout plus the top max(10, K/10) sampled rollouts by cost (:88-111).

The reference script gets its MPPI update from `pytorch_mppi.MPPI`, a package that is neither vendored nor pinned
in the reference tree, and passes `dynamics` / `running_cost` as Python callbacks.  Callbacks cannot cross into a
CUDA kernel, so this class takes the PARAMETERS of those two functions instead and evaluates them inside the tick
kernel (cost kind `target_soft`); they are pinned against the reference's own functions executed through its
`_compute_rollout_costs` loop (tests/golden/diffdrive_target_soft.npz).  The update rule is this repo's MPPI tick
(weights exp(-(S-min S)/lambda_), weighted noise, edge-padded moving average, shift): parity with pytorch_mppi's
update is UNPINNED and not claimed."""
import numpy as np

from ._base import ControllerBase


class MPPIDynamicObstacles(ControllerBase):
    _out_dtype = np.float32

    def __init__(self, noise_sigma, num_samples, horizon, lambda_, u_min, u_max, delta_t=0.05,
                 target=(6.0, 6.0, 1.57), Q=(30.0, 5.0, 9.0), R=(0.1, 0.1),
                 obstacle_positions=((5.0, 4.0), (3.5, 3.5)),
                 obstacle_velocities=((0.018, 0.009), (-0.009, 0.009)),
                 safety_distance=2.0, obstacle_weight=100.0, terminal_Q=(0.0, 0.0, 0.0),
                 *, seed=0, device=0, rank=0, world=1):
        u_min, u_max = np.asarray(u_min, float).reshape(2), np.asarray(u_max, float).reshape(2)
        if not np.allclose(u_min, -u_max):
            raise ValueError("the tick kernel clamps symmetrically: u_min must equal -u_max")
        self.delta_t = float(delta_t)
        self.dim_x, self.dim_u = 3, 2
        self.nx, self.nu = 3, 2
        self.T, self.K = int(horizon), int(num_samples)
        self.lambda_ = float(lambda_)
        self.noise_sigma = np.asarray(noise_sigma, dtype=np.float64).reshape(2, 2)
        self.u_min, self.u_max = u_min, u_max
        self.target = np.asarray(target, dtype=np.float64).reshape(3)
        self._obs_pos = np.asarray(obstacle_positions, dtype=np.float64).reshape(-1, 2)
        self._obs_vel = np.asarray(obstacle_velocities, dtype=np.float64).reshape(-1, 2)
        self.visualize_optimal_traj = self.visualze_sampled_trajs = False
        self._init_engine(
            ref_path=None, seed=seed, device=device, rank=rank, world=world,
            model="diffdrive", K=self.K, T=self.T, dt=self.delta_t, u_max=tuple(u_max), sigma=self.noise_sigma,
            stage_w=np.asarray(Q, float), term_w=np.asarray(terminal_Q, float),
            param_exploration=0.0, param_lambda=self.lambda_, param_alpha=1.0, temperature=self.lambda_,
            window=20, cost_mode="sum", waypoint_mode="frozen", filter_kind="racecar", yaw_wrap=False,
            collision="none", cost_kind="target_soft", goal=self.target, ctrl_w=tuple(R),
            soft_obs_weight=obstacle_weight, soft_obs_safety=safety_distance)
        self._engine.set_moving_obstacles(self._obs_pos, self._obs_vel)
        self._engine.set_keep_costs(True)            # get_trajectories sorts the samples by cost
        self._last = None

    def set_obstacles(self, positions, velocities):
        """Obstacle positions at horizon time 0 and their velocities; call before each tick to advance the scene the
        way the script's simulation loop does with `current_time` (:334-347)."""
        self._obs_pos = np.asarray(positions, dtype=np.float64).reshape(-1, 2)
        self._obs_vel = np.asarray(velocities, dtype=np.float64).reshape(-1, 2)
        self._engine.set_moving_obstacles(self._obs_pos, self._obs_vel)

    def command(self, state, noise=None):
        """One control tick; returns the action to apply now: row 0 of the updated nominal BEFORE the shift (what
        `MPPI.command` returns, :80), not the post-shift row the reference's own classes hand back (quirk Q8)."""
        x = np.asarray(state, dtype=np.float64).reshape(3)
        self._last = (x, noise)
        self._tick_impl(x, noise)
        return self._engine.stats()["u_first"].copy()

    def get_trajectories(self, state=None, n_top=None):
        """(optimal_traj (T,3), sampled_traj_list (n_top,T,3)) of the last `command`, the samples in ascending cost
        order, n_top = max(10, K/10) by default (:100-101), controls indexed t (:94,:108)."""
        import torch
        if self._last is None:
            self.command(state)
        x, noise = self._last
        n_top = max(10, self.K // 10) if n_top is None else int(n_top)
        n_top = min(n_top, self._K_local)
        dev = "cuda:%d" % self._engine.device
        d_samp = torch.empty(n_top, self.T, 3, dtype=torch.float32, device=dev)
        d_idx = torch.empty(n_top, dtype=torch.int32, device=dev)
        d_cost = torch.empty(n_top, dtype=torch.float32, device=dev)
        d_eps = None
        if noise is not None:
            from ._base import _to_device_noise
            d_eps = _to_device_noise(noise, self._K_local, self.T, self._engine.device)
        opt = self._engine.top_trajectories(x, d_samp, n_top, d_idx, d_cost, True, 0, d_eps, self.seed, self._tick - 1)
        self.last_top_idx, self.last_top_cost = d_idx.cpu().numpy(), d_cost.cpu().numpy()
        return opt, d_samp.cpu().numpy()

    def _viz_gates(self):
        return False, False
