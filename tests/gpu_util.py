"""Helpers for the `-m gpu` parity tests: build an engine / drop-in controller configured like an
oracle spec.  Everything goes through the C ABI of libmppi_b200.so."""
import numpy as np

from mppi_b200.engine import MPPIEngine


def engine_from_spec(spec, path, n_robots=1, device=0, **kw):
    eng = MPPIEngine(
        model=spec.model, K=spec.K, T=spec.T, dt=spec.dt, u_max=spec.u_max, sigma=spec.sigma,
        stage_w=spec.stage_w, term_w=spec.term_w, param_exploration=spec.param_exploration,
        param_lambda=spec.param_lambda, param_alpha=spec.param_alpha, temperature=spec.temperature,
        window=spec.window, cost_mode=spec.cost_mode, waypoint_mode=spec.waypoint_mode,
        filter_kind=spec.filter_kind, yaw_wrap=spec.yaw_wrap, collision=spec.collision,
        margin=spec.margin, wheel_base=spec.wheel_base,
        robot_radius=spec.robot_radius, vehicle_l=spec.vehicle_l, vehicle_w=spec.vehicle_w,
        n_robots=n_robots, device=device, cost_kind=spec.cost_kind, goal=spec.goal, ctrl_w=spec.ctrl_w,
        soft_obs_weight=spec.soft_w, soft_obs_safety=spec.soft_sd,
        **dict(dict(obstacles=None if spec.cost_kind == "target_soft" else spec.obstacles), **kw))
    if spec.cost_kind == "path":
        eng.set_ref_path(path)
    if spec.cost_kind == "target_soft":
        eng.set_moving_obstacles(spec.obstacles, spec.obs_vel)
    return eng


def cost_mismatch(S_gpu, S_ref, rtol=1e-5, atol=1e-6):
    """Fraction of samples outside tolerance and the worst relative error."""
    S_gpu = np.asarray(S_gpu, np.float64)
    S_ref = np.asarray(S_ref, np.float64)
    err = np.abs(S_gpu - S_ref)
    bad = err > (atol + rtol * np.abs(S_ref))
    rel = err / np.maximum(np.abs(S_ref), 1e-12)
    return float(bad.mean()), float(rel.max())
