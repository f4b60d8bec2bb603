"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/mppi_oracle.c (the FP64 C
restatement).  Same import restrictions as oracle/mppi_oracle.py."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import mppi_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmppi_oracle.so")


class _Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("model", "K", "T", "n_exploit", "window", "cost_mode",
                                       "waypoint_mode", "filter_kind", "yaw_wrap", "collision",
                                       "n_obstacles", "n_path", "path_cols")] + \
               [("dt", C.c_double), ("wheel_base", C.c_double), ("u_max", C.c_double * 2),
                ("gamma", C.c_double), ("temperature", C.c_double), ("sig_inv", C.c_double * 4),
                ("chol", C.c_double * 4), ("stage_w", C.c_double * 4), ("term_w", C.c_double * 4),
                ("margin", C.c_double), ("robot_radius", C.c_double), ("vehicle_l", C.c_double),
                ("vehicle_w", C.c_double), ("goal", C.c_double * 3), ("ctrl_w", C.c_double * 2),
                ("soft_w", C.c_double), ("soft_sd", C.c_double), ("obs_vel", C.c_double * 32), ("cost_kind", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "mppi_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmppi_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mppi_oracle_max_threads.restype = C.c_int
    return _lib


def _cfg(spec: orc.MPPISpec, path):
    c = _Cfg()
    c.model = 1 if spec.model == "bicycle" else 0
    if spec.model == "diffdrive_mlp":
        raise ValueError("the C oracle has no MLP variant; use oracle.mppi_oracle.tick_vec")
    c.K, c.T, c.n_exploit, c.window = spec.K, spec.T, spec.n_exploit(), spec.window
    c.cost_mode = 0 if spec.cost_mode == "last" else 1
    c.waypoint_mode = 0 if spec.waypoint_mode == "strict" else 1
    c.filter_kind = 0 if spec.filter_kind == "diffdrive" else 1
    c.yaw_wrap = int(spec.yaw_wrap)
    c.collision = {"none": 0, "circle": 1, "footprint": 2}[spec.collision]
    c.n_obstacles = int(spec.obstacles.shape[0])
    c.n_path, c.path_cols = path.shape if path is not None else (0, 3)
    c.cost_kind = {"path": 0, "goal": 1, "target_soft": 2}[spec.cost_kind]
    c.goal[:] = list(np.asarray(spec.goal, float).reshape(-1)[:3]) + [0.0] * (3 - min(3, np.asarray(spec.goal).size))
    c.ctrl_w[:] = list(spec.ctrl_w)
    c.soft_w, c.soft_sd = spec.soft_w, spec.soft_sd
    ov = np.zeros(32); ov[:spec.obs_vel.size] = np.asarray(spec.obs_vel, float).reshape(-1)
    c.obs_vel[:] = list(ov)
    c.dt, c.wheel_base = spec.dt, spec.wheel_base
    c.u_max[:] = spec.u_max
    c.gamma, c.temperature = spec.gamma, spec.temperature
    c.sig_inv[:] = np.linalg.inv(spec.sigma).reshape(-1)
    c.chol[:] = np.linalg.cholesky(spec.sigma).reshape(-1)
    sw = np.zeros(4); sw[:spec.stage_w.size] = spec.stage_w
    tw = np.zeros(4); tw[:spec.term_w.size] = spec.term_w
    c.stage_w[:] = sw
    c.term_w[:] = tw
    c.margin, c.robot_radius = spec.margin, spec.robot_radius
    c.vehicle_l, c.vehicle_w = spec.vehicle_l, spec.vehicle_w
    return c


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def costs(spec, path, U, idx, x0, eps=None, seed=0, tick=0, k_offset=0, nthreads=0, n_exploit=None):
    """Per-sample costs S (K,) float64, index after step 1, index after the tick.  For a shard of a
    larger sample set pass `k_offset` and the GLOBAL explore/exploit threshold `n_exploit` (Q6)."""
    path = np.ascontiguousarray(path, np.float64) if path is not None else None
    c = _cfg(spec, path)
    if path is None:
        path = np.zeros((1, 3))
    if n_exploit is not None:
        c.n_exploit = int(n_exploit)
    U = np.ascontiguousarray(U, np.float64)
    x0 = np.ascontiguousarray(x0, np.float64)
    obs = np.ascontiguousarray(spec.obstacles, np.float64).reshape(-1, 2 if spec.cost_kind == "target_soft" else 3)
    S = np.zeros(spec.K)
    i1, i2 = C.c_int(), C.c_int()
    if eps is not None:
        eps = np.ascontiguousarray(eps, np.float32)
        assert eps.shape == (spec.K, spec.T, 2)
    lib().mppi_oracle_costs(C.byref(c), _p(path, C.c_double), _p(obs, C.c_double), _p(U, C.c_double),
                            C.c_int(int(idx)), _p(x0, C.c_double),
                            _p(eps, C.c_float) if eps is not None else None,
                            C.c_uint64(seed), C.c_uint32(tick), C.c_uint32(k_offset), C.c_int(nthreads),
                            _p(S, C.c_double), C.byref(i1), C.byref(i2))
    return S, i1.value, i2.value


def update(spec, path, U, S, eps=None, seed=0, tick=0, k_offset=0, nthreads=0):
    path = np.ascontiguousarray(path, np.float64) if path is not None else None
    c = _cfg(spec, path)
    if path is None:
        path = np.zeros((1, 3))
    U = np.ascontiguousarray(U, np.float64)
    S = np.ascontiguousarray(S, np.float64)
    w_eps = np.zeros((spec.T, 2))
    U_after = np.zeros((spec.T, 2))
    rho, eta = C.c_double(), C.c_double()
    if eps is not None:
        eps = np.ascontiguousarray(eps, np.float32)
    lib().mppi_oracle_update(C.byref(c), _p(U, C.c_double), _p(S, C.c_double),
                             _p(eps, C.c_float) if eps is not None else None,
                             C.c_uint64(seed), C.c_uint32(tick), C.c_uint32(k_offset), C.c_int(nthreads),
                             _p(w_eps, C.c_double), _p(U_after, C.c_double), C.byref(rho), C.byref(eta))
    return dict(w_eps=w_eps, U_after=U_after, u0=U_after[0].copy(), rho=rho.value, eta=eta.value)


def tick(spec, path, U, idx, x0, eps=None, seed=0, tick=0, nthreads=0):
    S, _, s_end = costs(spec, path, U, idx, x0, eps, seed, tick, 0, nthreads)
    out = update(spec, path, U, S, eps, seed, tick, 0, nthreads)
    out.update(S=S, idx_after=s_end)
    return out


def filter_matrix(T, kind):
    M = np.zeros((T, T))
    lib().mppi_oracle_filter_matrix(C.c_int(T), C.c_int(0 if kind == "diffdrive" else 1), _p(M, C.c_double))
    return M


def philox(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().mppi_oracle_philox(_p(ctr, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32))
    return out


def max_threads():
    return lib().mppi_oracle_max_threads()
