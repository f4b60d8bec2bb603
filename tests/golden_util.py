"""Loads tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified
reference classes) and maps each file's recorded constructor kwargs to an oracle spec."""
import json
import os

import numpy as np

from oracle import mppi_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DIFFDRIVE_CASES = ["diffdrive_pe1e-4", "diffdrive_pe0.05", "diffdrive_closed_loop", "diffdrive_obs"]
RACECAR_CASES = ["racecar_default", "racecar_alpha0.9", "racecar_noobs"]
VIZ_CASES = ["diffdrive_viz", "racecar_viz"]
GOAL_CASES = ["diffdrive_goal"]                 # test/mppi_differential_drive_obs.py (goal-point class)
TARGET_SOFT_CASE = "diffdrive_target_soft"       # test/test_mppi_diff_obs.py (running cost only; no class to pin)
ALL_CASES = DIFFDRIVE_CASES + RACECAR_CASES + GOAL_CASES


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = json.loads(str(z["meta"]))
        self.path = z["path"]
        self.eps = z["eps"]                       # (ticks, K, T, 2) float32
        self.obstacles = z["obstacles"] if "obstacles" in z.files else None
        self.rec = {k: z[k] for k in ("x0", "U0", "idx0", "S", "w", "w_eps", "w_eps_filt",
                                      "U_after", "u0", "idx_after", "optimal_traj", "sampled_traj", "V") if k in z.files}
        self.n_ticks = self.eps.shape[0]

    def spec(self, **override):
        m = self.meta
        if m["kind"] == "diffdrive_goal":
            s = orc.goal_spec(K=m["num_samples_K"], T=m["num_horizons_T"], goal=m["goal"], dt=m["delta_t"],
                              max_speed=m["max_speed"], max_omega=m["max_omega"],
                              param_exploration=m["param_exploration"], param_lambda=m["param_lambda"],
                              param_alpha=m["param_alpha"], obstacles=self.obstacles,
                              margin=m["safety_margin_rate"])
        elif m["kind"] == "diffdrive_target_soft":
            s = orc.target_soft_spec(K=m["K"], T=m["T"], target=m["target"], Q=m["Q"], R=m["R"], dt=m["delta_t"],
                                     u_max=m["u_max"], sigma=np.array(m["sigma"]), obs_pos=m["obs_pos"],
                                     obs_vel=m["obs_vel"], soft_w=m["soft_w"], soft_sd=m["soft_sd"],
                                     param_exploration=0.0)
        elif m["kind"].startswith("diffdrive"):
            s = orc.diffdrive_spec(
                K=m["num_samples_K"], T=m["num_horizons_T"], dt=m["delta_t"],
                max_speed=m["max_speed"], max_omega=m["max_omega"],
                param_exploration=m["param_exploration"], param_lambda=m["param_lambda"],
                param_alpha=m["param_alpha"], sigma=m.get("sigma"),
                stage_w=(10 * np.array([5.0, 6.0, 9.0]) if m["kind"] == "diffdrive_obs" else None),
                term_w=(10 * np.array([5.0, 6.0, 9.0]) if m["kind"] == "diffdrive_obs" else None),
                obstacles=self.obstacles, margin=m.get("safety_margin_rate", 1.0))
        else:
            s = orc.racecar_spec(K=m["number_of_samples_K"], T=m["horizon_step_T"],
                                 param_alpha=m["param_alpha"], obstacles=self.obstacles,
                                 max_steer=m.get("max_steer_abs", 0.523), max_accel=m.get("max_accel_abs", 2.0),
                                 param_lambda=m.get("param_lambda", 50.0))
        if m.get("viz"):
            s.viz_optimal = s.viz_sampled = True
        for k, v in override.items():
            setattr(s, k, v)
        return s

    def tick_inputs(self, i):
        r = self.rec
        return dict(path=self.path, U=r["U0"][i], idx=int(r["idx0"][i]) if "idx0" in r else 0, x0=r["x0"][i], eps=self.eps[i])


def rel_err(a, b, floor=1e-12):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


# ---- BASELINE config 1 at its literal size: c1_K1000_T30_seed{0..4}_pe{1e-4,0.05}.npz ------------------------
C1_SEEDS, C1_PES = (0, 1, 2, 3, 4), ("1e-4", "0.05")


def c1_eps(rng, K=1000, T=30):
    """Noise of one C1 tick: eps ~ N(0, diag(0.1, 0.01)) from np.random.default_rng(seed), drawn as standard normals times
    the Cholesky factor (PCG64 + ziggurat: the stream is stable across numpy versions, unlike multivariate_normal's SVD
    signs), rounded to float32.  make_golden.py fed exactly these values to the reference class; the tests regenerate
    them from the seed, so the 48 MB of noise per case is not stored."""
    return (rng.standard_normal((K, T, 2)) * np.sqrt(np.array([0.1, 0.01]))).astype(np.float32)


class C1Golden:
    """One 200-tick closed loop of the unmodified reference class at K=1000, T=30 (x0, idx, u0, U for every tick; the
    per-sample costs S for the ticks listed in S_ticks)."""

    def __init__(self, seed, pe_tag):
        z = np.load(os.path.join(GOLDEN_DIR, "c1_K1000_T30_seed%d_pe%s.npz" % (seed, pe_tag)))
        self.meta = json.loads(str(z["meta"]))
        self.path = z["path"]
        self.z = {k: z[k] for k in z.files if k not in ("meta", "path")}
        self.n_ticks = self.z["x0"].shape[0]
        self.seed = seed

    def spec(self, **override):
        m = self.meta
        s = orc.diffdrive_spec(K=m["num_samples_K"], T=m["num_horizons_T"], dt=m["delta_t"], max_speed=m["max_speed"],
                               max_omega=m["max_omega"], param_exploration=m["param_exploration"],
                               param_lambda=m["param_lambda"], param_alpha=m["param_alpha"])
        for k, v in override.items():
            setattr(s, k, v)
        return s

    def eps_stream(self):
        """Yields the noise of tick 0, 1, 2, ... (one generator call per tick, as the fixture was made)."""
        rng = np.random.default_rng(self.seed)
        while True:
            yield c1_eps(rng, self.meta["num_samples_K"], self.meta["num_horizons_T"])
