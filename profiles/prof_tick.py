"""Tiny driver for ncu: a few ticks of the headline workload (diff-drive K=1M, H=50, sum/frozen, Philox).
Usage (GPU box):  python profiles/prof_tick.py [K] [T] [ticks] [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-mppi-mpc_b200")]
import numpy as np  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 5
workload = sys.argv[4] if len(sys.argv) > 4 else "diffdrive"

if workload == "diffdrive":
    from bench import diffdrive_kwargs
    from mppi_b200.mppi_differential_drive import MPPIAlgorithms
    ctrl = MPPIAlgorithms(**diffdrive_kwargs(K, T, 10.0), seed=7)
    x0 = np.zeros(3)
    for i in range(ticks):
        ctrl._calc_input_control(x0)
else:
    from mppi_b200.mppi_race_car_obstacle import MPPIRacecarController
    rc = MPPIRacecarController(horizon_step_T=T, number_of_samples_K=K, visualize_optimal_traj=False,
                               visualze_sampled_trajs=False, seed=3)
    lp = rc.generate_lemniscate_trajectory(100, 10.0).astype(np.float32)
    rc.ref_path = lp
    for i in range(ticks):
        rc._calc_control_input(lp[i])
print("ok")
