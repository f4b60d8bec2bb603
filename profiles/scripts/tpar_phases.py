"""Where a small-K race-car tick goes: per-CTA %globaltimer stamps (mppi_set_trace) and, with a -DMPPI_TPAR_PROBE=1 build
(profiles/scripts/build_variant.sh tparprobe "-DMPPI_TPAR_PROBE=1"; MPPI_B200_LIB=...), the cycles of the four phases of the
time-parallel rollout printed by two CTAs at tick 25.  Usage (GPU box): python profiles/scripts/tpar_phases.py [K] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-mppi-mpc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mppi_b200.mppi_race_car_obstacle import MPPIRacecarController  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rc = MPPIRacecarController(horizon_step_T=T, number_of_samples_K=K, visualize_optimal_traj=False, visualze_sampled_trajs=False, seed=3)
lp = rc.generate_lemniscate_trajectory(100, 10.0).astype(np.float32)
rc.ref_path = lp
eng = rc.engine
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
eng.set_trace(True)
x0 = lp[0].astype(np.float64)
for i in range(20):
    eng.step_async(x0, None, 3, i)
torch.cuda.synchronize()
rows = []
for i in range(20, 40):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    eng.step_async(x0, None, 3, i)
    b.record(st)
    torch.cuda.synchronize()
    ctas, last = eng.trace()
    t0 = ctas[:, 0].min()
    dur = ctas[:, 1] - ctas[:, 0]
    rows.append((a.elapsed_time(b) * 1e3, (ctas[:, 0].max() - t0) / 1e3, np.median(dur) / 1e3, dur.max() / 1e3, (ctas[:, 1].max() - t0) / 1e3,
                 (last[0] - ctas[:, 1].max()) / 1e3, (last[1] - last[0]) / 1e3, (last[1] - t0) / 1e3))
r = np.median(np.array(rows), axis=0)
print("race-car K=%d T=%d, %d CTAs, medians over 20 ticks (us):" % (K, T, len(ctas)))
for name, v in zip(("event-to-event (single tick)", "CTA start skew", "CTA prologue + rollout + K2, median", "... max",
                    "first start -> last CTA done", "last CTA done -> partials merged", "merged -> nominal updated",
                    "first start -> nominal updated"), r):
    print("  %-44s %8.2f" % (name, v))
