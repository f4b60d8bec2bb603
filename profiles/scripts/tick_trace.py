"""Where a tick's time outside the rollouts goes: per-CTA %globaltimer stamps of one K = 1M, H = 50 tick
(mppi_set_trace / mppi_get_trace).  Usage (GPU box): python profiles/scripts/tick_trace.py [K] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-mppi-mpc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bench import diffdrive_kwargs  # noqa: E402
from mppi_b200.mppi_differential_drive import MPPIAlgorithms  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
c = MPPIAlgorithms(**diffdrive_kwargs(K, T, 10.0), seed=7)
eng = c.engine
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
eng.set_trace(True)
x0 = np.zeros(3)
for i in range(5):
    eng.step_async(x0, None, 7, i)
torch.cuda.synchronize()
rows = []
for i in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    eng.step_async(x0, None, 7, 10 + i)
    b.record(st)
    torch.cuda.synchronize()
    ctas, last = eng.trace()
    t0 = ctas[:, 0].min()
    dur = ctas[:, 1] - ctas[:, 0]
    rows.append((a.elapsed_time(b) * 1e3, (ctas[:, 0].max() - t0) / 1e3, dur.min() / 1e3, np.median(dur) / 1e3, dur.max() / 1e3,
                 (ctas[:, 1].max() - t0) / 1e3, (last[0] - ctas[:, 1].max()) / 1e3, (last[1] - last[0]) / 1e3, (last[1] - t0) / 1e3,
                 (np.sort(ctas[:, 1])[-1] - np.sort(ctas[:, 1])[len(ctas) // 2]) / 1e3))
r = np.median(np.array(rows), axis=0)
print("K=%d T=%d, %d CTAs, medians over 20 ticks (us):" % (K, T, len(ctas)))
for name, v in zip(("event-to-event", "CTA start skew (last start - first start)", "CTA rollout time min", "CTA rollout time median",
                    "CTA rollout time max", "first start -> last CTA rollouts done", "last rollouts done -> partials merged",
                    "merged -> nominal updated", "first start -> nominal updated", "last CTA done - median CTA done (tail)"), r):
    print("  %-48s %9.2f" % (name, v))
