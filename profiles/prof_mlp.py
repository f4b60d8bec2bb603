"""Tiny driver for ncu: a few ticks of config 3 (diff-drive + simple_mlp residual, K=65536, H=30)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-mppi-mpc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
from bench import diffdrive_kwargs  # noqa: E402
from mppi_b200.mppi_differential_drive import MPPIAlgorithms  # noqa: E402


def make_mlp(seed=0, out_scale=0.01):
    rng = np.random.default_rng(seed)
    out = {}
    for i, (o, n) in enumerate([(512, 3), (512, 512), (512, 512), (3, 512)]):
        b = 1.0 / np.sqrt(n)
        s = out_scale if i == 3 else 1.0
        out["W%d" % i] = (rng.uniform(-b, b, (o, n)) * s).astype(np.float32)
        out["b%d" % i] = (rng.uniform(-b, b, (o,)) * s).astype(np.float32)
    return out


K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 3
variant = sys.argv[4] if len(sys.argv) > 4 else "2l"
mlp = make_mlp()
if variant == "3l":          # the trained checkpoints' shape: 5 inputs, three hidden layers, scalers
    rng = np.random.default_rng(1)
    mlp = {"W0": (rng.uniform(-1, 1, (512, 5)) / np.sqrt(5)).astype(np.float32), "b0": mlp["b0"], "W1": mlp["W1"], "b1": mlp["b1"],
           "W2": mlp["W2"], "b2": mlp["b2"], "W3": (rng.uniform(-1, 1, (512, 512)) / np.sqrt(512)).astype(np.float32),
           "b3": (rng.uniform(-1, 1, (512,)) / np.sqrt(512)).astype(np.float32), "W4": mlp["W3"], "b4": mlp["b3"],
           "in_mean": [4.39, -0.126, -0.08, 0.359, -0.031], "in_scale": [5.587, 3.641, 1.06, 1.024, 1.836],
           "out_mean": [-0.561, 0.029, -0.015], "out_scale": [5.701, 3.59, 0.996]}
if variant == "5in":         # five inputs (state + control) + scalers, two hidden layers: layer 1 on the tcgen05 tensor core
    rng = np.random.default_rng(1)
    mlp = dict(mlp, W0=(rng.uniform(-1, 1, (512, 5)) / np.sqrt(5)).astype(np.float32),
               in_mean=[4.39, -0.126, -0.08, 0.359, -0.031], in_scale=[5.587, 3.641, 1.06, 1.024, 1.836],
               out_mean=[-0.561, 0.029, -0.015], out_scale=[5.701, 3.59, 0.996])
ctrl = MPPIAlgorithms(**diffdrive_kwargs(K, T, 2.0), seed=7, dynamics=mlp)
for i in range(ticks):
    ctrl._calc_input_control(np.array([0.4, 0.3, 0.5]))
print("ok")
