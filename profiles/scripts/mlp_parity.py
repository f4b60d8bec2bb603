"""Parity of the learned-dynamics kernel (K3) decomposed (run on the GPU box):
  device costs  vs  FP64 oracle                         -- operand rounding + hardware tanh
  device costs  vs  device-faithful oracle, exact tanh  -- hardware tanh only
  device costs  vs  device-faithful oracle, HW tanh     -- everything else (FP32 accumulation order)
for synthetic weights with a large output layer and for the reference's trained checkpoints.
Usage: python profiles/scripts/mlp_parity.py [K] [T]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-mppi-mpc_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from golden_util import GOLDEN_DIR, Golden  # noqa: E402
from gpu_util import engine_from_spec  # noqa: E402
from mppi_b200 import _lib  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402

OP = os.environ.get("MLP_OP", "f16")


def hw_tanh(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty_like(a)
    rc = _lib.load().mppi_probe_tanh(0, a.ctypes.data_as(_lib._PF), out.ctypes.data_as(_lib._PF), a.size)
    assert rc == 0
    return out


def trained(tag):
    z = np.load(os.path.join(GOLDEN_DIR, "trained_%s.npz" % tag))
    return {k: z[k].astype(np.float64) for k in z.files if k not in ("meta", "X", "Y_ref")}


def stats(a, b):
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-9)
    return "median %.2e  p99 %.2e  max %.2e" % (np.median(rel), np.quantile(rel, 0.99), rel.max())


def run(name, mlp, K, T, hw=True):
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen", model="diffdrive_mlp", mlp=mlp)
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    n = len([k for k in mlp if k[0] == "W" and k[1:].isdigit()])
    sc = [mlp[k] for k in ("in_mean", "in_scale", "out_mean", "out_scale")] if "in_scale" in mlp else []
    eng.set_mlp([mlp["W%d" % i] for i in range(n)], [mlp["b%d" % i] for i in range(n)], *sc)
    eps = torch.zeros(K, T, 2, dtype=torch.float32, device="cuda")
    eng.generate_noise(eps, seed=3, tick=1)
    S = torch.zeros(K, dtype=torch.float32, device="cuda")
    x0 = np.array([0.4, 0.3, 0.5])
    U = np.random.default_rng(2).normal(0, 0.5, (T, 2)).astype(np.float32)
    eng.set_nominal(U)
    eng.rollout_costs(x0, S, None, seed=3, tick=1)
    Sg = S.cpu().numpy().astype(np.float64)
    e64 = eps.cpu().numpy().astype(np.float64)
    t0 = time.time()
    S64, _, _ = orc.costs_vec(sp, g.path, U.astype(np.float64), 0, x0, e64)
    print("%-34s K=%d T=%d  |S| median %.3g  (oracle %.1fs)" % (name, K, T, np.median(np.abs(S64)), time.time() - t0))
    print("    vs FP64 oracle               ", stats(Sg, S64))
    sp.mlp_precision = OP
    Sf, _, _ = orc.costs_vec(sp, g.path, U.astype(np.float64), 0, x0, e64)
    print("    vs faithful(%s), exact tanh  " % OP, stats(Sg, Sf), "   [faithful vs FP64:", stats(Sf, S64), "]")
    if hw:
        sp.mlp_tanh = hw_tanh
        Sh, _, _ = orc.costs_vec(sp, g.path, U.astype(np.float64), 0, x0, e64)
        print("    vs faithful(%s), HW tanh     " % OP, stats(Sg, Sh))
    eng.close()


if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    x = np.linspace(-6, 6, 2000001).astype(np.float32)
    y = hw_tanh(x).astype(np.float64)
    ex = np.tanh(x.astype(np.float64))
    rel = np.abs(y - ex) / np.maximum(np.abs(ex), 1e-30)
    print("MUFU.TANH vs exact on [-6,6]: max rel %.3e (2^-11 = %.3e), rms rel %.3e, max abs %.3e" % (rel.max(), 2.0 ** -11, np.sqrt((rel ** 2).mean()), np.abs(y - ex).max()))
    run("synthetic 3-in out_scale 0.5", orc.make_mlp(seed=0, out_scale=0.5), K, T)
    run("synthetic 3-in out_scale 0.01", orc.make_mlp(seed=0, out_scale=0.01), K, T)
    m5 = orc.make_mlp(seed=0, out_scale=0.05, n_in=5, scalers=True, scaler_gain=1.0)
    run("synthetic 5-in scalers", m5, K, T)
    run("synthetic 3 hidden out_scale 0.5", orc.make_mlp(seed=0, out_scale=0.5, n_hidden=3), K, T)
    run("TRAINED mlp_diff_300x100", trained("mlp_diff_300x100"), K, T)
    run("TRAINED mlp_diff_300x100_3l_mppi", trained("mlp_diff_300x100_3l_mppi"), K, T)
