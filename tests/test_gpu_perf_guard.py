"""Coarse device-time guards for the kernels behind BASELINE.json's configurations.  They are not benchmarks (bench.py is):
each bound sits ~1.6-2x above what a B200 measures, tight enough to catch a code-generation accident -- e.g. the 16-entry
argmin tree of the dynamic waypoint window once compiled into local-memory arrays (91 registers + a stack frame instead of 126)
and made every race-car tick 2.5x slower with all parity tests green -- and loose enough not to trip on a power-capped part:
every bound is scaled by the FP32 rate the library's own register-only FFMA probe reaches in this process (mppi_probe_fp32_peak),
so a GPU whose clocks are locked low (a fresh lease can start at 1050 of 1965 MHz) moves the bounds with it."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import Golden  # noqa: E402
from gpu_util import engine_from_spec  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402


_FP32_NOMINAL_TFLOPS = 71.0          # what the probe measures on an unthrottled B200 (bench.py: 71-73)


def _clock_scale():
    """>= 1: how much slower than a full-clock B200 this GPU runs FP32 right now."""
    import mppi_b200
    v = ctypes.c_double(0.0)
    rc = mppi_b200.load().mppi_probe_fp32_peak(0, 0, ctypes.byref(v))
    assert rc == 0 and v.value > 0.0, (rc, v.value)
    return max(1.0, _FP32_NOMINAL_TFLOPS / v.value)


def _device_ms_per_tick(eng, x0, n=30, warm=8):
    st = torch.cuda.Stream()
    eng.set_stream(st.cuda_stream)
    for i in range(warm):
        eng.step_async(x0, None, 7, i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(n):
        eng.step_async(x0, None, 7, warm + i)
    b.record(st)
    torch.cuda.synchronize()
    eng.set_stream(0)
    return a.elapsed_time(b) / n


def test_racecar_K16384_H50_tick_stays_under_its_device_time_bound():
    """configs[1] (controllers/mppi_race_car_obstacle.py:65-131 at K = 16 384, T = 50): 47 us per tick measured with the time-parallel
    rollout (74 us with the serial one, which MPPI_TPAR=0 still selects); bound 90 us."""
    g = Golden("racecar_default")
    sp = g.spec()
    sp.K, sp.T = 16384, 50
    eng = engine_from_spec(sp, g.path)
    ms = _device_ms_per_tick(eng, np.asarray(g.rec["x0"][0], np.float64))
    eng.close()
    scale = _clock_scale()
    assert ms < 0.090 * scale, (ms, scale)


def test_diffdrive_K1M_H50_tick_stays_under_its_device_time_bound():
    """The headline workload (controllers/mppi_differential_drive.py:87-165 at K = 1 048 576, T = 50, sum / frozen): 0.418 ms; bound 0.65 ms."""
    g = Golden("diffdrive_pe0.05")
    sp = orc.diffdrive_spec(K=1 << 20, T=50, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    ms = _device_ms_per_tick(eng, np.array([0.3, 0.2, 0.4]), n=12, warm=4)
    eng.close()
    scale = _clock_scale()
    assert ms < 0.65 * scale, (ms, scale)


@pytest.mark.parametrize("n_in,bound", [(3, 1.45), (5, 1.55)])
def test_learned_dynamics_K65536_H30_tick_stays_under_its_device_time_bound(n_in, bound):
    """configs[2] (dnn/simple_mlp.py:18-23 residual, K = 65 536, T = 30): 0.90 ms (3 inputs, pair MMAs) / 0.97 ms (5 inputs, layer 1 on
    the tensor core) measured."""
    g = Golden("diffdrive_pe0.05")
    mlp = orc.make_mlp(seed=0, out_scale=0.01, n_hidden=2, n_in=n_in)
    sp = orc.diffdrive_spec(K=65536, T=30, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen", model="diffdrive_mlp", mlp=mlp)
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    eng.set_mlp([mlp["W%d" % i] for i in range(4)], [mlp["b%d" % i] for i in range(4)])
    ms = _device_ms_per_tick(eng, np.array([0.4, 0.3, 0.5]), n=10, warm=3)
    eng.close()
    scale = _clock_scale()
    assert ms < bound * scale, (ms, scale)
