"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's course generator
`calc_spline_course` (path_generator/cubic_spline_planner.py:311-323): an arclength-parameterised
natural cubic spline through the waypoints (CubicSpline2D :216-234 over CubicSpline1D :44-172),
sampled every `ds`, heading = atan2 of the first derivatives (:294-309).  Checker for
`mppi_set_ref_paths_spline` (per-robot reference paths generated on the device).

Pinned by tests/test_oracle_golden.py against tests/golden/paths.npz['spline'] and
tests/golden/spline_courses.npz, both produced by the unmodified reference function."""
import bisect
import math

import numpy as np


class _Spline1D:
    """CubicSpline1D (:44-172): natural end conditions, coefficients by np.linalg.solve."""

    def __init__(self, x, y):
        h = np.diff(x)
        n = len(x)
        self.x, self.a = list(x), [v for v in y]
        A = np.zeros((n, n))                      # __calc_A (:146-160)
        A[0, 0] = 1.0
        for i in range(n - 1):
            if i != n - 2:
                A[i + 1, i + 1] = 2.0 * (h[i] + h[i + 1])
            A[i + 1, i] = h[i]
            A[i, i + 1] = h[i]
        A[0, 1] = 0.0
        A[n - 1, n - 2] = 0.0
        A[n - 1, n - 1] = 1.0
        B = np.zeros(n)                           # __calc_B (:162-172)
        for i in range(n - 2):
            B[i + 1] = 3.0 * (self.a[i + 2] - self.a[i + 1]) / h[i + 1] - 3.0 * (self.a[i + 1] - self.a[i]) / h[i]
        self.c = np.linalg.solve(A, B)
        self.b, self.d = [], []
        for i in range(n - 1):
            self.d.append((self.c[i + 1] - self.c[i]) / (3.0 * h[i]))
            self.b.append(1.0 / h[i] * (self.a[i + 1] - self.a[i]) - h[i] / 3.0 * (2.0 * self.c[i] + self.c[i + 1]))

    def _seg(self, t):
        return bisect.bisect(self.x, t) - 1       # __search_index (:140-144)

    def pos(self, t):
        i = self._seg(t)
        dx = t - self.x[i]
        return self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2.0 + self.d[i] * dx ** 3.0

    def d1(self, t):
        i = self._seg(t)
        dx = t - self.x[i]
        return self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2.0


def spline_course(wx, wy, ds=0.1):
    """Returns the (N,3) [x, y, yaw] course of `calc_spline_course(wx, wy, ds)` (curvature omitted: the MPPI
    controllers only read x, y, yaw -- controllers/mppi_differential_drive_cuda.py:413-415)."""
    wx, wy = [float(v) for v in wx], [float(v) for v in wy]
    s = [0]
    s.extend(np.cumsum(np.hypot(np.diff(wx), np.diff(wy))))
    sx, sy = _Spline1D(s, wx), _Spline1D(s, wy)
    out = []
    for t in np.arange(0, s[-1], ds):
        out.append([sx.pos(t), sy.pos(t), math.atan2(sy.d1(t), sx.d1(t))])
    return np.array(out)
