// Per-robot reference paths generated ON THE DEVICE: the course generator of the reference,
// `calc_spline_course` (path_generator/cubic_spline_planner.py:311-323) -- an arclength-parameterised natural cubic
// spline through a robot's waypoints (CubicSpline2D :216-234 over CubicSpline1D :44-172), sampled every `ds`, heading =
// atan2 of the first derivatives (:294-309) -- one CTA per robot, so a fleet of thousands of robots gets its paths in
// one launch instead of thousands of Python spline fits.  Spec / checker: oracle/spline_oracle.py.
//
// The spline set-up (n_wp <= 32 knots: chord lengths, one tridiagonal solve per coordinate) is a few hundred FP64
// operations done by one thread per coordinate; the N = ceil(s_end / ds) samples are evaluated by the whole CTA.  All
// arithmetic is FP64 like the reference (np.linalg.solve there, the Thomas recurrence here: the natural-spline matrix
// is diagonally dominant, both are accurate to a few ulp), results are rounded once to the float4 path layout.
#include <cuda_runtime.h>
#include <math.h>

#include "mppi_launch.h"

namespace {

constexpr int NWP_MAX = MPPI_SPLINE_MAX_WAYPOINTS;

struct Spline1D {
    double a[NWP_MAX], b[NWP_MAX], c[NWP_MAX], d[NWP_MAX];
};

// natural cubic spline coefficients through (s_i, y_i): __calc_A / __calc_B (:146-172), coefficients b, d (:61-67)
__device__ void fit_spline(const double *s, const double *y, int n, Spline1D &sp) {
    double h[NWP_MAX], diag[NWP_MAX], rhs[NWP_MAX], upper[NWP_MAX];
    for (int i = 0; i < n - 1; ++i) h[i] = s[i + 1] - s[i];
    for (int i = 0; i < n; ++i) sp.a[i] = y[i];
    // rows: 0 -> c0 = 0; i = 1..n-2 -> h[i-1] c[i-1] + 2 (h[i-1] + h[i]) c[i] + h[i] c[i+1] = B[i]; n-1 -> c = 0
    diag[0] = 1.0; upper[0] = 0.0; rhs[0] = 0.0;
    for (int i = 1; i < n - 1; ++i) {
        const double lower = h[i - 1];
        const double bi = 3.0 * (sp.a[i + 1] - sp.a[i]) / h[i] - 3.0 * (sp.a[i] - sp.a[i - 1]) / h[i - 1];
        const double w = lower / diag[i - 1];
        diag[i] = 2.0 * (h[i - 1] + h[i]) - w * upper[i - 1];
        upper[i] = h[i];
        rhs[i] = bi - w * rhs[i - 1];
    }
    sp.c[n - 1] = 0.0;
    for (int i = n - 2; i >= 1; --i) sp.c[i] = (rhs[i] - upper[i] * sp.c[i + 1]) / diag[i];
    sp.c[0] = 0.0;
    for (int i = 0; i < n - 1; ++i) {
        sp.d[i] = (sp.c[i + 1] - sp.c[i]) / (3.0 * h[i]);
        sp.b[i] = 1.0 / h[i] * (sp.a[i + 1] - sp.a[i]) - h[i] / 3.0 * (2.0 * sp.c[i] + sp.c[i + 1]);
    }
}

__global__ void __launch_bounds__(128) mppi_spline_kernel(const float *__restrict__ wx, const float *__restrict__ wy, int n_wp,
                                                          double ds, int max_pts, float4 *__restrict__ paths,
                                                          int *__restrict__ path_len) {
    __shared__ double s[NWP_MAX];
    __shared__ Spline1D sx, sy;
    __shared__ int n_pts;
    const int r = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {                                   // __calc_s (:226-232): cumulative chord length
        s[0] = 0.0;
        for (int i = 0; i + 1 < n_wp; ++i) {
            const double dx = (double)wx[r * n_wp + i + 1] - (double)wx[r * n_wp + i];
            const double dy = (double)wy[r * n_wp + i + 1] - (double)wy[r * n_wp + i];
            s[i + 1] = s[i] + hypot(dx, dy);
        }
        const double cnt = ceil(s[n_wp - 1] / ds);    // len(np.arange(0, s[-1], ds))
        n_pts = cnt > 2147483647.0 ? 2147483647 : (int)cnt;
    }
    __syncthreads();
    if (tid < 2) {                                    // one thread per coordinate
        double y[NWP_MAX];
        const float *w = tid == 0 ? wx : wy;
        for (int i = 0; i < n_wp; ++i) y[i] = (double)w[r * n_wp + i];
        fit_spline(s, y, n_wp, tid == 0 ? sx : sy);
    }
    __syncthreads();
    const int N = n_pts;
    if (N > max_pts || N < 1) {                       // capacity exceeded / degenerate waypoints: reported by the host
        if (tid == 0) path_len[r] = -1;
        return;
    }
    for (int j = tid; j < N; j += blockDim.x) {
        const double t = (double)j * ds;              // np.arange: start + j * step
        int i = 0;                                    // bisect.bisect(s, t) - 1 (:140-144): last knot <= t
        while (i + 2 < n_wp && s[i + 1] <= t) ++i;
        const double dx = t - s[i];
        const double px = sx.a[i] + sx.b[i] * dx + sx.c[i] * (dx * dx) + sx.d[i] * (dx * dx * dx);
        const double py = sy.a[i] + sy.b[i] * dx + sy.c[i] * (dx * dx) + sy.d[i] * (dx * dx * dx);
        const double vx = sx.b[i] + 2.0 * sx.c[i] * dx + 3.0 * sx.d[i] * (dx * dx);
        const double vy = sy.b[i] + 2.0 * sy.c[i] * dx + 3.0 * sy.d[i] * (dx * dx);
        paths[(size_t)r * max_pts + j] = make_float4((float)px, (float)py, (float)atan2(vy, vx), 0.f);
    }
    if (tid == 0) path_len[r] = N;
}

}  // namespace

cudaError_t mppi_launch_spline(const float *d_wx, const float *d_wy, int n_robots, int n_wp, double ds, int max_pts,
                               float4 *d_paths, int *d_path_len, cudaStream_t st) {
    mppi_spline_kernel<<<n_robots, 128, 0, st>>>(d_wx, d_wy, n_wp, ds, max_pts, d_paths, d_path_len);
    return cudaGetLastError();
}
