/* TEST INFRASTRUCTURE ONLY -- plain-C (FP64) restatement of the reference MPPI tick.
 *
 * This is the *checker* for the CUDA path at sizes the Python oracle cannot reach
 * (K = 1M samples) and the "port" CPU baseline bench.py times.  It is never linked
 * into, loaded by, or called from the product library (dnn-mppi-mpc_b200/).
 *
 * Follows (reference file:line, relative to the reference tree):
 *   controllers/mppi_differential_drive.py:87-289      diff-drive tick
 *   controllers/mppi_differential_drive_obs.py:301-313 circle-circle collision
 *   controllers/mppi_race_car_obstacle.py:65-274       bicycle tick, footprint collision
 *   test/mppi_differential_drive_obs.py:202-232        goal-point cost (cost_kind 1)
 *   test/test_mppi_diff_obs.py:14-20,44-66,113-140     target pose + soft moving obstacles (cost_kind 2)
 * Semantics are those of Appendix A of SURVEY.md; the Python restatement
 * (oracle/mppi_oracle.py) is pinned bit-for-bit to the reference classes and this file
 * is pinned to the same golden vectors (tests/test_oracle_golden.py): diff-drive to
 * 1e-12 relative, race-car (whose reference accumulates in FP32) to 2e-6.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int model;          /* 0 diff-drive, 1 kinematic bicycle */
    int K, T;
    int n_exploit;      /* Q6: #{k : k < (1-param_exploration)*K}, computed by the caller */
    int window;         /* 20 / 200 */
    int cost_mode;      /* 0 = last (Q1), 1 = sum */
    int waypoint_mode;  /* 0 = strict (Q3), 1 = frozen */
    int filter_kind;    /* 0 = diff-drive convolve+tail bug, 1 = race-car edge padded */
    int yaw_wrap;       /* Q11 */
    int collision;      /* 0 none, 1 circle-circle, 2 footprint points */
    int n_obstacles;
    int n_path, path_cols;
    double dt, wheel_base;
    double u_max[2];
    double gamma;       /* lambda * (1 - alpha) */
    double temperature; /* Q2 */
    double sig_inv[4];  /* row-major inverse of Sigma */
    double chol[4];     /* row-major lower Cholesky factor of Sigma (Philox mode) */
    double stage_w[4], term_w[4];
    double margin, robot_radius, vehicle_l, vehicle_w;
    double goal[3];     /* goal point (x,y) / desired pose (x,y,yaw) */
    double ctrl_w[2], soft_w, soft_sd;
    double obs_vel[32]; /* cost_kind 2: (vx,vy) per obstacle; `obs` then holds (x,y) pairs */
    int cost_kind;      /* 0 path, 1 goal point, 2 target pose + soft moving obstacles */
} oracle_cfg_t;

#define PENALTY 1.0e10

/* ---------------------------------------------------------------- Philox4x32-10 */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void mppi_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    memcpy(out, ctr, 16);
    philox4x32_10(out, key[0], key[1]);
}

/* eps for (k, t): counter = (k, t/2, tick, robot), key = seed; outputs (0,1)->t even, (2,3)->t odd */
static inline void philox_eps(const oracle_cfg_t *c, uint64_t seed, uint32_t tick, uint32_t robot,
                              uint32_t k, int t, double e[2]) {
    uint32_t ctr[4] = {k, (uint32_t)(t >> 1), tick, robot};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    int h = (t & 1) * 2;
    double u1 = ((ctr[h] >> 9) + 0.5) * 0x1p-23, u2 = ((ctr[h + 1] >> 9) + 0.5) * 0x1p-23;
    double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586476925 * (u2 - 0.5);
    double z0 = rad * cos(ang), z1 = rad * sin(ang);
    e[0] = c->chol[0] * z0;
    e[1] = c->chol[2] * z0 + c->chol[3] * z1;
}

/* ---------------------------------------------------------------- pieces of the tick */
static inline int nearest(const oracle_cfg_t *c, const double *path, int s, double x, double y) {
    int end = s + c->window; if (end > c->n_path) end = c->n_path;
    int best = s; double bd = INFINITY;
    for (int j = s; j < end; ++j) {
        double dx = x - path[j * c->path_cols], dy = y - path[j * c->path_cols + 1];
        double d = dx * dx + dy * dy;
        if (d < bd) { bd = d; best = j; }          /* strict < : first minimum */
    }
    return best;
}

static inline double collided(const oracle_cfg_t *c, const double *obs, const double *z) {
    if (c->collision == 0) return 0.0;
    if (c->collision == 1) {
        double rr = c->robot_radius * c->margin;
        for (int m = 0; m < c->n_obstacles; ++m) {
            double dx = z[0] - obs[3 * m], dy = z[1] - obs[3 * m + 1], r = rr + obs[3 * m + 2];
            if (dx * dx + dy * dy < r * r) return 1.0;
        }
        return 0.0;
    }
    double L = c->vehicle_l * c->margin, W = c->vehicle_w * c->margin;
    const double bx[8] = {-0.5 * L, -0.5 * L, 0.0, 0.5 * L, 0.5 * L, 0.5 * L, 0.0, -0.5 * L};
    const double by[8] = {0.0, 0.5 * W, 0.5 * W, 0.5 * W, 0.0, -0.5 * W, -0.5 * W, -0.5 * W};
    double cs = cos(z[2]), sn = sin(z[2]);          /* raw yaw (Q11) */
    for (int p = 0; p < 8; ++p) {
        double px = bx[p] * cs - by[p] * sn + z[0], py = bx[p] * sn + by[p] * cs + z[1];
        for (int m = 0; m < c->n_obstacles; ++m) {
            double dx = px - obs[3 * m], dy = py - obs[3 * m + 1], r = obs[3 * m + 2];
            if (dx * dx + dy * dy < r * r) return 1.0;
        }
    }
    return 0.0;
}

/* test/mppi_differential_drive_obs.py:202-232 */
static inline double goal_cost(const oracle_cfg_t *c, const double *obs, const double *z, const double *w) {
    double dx = z[0] - c->goal[0], dy = z[1] - c->goal[1];
    double dist = sqrt(dx * dx + dy * dy);
    double a = atan2(dy, dx) - z[2];
    a = atan2(sin(a), cos(a));
    return w[0] * dist * dist + w[1] * a * a + collided(c, obs, z) * PENALTY;
}

/* test/test_mppi_diff_obs.py:44-66; obstacle m at obs[2m..] + vel * (t*dt) (:14-20) */
static inline double target_soft_cost(const oracle_cfg_t *c, const double *obs, const double *z, const double *v,
                                      int t, const double *w, int full) {
    double ex = z[0] - c->goal[0], ey = z[1] - c->goal[1], eth = z[2] - c->goal[2];
    double cost = w[0] * ex * ex + w[1] * ey * ey + w[2] * eth * eth;
    if (full) {
        cost += c->ctrl_w[0] * v[0] * v[0] + c->ctrl_w[1] * v[1] * v[1];
        double tt = t * c->dt, soft = 0.0;
        for (int m = 0; m < c->n_obstacles; ++m) {
            double dx = z[0] - (obs[2 * m] + c->obs_vel[2 * m] * tt), dy = z[1] - (obs[2 * m + 1] + c->obs_vel[2 * m + 1] * tt);
            double d = sqrt(dx * dx + dy * dy);
            if (d < c->soft_sd) soft += exp(c->soft_sd - d);
        }
        cost += c->soft_w * soft;
    }
    return cost;
}

static inline double state_cost(const oracle_cfg_t *c, const double *path, const double *obs,
                                const double *z, int j, const double *w) {
    const double *r = path + (size_t)j * c->path_cols;
    double dx = z[0] - r[0], dy = z[1] - r[1], cost;
    if (c->model == 1) {
        double yaw = z[2];
        if (c->yaw_wrap) { yaw = fmod(yaw + 2.0 * M_PI, 2.0 * M_PI); if (yaw < 0) yaw += 2.0 * M_PI; }
        double dyaw = yaw - r[2], dv = z[3] - r[3];
        cost = w[0] * dx * dx + w[1] * dy * dy + w[2] * dyaw * dyaw + w[3] * dv * dv;
    } else {
        double dyaw = z[2] - r[2];
        cost = w[0] * dx * dx + w[1] * dy * dy + w[2] * dyaw * dyaw;
    }
    return cost + collided(c, obs, z) * PENALTY;
}

static inline void dyn_step(const oracle_cfg_t *c, double *z, const double *v) {
    double cs = cos(z[2]), sn = sin(z[2]);
    if (c->model == 1) {
        double vel = z[3];
        z[0] += vel * cs * c->dt; z[1] += vel * sn * c->dt;
        z[2] += vel / c->wheel_base * tan(v[0]) * c->dt; z[3] += v[1] * c->dt;
    } else {
        z[0] += v[0] * cs * c->dt; z[1] += v[0] * sn * c->dt; z[2] += v[1] * c->dt;
    }
}

static inline double clampd(double v, double lim) { return v < -lim ? -lim : (v > lim ? lim : v); }

/* One sample: rolls out, returns its cost.  `s` is the carried waypoint index (mutated in
 * strict mode).  eps_k points at this sample's (T,2) float noise or NULL for Philox. */
static double sample_cost(const oracle_cfg_t *c, const double *path, const double *obs,
                          const double *U, const double *x0, const float *eps_k,
                          uint64_t seed, uint32_t tick, uint32_t k_global, int exploit, int *s) {
    int nx = c->model == 1 ? 4 : 3;
    double z[4] = {0, 0, 0, 0};
    for (int i = 0; i < nx; ++i) z[i] = x0[i];
    double S = 0.0;
    for (int t = 0; t < c->T; ++t) {
        double e[2];
        if (eps_k) { e[0] = eps_k[2 * t]; e[1] = eps_k[2 * t + 1]; }
        else philox_eps(c, seed, tick, 0, k_global, t, e);
        double v[2];
        v[0] = clampd(exploit ? U[2 * t] + e[0] : e[0], c->u_max[0]);
        v[1] = clampd(exploit ? U[2 * t + 1] + e[1] : e[1], c->u_max[1]);
        dyn_step(c, z, v);
        if (c->cost_kind != 0) {
            if (c->cost_mode == 1 || t == c->T - 1) {
                double q0 = U[2 * t] * c->sig_inv[0] + U[2 * t + 1] * c->sig_inv[2];
                double q1 = U[2 * t] * c->sig_inv[1] + U[2 * t + 1] * c->sig_inv[3];
                double cst = (c->cost_kind == 1 ? goal_cost(c, obs, z, c->stage_w)
                                                : target_soft_cost(c, obs, z, v, t, c->stage_w, 1)) +
                             c->gamma * (q0 * v[0] + q1 * v[1]);
                if (c->cost_mode == 1) S += cst; else S = cst;
            }
        } else if (c->cost_mode == 1 || c->waypoint_mode == 0 || t == c->T - 1) {
            int j = nearest(c, path, *s, z[0], z[1]);
            if (c->waypoint_mode == 0) *s = j;
            if (c->cost_mode == 1 || t == c->T - 1) {
                double q0 = U[2 * t] * c->sig_inv[0] + U[2 * t + 1] * c->sig_inv[2];
                double q1 = U[2 * t] * c->sig_inv[1] + U[2 * t + 1] * c->sig_inv[3];
                double cst = state_cost(c, path, obs, z, j, c->stage_w) + c->gamma * (q0 * v[0] + q1 * v[1]);
                if (c->cost_mode == 1) S += cst; else S = cst;
            }
        }
    }
    if (c->cost_kind == 1) return S + goal_cost(c, obs, z, c->term_w);
    if (c->cost_kind == 2) return S + target_soft_cost(c, obs, z, NULL, 0, c->term_w, 0);
    int j = nearest(c, path, *s, z[0], z[1]);
    if (c->waypoint_mode == 0) *s = j;
    return S + state_cost(c, path, obs, z, j, c->term_w);
}

/* Per-sample costs.  eps: (K,T,2) float32 row-major or NULL (Philox with seed/tick,
 * global sample index = k_offset + k).  Returns the index after step 1 in *idx_step1 and
 * after the whole tick in *idx_after. */
int mppi_oracle_costs(const oracle_cfg_t *c, const double *path, const double *obs,
                      const double *U, int idx, const double *x0, const float *eps,
                      uint64_t seed, uint32_t tick, uint32_t k_offset, int nthreads,
                      double *S, int *idx_step1, int *idx_after) {
    int s0 = c->cost_kind == 0 ? nearest(c, path, idx, x0[0], x0[1]) : 0;
    if (idx_step1) *idx_step1 = s0;
    size_t stride = (size_t)c->T * 2;
    if (c->waypoint_mode == 0) {                    /* strict: inherently sequential */
        int s = s0;
        for (int k = 0; k < c->K; ++k)
            S[k] = sample_cost(c, path, obs, U, x0, eps ? eps + k * stride : NULL, seed, tick,
                               k_offset + (uint32_t)k, (int)(k_offset + (uint32_t)k) < c->n_exploit, &s);
        if (idx_after) *idx_after = s;
        return 0;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(static)
    for (int k = 0; k < c->K; ++k) {
        int s = s0;
        S[k] = sample_cost(c, path, obs, U, x0, eps ? eps + k * stride : NULL, seed, tick,
                           k_offset + (uint32_t)k, (int)(k_offset + (uint32_t)k) < c->n_exploit, &s);
    }
    if (idx_after) *idx_after = s0;
    return 0;
}

/* The fixed T x T filter operators (A14), row-major M with filtered = M * x. */
void mppi_oracle_filter_matrix(int T, int kind, double *M) {
    memset(M, 0, sizeof(double) * T * T);
    if (kind == 0) {
        /* np.convolve(x, ones(10)/10, 'same'): y[n] = 0.1 * sum_{m=n-5}^{n+4} x[m]; rows 0..4
         * rescaled by 10/(i+5); the tail rescale hits row T-1 four times: 10/(i+5), i=1..4 */
        for (int n = 0; n < T; ++n)
            for (int m = n - 5; m <= n + 4; ++m)
                if (m >= 0 && m < T) M[n * T + m] = 0.1;
        for (int i = 0; i < 5 && i < T; ++i)
            for (int m = 0; m < T; ++m) M[i * T + m] *= 10.0 / (i + 5);
        for (int i = 1; i < 5; ++i)
            for (int m = 0; m < T; ++m) M[(T - 1) * T + m] *= 10.0 / (i + 5);
    } else {
        /* padded xp = [x[0:5], x, x[T-5:T]]; y[n] = 0.1 * sum_{m=n}^{n+9} xp[m] */
        for (int n = 0; n < T; ++n)
            for (int m = n; m <= n + 9; ++m) {
                int src = m < 5 ? m : (m < T + 5 ? m - 5 : m - 10);
                M[n * T + src] += 0.1;
            }
    }
}

/* A12-A15 given S.  Outputs: w_eps (T,2) raw weighted noise, U_after (T,2) shifted nominal. */
int mppi_oracle_update(const oracle_cfg_t *c, const double *U, const double *S, const float *eps,
                       uint64_t seed, uint32_t tick, uint32_t k_offset, int nthreads,
                       double *w_eps, double *U_after, double *rho_out, double *eta_out) {
    int K = c->K, T = c->T;
    double rho = INFINITY;
    for (int k = 0; k < K; ++k) if (S[k] < rho) rho = S[k];
    double eta = 0.0;
    double *acc = (double *)calloc((size_t)T * 2, sizeof(double));
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel
    {
        double *loc = (double *)calloc((size_t)T * 2, sizeof(double));
        double leta = 0.0;
        #pragma omp for schedule(static)
        for (int k = 0; k < K; ++k) {
            double w = exp(-(S[k] - rho) / c->temperature);
            leta += w;
            if (w == 0.0) continue;
            for (int t = 0; t < T; ++t) {
                double e[2];
                if (eps) { e[0] = eps[((size_t)k * T + t) * 2]; e[1] = eps[((size_t)k * T + t) * 2 + 1]; }
                else philox_eps(c, seed, tick, 0, k_offset + (uint32_t)k, t, e);
                loc[2 * t] += w * e[0]; loc[2 * t + 1] += w * e[1];
            }
        }
        #pragma omp critical
        { eta += leta; for (int i = 0; i < 2 * T; ++i) acc[i] += loc[i]; }
        free(loc);
    }
    for (int i = 0; i < 2 * T; ++i) w_eps[i] = acc[i] / eta;
    free(acc);
    double *M = (double *)malloc(sizeof(double) * T * T);
    double *Upre = (double *)malloc(sizeof(double) * T * 2);
    mppi_oracle_filter_matrix(T, c->filter_kind, M);
    for (int n = 0; n < T; ++n)
        for (int u = 0; u < 2; ++u) {
            double f = 0.0;
            for (int m = 0; m < T; ++m) f += M[n * T + m] * w_eps[2 * m + u];
            Upre[2 * n + u] = U[2 * n + u] + f;
        }
    for (int n = 0; n < T; ++n)                   /* shift, last row duplicated (A15) */
        for (int u = 0; u < 2; ++u) U_after[2 * n + u] = Upre[2 * (n + 1 < T ? n + 1 : T - 1) + u];
    free(M); free(Upre);
    if (rho_out) *rho_out = rho;
    if (eta_out) *eta_out = eta;
    return 0;
}

int mppi_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
