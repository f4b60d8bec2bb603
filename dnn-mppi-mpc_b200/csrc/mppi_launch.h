// Launch entry points of mppi_kernels.cu used by the C-ABI host code (mppi_api.cu).
#pragma once
#include <cuda_runtime.h>
struct TickArgs;
cudaError_t mppi_launch_tick(const TickArgs &a, int model, int coll, int cost_kind, bool sum, bool inj, bool stash, dim3 grid, cudaStream_t st);
size_t mppi_tick_dyn_smem(int T, bool stash);
cudaError_t mppi_launch_strict(const TickArgs &a, int model, int coll, bool sum, bool inj, const unsigned *bp_n,
                               const int *bp_s, int nbp, int k_first, unsigned check_from,
                               unsigned long long *first_change, cudaStream_t st);
cudaError_t mppi_launch_merge(const TickArgs &a, const float *triples, int G, cudaStream_t st);
cudaError_t mppi_launch_traj(const TickArgs &a, int model, const float *rec, float *d_opt, float *d_samp, cudaStream_t st);
cudaError_t mppi_launch_noise(const TickArgs &a, float *d_out, int robot, cudaStream_t st);
int mppi_tick_occupancy(int model, int coll, int cost_kind, bool sum, bool inj, int window, int T, bool stash);
