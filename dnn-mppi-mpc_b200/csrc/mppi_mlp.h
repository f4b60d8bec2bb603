// Learned-dynamics rollout (K3): unicycle + simple_mlp residual, tcgen05 tensor-core MLP.
#pragma once
#include <cuda_runtime.h>
struct TickArgs;
struct MlpState;
MlpState *mlp_create(int K, int T);
void mlp_destroy(MlpState *m);
cudaError_t mlp_set_weights(MlpState *m, const float *const W[4], const float *const b[4], cudaStream_t st);
// index update + K x T rollout through the MLP + costs -> d_S; returns 0 on success
int mlp_rollout_costs(MlpState *m, const TickArgs &a, bool sum, const float *d_eps, float *d_S, cudaStream_t st);
int mlp_launches_per_tick(const MlpState *m);
