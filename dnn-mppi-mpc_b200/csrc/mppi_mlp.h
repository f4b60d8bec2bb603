// Learned-dynamics rollout (K3): unicycle + simple_mlp residual, tcgen05 tensor-core MLP.
#pragma once
#include <cuda_runtime.h>
struct TickArgs;
struct MlpState;
MlpState *mlp_create(int K, int T);
void mlp_destroy(MlpState *m);
// n_in = 3 (state) or 5 (state + control); n_hidden = 2 or 3 tanh layers (W / b: n_hidden + 2 layers); scaler pointers
// may be null (identity)
cudaError_t mlp_set_weights(MlpState *m, int n_in, int n_hidden, const float *const *W, const float *const *b, const double *in_mean,
                            const double *in_scale, const double *out_mean, const double *out_scale, cudaStream_t st);
// index update + K x T rollout through the MLP + costs -> d_S; returns 0 on success
int mlp_rollout_costs(MlpState *m, const TickArgs &a, bool sum, const float *d_eps, float *d_S, cudaStream_t st);
int mlp_launches_per_tick(const MlpState *m);
