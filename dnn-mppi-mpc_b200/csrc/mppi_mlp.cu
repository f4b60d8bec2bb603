// K3 placeholder: filled in by the tcgen05 implementation.
#include "mppi_mlp.h"
struct MlpState { int K, T; };
MlpState *mlp_create(int, int) { return nullptr; }
void mlp_destroy(MlpState *m) { delete m; }
cudaError_t mlp_set_weights(MlpState *, const float *const[4], const float *const[4], cudaStream_t) { return cudaErrorNotSupported; }
int mlp_rollout_costs(MlpState *, const TickArgs &, const float *, float *, cudaStream_t) { return -1; }
int mlp_launches_per_tick(const MlpState *) { return 0; }
