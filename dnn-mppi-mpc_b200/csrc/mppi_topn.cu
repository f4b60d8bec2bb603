// Cost ordering for the trajectory viewers: `sorted_idx = np.argsort(S)` (controllers/mppi_differential_drive.py:153)
// and `torch.argsort(self.cost_total)[:num_top_samples]` (test/test_mppi_diff_obs.py:102-104).  Not on the control
// path -- it only runs when a caller asks for the top-N sampled trajectories -- so the sort itself is the CUDA
// toolkit's cub::DeviceRadixSort (stable: equal costs stay in sample order).
#include <cub/device/device_radix_sort.cuh>

#include "mppi_launch.h"

namespace {
__global__ void iota_kernel(int *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
}  // namespace

size_t mppi_sort_costs_temp_bytes(int K) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const float *)nullptr, (float *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, K);
    return bytes;
}

cudaError_t mppi_sort_costs(const float *d_S, int K, float *d_S_sorted, int *d_idx_sorted, int *d_iota, void *d_temp,
                            size_t temp_bytes, cudaStream_t st) {
    iota_kernel<<<(K + 255) / 256, 256, 0, st>>>(d_iota, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_S, d_S_sorted, d_iota, d_idx_sorted, K, 0, 32, st);
}
