// K2: alpha/beta lattice recursion + posterior occupancy update of the gradient rows.
// K3: fixed-order sum of the per-utterance costs.
#include "common.cuh"
#include "lattice_safe.cuh"

namespace b200ctc {

namespace {

constexpr int kSafeThreads = 256;

__global__ void __launch_bounds__(kSafeThreads) lattice_safe_kernel(CallParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = p.order[blockIdx.x];
  const UttMeta m = p.meta[b];
  if (!m.feasible) {  // K1 already zero-filled its gradient rows
    if (threadIdx.x == 0) p.costs[b] = INFINITY;
    return;
  }
  if (m.T == 0) {  // empty utterance with an empty target: probability one
    if (threadIdx.x == 0) p.costs[b] = 0.f;
    return;
  }
  lattice_safe_utterance(p, b, smem, /*rows_dirty=*/false);
}

// Single CTA, fixed summation tree: the returned loss is bit-reproducible run to run.
__global__ void __launch_bounds__(256) cost_sum_kernel(const float* __restrict__ costs, int B,
                                                        float* __restrict__ loss_sum) {
  __shared__ double part[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) acc += (double)costs[i];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_sum = (float)part[0];
}

}  // namespace

cudaError_t launch_lattice(const CallParams& p, int max_L, cudaStream_t stream) {
  if (p.B == 0) return cudaSuccess;
  const size_t smem = safe_smem_bytes(max_L);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lattice_safe_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  lattice_safe_kernel<<<p.B, kSafeThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_cost_sum(const CallParams& p, cudaStream_t stream) {
  if (!p.loss_sum) return cudaSuccess;
  cost_sum_kernel<<<1, 256, 0, stream>>>(p.costs, p.B, p.loss_sum);
  return cudaGetLastError();
}

}  // namespace b200ctc
