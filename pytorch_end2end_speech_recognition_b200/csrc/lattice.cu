// K2: alpha/beta lattice recursion + posterior occupancy update of the gradient rows.
// K3 (fused into K2's last CTA): fixed-order sum of the per-utterance costs.
#include <cstdlib>

#include "common.cuh"
#include "lattice_fast.cuh"
#include "lattice_safe.cuh"

namespace b200ctc {

namespace {

constexpr int kChunk = 4;  // frames between halo exchanges (K)

// One CTA per utterance, longest lattice first (p.order): per side NWMAX lattice warps and kReducers
// reducer warps (lattice_fast.cuh); NS lattice states per lane.  The block-exponent fast path runs unless
// the utterance was flagged by K1 or is too long for the lattice window; when the fast path gives
// up (range lost, zero probability) the same CTA redoes the utterance with the fp64 safe path.
template <int K, int NWMAX, int NS>
__global__ void __launch_bounds__(2 * (NWMAX + kReducers) * 32, 1) lattice_kernel(CallParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = p.order[blockIdx.x];
  const UttMeta m = p.meta[b];
#ifdef B200CTC_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 2048) {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_cta_time[blockIdx.x * 2] = (long long)t;
  }
#endif
  if (!m.feasible) {  // K1 already zero-filled its gradient rows
    if (threadIdx.x == 0) p.costs[b] = INFINITY;
  } else if (m.T == 0) {  // empty utterance with an empty target: probability one
    if (threadIdx.x == 0) p.costs[b] = 0.f;
  } else {
    bool use_safe = fast_warps_needed<K, NS>(m.L) > NWMAX;
    bool dirty = false;
    if (!use_safe) {
      // prologue (no K1 output needed), then wait for K1, then the sweeps -- unless K1 flagged an extreme row
      int* abort_word = nullptr;
      lattice_fast_utterance<K, NWMAX, NS>(p, b, smem, &abort_word);
      __syncthreads();
      const int why = *abort_word;
      use_safe = why != 0;
      if (use_safe && why != kAbortExtremeRow) {
        dirty = p.grads != nullptr;  // part of the gradient rows may already have been rewritten
        if (threadIdx.x == 0) atomicOr(p.flags + b, FLAG_PRECISION_LOST);
        __threadfence();             // order this thread's row updates before the rows are rebuilt
        __syncthreads();
      }
    } else {
      pdl_wait_primary();
    }
    if (use_safe) lattice_safe_utterance(p, b, smem, dirty);
  }

#ifdef B200CTC_TRACE
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x < 2048) {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_cta_time[blockIdx.x * 2 + 1] = (long long)t;
  }
#endif
  // K3, fused: the CTA that finishes last sums the per-utterance costs in a fixed order (256 strided
  // partial sums in double, then a tree), so the returned loss is bit-reproducible run to run.
  if (p.loss_sum == nullptr) return;
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                 // this CTA's cost is visible before the ticket is taken
    s_last = atomicAdd(p.done_counter, 1) == p.B - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* part = reinterpret_cast<double*>(smem);
  if (threadIdx.x < 256) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < p.B; i += 256) acc += (double)__ldcg(p.costs + i);
    part[threadIdx.x] = acc;
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *p.loss_sum = (float)part[0];
}

template <int K, int NWMAX, int NS>
cudaError_t launch_lattice_t(const CallParams& p, int max_L, cudaStream_t stream) {
  // longest label sequence the fast path's lattice window holds with NWMAX warps per side
  int l_cap = max_L;
  while (l_cap > 0 && fast_warps_needed<K, NS>(l_cap) > NWMAX) --l_cap;
  const int rw = p.gathered ? (l_cap + 1 + 3) / 4 * 4 : (p.V + 3) / 4 * 4;
  // the posterior row is not monotonic in L (its slot width steps): size for the worst L <= l_cap
  size_t smem = 0;
  for (int L = 0; L <= l_cap; ++L) {
    const size_t s = fast_smem_bytes<K, NWMAX, NS>(L, rw, p.V);
    smem = s > smem ? s : smem;
  }
  smem = smem > safe_smem_bytes(max_L) ? smem : safe_smem_bytes(max_L);
  smem = smem > 256 * sizeof(double) ? smem : 256 * sizeof(double);   // the fused cost sum
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lattice_kernel<K, NWMAX, NS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // programmatic dependent launch: this kernel may start while K1 (the previous kernel in the stream) runs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.B);
  cfg.blockDim = dim3(2 * (NWMAX + kReducers) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lattice_kernel<K, NWMAX, NS>, p);
}

}  // namespace

cudaError_t launch_lattice(const CallParams& p, int max_L, cudaStream_t stream) {
  if (p.B == 0) return cudaSuccess;
  // eight lattice states per lane; one, two or four 256-state windows per sweep
  const int nw = fast_warps_needed<kChunk, 8>(max_L);
  if (nw <= 1) return launch_lattice_t<kChunk, 1, 8>(p, max_L, stream);
  if (nw <= 2) return launch_lattice_t<kChunk, 2, 8>(p, max_L, stream);
  return launch_lattice_t<kChunk, 4, 8>(p, max_L, stream);   // longer label sequences (L > 463) take the safe lattice
}

#ifdef B200CTC_TRACE
// developer hook: copies the timeline of CTA 0 of the last lattice launch and clears it
extern "C" __attribute__((visibility("default"))) int b200ctc_debug_set_trace_cta(int cta) {
  return cudaMemcpyToSymbol(g_trace_cta, &cta, sizeof(int)) == cudaSuccess ? 0 : 2;
}
extern "C" __attribute__((visibility("default"))) int b200ctc_debug_read_cta_times(long long* host, int n) {
  if (cudaDeviceSynchronize() != cudaSuccess) return 2;
  return cudaMemcpyFromSymbol(host, g_cta_time, sizeof(long long) * 2 * n) == cudaSuccess ? 0 : 2;
}
extern "C" __attribute__((visibility("default"))) int b200ctc_debug_read_trace(long long* host, int* counts) {
  if (cudaDeviceSynchronize() != cudaSuccess) return 2;
  if (cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * 64 * kTraceCap) != cudaSuccess) return 2;
  if (cudaMemcpyFromSymbol(counts, g_trace_cnt, sizeof(int) * 64) != cudaSuccess) return 2;
  static int zeros[64];
  if (cudaMemcpyToSymbol(g_trace_cnt, zeros, sizeof(zeros)) != cudaSuccess) return 2;
  return 0;
}
#endif

}  // namespace b200ctc
