// Batch planning on the device, for the call with device-resident labels and lengths
// (b200ctc_loss_and_grad_dev): what api.cu's host loop does for the warp-ctc style call -- per-utterance
// metadata, input validation, the feasibility test L + repeats <= T, and the launch order of the lattice
// CTAs (longest lattice first) -- as one small kernel, so that a call consists of kernel launches only and
// can be captured into a CUDA graph (reference call site: models/pytorch_v3/ctc/ctc.py:294-326, whose
// x_lens.cpu() / python label loop this replaces).
#include "common.cuh"

namespace b200ctc {

namespace {

constexpr int kPlanThreads = 1024;
constexpr int kPlanSmemKeys = 6144;   // work keys of up to this many utterances are ranked from shared memory

__device__ __forceinline__ long long work_key(const UttMeta& m) {
  return (long long)m.T * (2 * m.L + 1) * m.feasible;
}

// One CTA.  Phase 1: one warp per utterance (labels checked and repeats counted 32 at a time).
// Phase 2: rank sort of the utterances by decreasing lattice work, ties by index (stable).
__global__ void __launch_bounds__(kPlanThreads) plan_kernel(CallParams p, UttMeta* __restrict__ meta,
                                                            int* __restrict__ order, int* __restrict__ flags) {
  extern __shared__ long long s_key[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const int B = p.B;
  const int J_max = groups_of(p.max_label_len), W_max = em_width_of(p.max_label_len);
  for (int b = warp; b < B; b += n_warps) {
    int L = p.dev_label_lens[b], T = p.dev_act_lens[b];
    const bool bad_len = L < 0 || L > p.max_label_len || T < 0 || T > p.T;
    if (bad_len) { L = 0; T = 0; }
    const int* lab = p.labels + (long long)b * p.label_stride;
    int repeats = 0;
    bool bad_lab = false;
    for (int i = lane; i < L; i += 32) {
      const int s = lab[i];
      bad_lab |= s < 0 || s >= p.V || s == p.blank;
      repeats += (i > 0 && s == lab[i - 1]) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) repeats += __shfl_xor_sync(0xffffffffu, repeats, o);
    bad_lab = __any_sync(0xffffffffu, bad_lab);
    if (lane == 0) {
      UttMeta m;
      m.T = T;
      m.L = L;
      m.lab_off = b * p.label_stride;
      m.feasible = (!bad_len && !bad_lab && L + repeats <= T) ? 1 : 0;
      m.J = groups_of(L);
      m.W = em_width_of(L);
      m.scratch_off = (long long)b * (p.T + 1) * J_max;     // worst-case regions: no prefix sum over the mini-batch
      m.em_off = (long long)b * p.T * W_max;
      m.sym_off = b * p.max_label_len;
      m.pad_ = 0;
      meta[b] = m;
      flags[b] = (bad_len || bad_lab) ? FLAG_INVALID_INPUT : 0;
      if (b < kPlanSmemKeys) s_key[b] = work_key(m);
    }
  }
  if (tid == 0) flags[B] = 0;   // finished-utterance counter of the lattice kernel
  __threadfence_block();
  __syncthreads();
  for (int b = tid; b < B; b += blockDim.x) {
    const long long kb = b < kPlanSmemKeys ? s_key[b] : work_key(meta[b]);
    int rank = 0;
    const int ns = B < kPlanSmemKeys ? B : kPlanSmemKeys;
    for (int j = 0; j < ns; ++j) {
      const long long kj = s_key[j];
      rank += (kj > kb || (kj == kb && j < b)) ? 1 : 0;
    }
    for (int j = ns; j < B; ++j) {
      const long long kj = work_key(meta[j]);
      rank += (kj > kb || (kj == kb && j < b)) ? 1 : 0;
    }
    order[rank] = b;
  }
}

}  // namespace

cudaError_t launch_plan(const CallParams& p, UttMeta* meta, int* order, int* flags, cudaStream_t stream) {
  const int n_keys = p.B < kPlanSmemKeys ? p.B : kPlanSmemKeys;
  const size_t smem = (size_t)n_keys * sizeof(long long);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  plan_kernel<<<1, kPlanThreads, smem, stream>>>(p, meta, order, flags);
  return cudaGetLastError();
}

}  // namespace b200ctc
