// Batch planning on the device, for the call with device-resident labels and lengths
// (b200ctc_loss_and_grad_dev): what api.cu's host loop does for the warp-ctc style call -- per-utterance
// metadata, input validation, the feasibility test L + repeats <= T, and the launch order of the lattice
// CTAs (longest lattice first) -- as one small kernel, so that a call consists of kernel launches only and
// can be captured into a CUDA graph (reference call site: models/pytorch_v3/ctc/ctc.py:294-326, whose
// x_lens.cpu() / python label loop this replaces).
#include "common.cuh"

namespace b200ctc {

namespace {

constexpr int kPlanThreads = 256;    // few registers and threads: the CTA fits next to a lattice CTA of the previous call
constexpr int kPlanSmemUtts = 704;   // up to this many utterances are planned in shared memory, underneath the previous call

__device__ __forceinline__ long long work_key(const UttMeta& m) {
  return (long long)m.T * (2 * m.L + 1) * m.feasible;
}

// Metadata of utterance b, computed by one warp (labels checked and repeats counted 32 at a time).
__device__ __forceinline__ UttMeta plan_utterance(const CallParams& p, int b, int lane, int J_max, int W_max, int* flag) {
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
  *flag = (bad_len || bad_lab) ? FLAG_INVALID_INPUT : 0;
  return m;
}

// One CTA, launched PROGRAMMATICALLY behind the previous call's lattice kernel (which signals its dependents at
// its start): everything that only reads the caller's labels and lengths -- the per-utterance scan and the rank
// sort by decreasing lattice work, ties by index -- happens in shared memory while that kernel still runs; the
// tables in the workspace, which the previous call may still be reading, are written after griddepcontrol.wait.
// The wait also keeps this kernel from COMPLETING before the previous call: the softmax-rows kernel that
// follows (a normal launch) is ordered behind this kernel only, and it overwrites the shared workspace.
__global__ void __launch_bounds__(kPlanThreads) plan_kernel(CallParams p, UttMeta* __restrict__ meta,
                                                            int* __restrict__ order, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const int B = p.B;
  const int J_max = groups_of(p.max_label_len), W_max = em_width_of(p.max_label_len);
  if (B <= kPlanSmemUtts) {
    UttMeta* s_meta = reinterpret_cast<UttMeta*>(s_raw);
    long long* s_key = reinterpret_cast<long long*>(s_meta + B);
    int* s_flag = reinterpret_cast<int*>(s_key + B);
    int* s_order = s_flag + B;
    // lengths: one thread per utterance
    for (int b = tid; b < B; b += blockDim.x) {
      int L = p.dev_label_lens[b], T = p.dev_act_lens[b];
      const bool bad_len = L < 0 || L > p.max_label_len || T < 0 || T > p.T;
      if (bad_len) { L = 0; T = 0; }
      s_meta[b].T = T;
      s_meta[b].L = L;
      s_meta[b].feasible = bad_len ? 0 : 1;
      s_flag[b] = bad_len ? FLAG_INVALID_INPUT : 0;
      s_order[b] = 0;                                     // repeats of the utterance, for now
    }
    __syncthreads();
    // labels: one work item per 32 consecutive labels of an utterance, so that the loads of all utterances are
    // in flight together (a warp per utterance, one utterance after the other, took 40 us for 128 utterances)
    {
      const int n_chunks = (p.max_label_len + 31) >> 5, items = B * n_chunks;
      for (int it = tid; it < items; it += blockDim.x) {
        const int b = it / n_chunks, i0 = (it - b * n_chunks) << 5, L = s_meta[b].L;
        if (i0 >= L) continue;
        const int* lab = p.labels + (long long)b * p.label_stride + i0;
        const int n = min(32, L - i0);
        int prev = i0 > 0 ? lab[-1] : -1, rep = 0;
        bool bad = false;
#pragma unroll 8
        for (int k = 0; k < n; ++k) {
          const int sym = lab[k];
          bad |= sym < 0 || sym >= p.V || sym == p.blank;
          rep += (sym == prev) ? 1 : 0;                   // prev == -1 at the first label: never equal to a valid symbol
          prev = sym;
        }
        if (rep) atomicAdd(&s_order[b], rep);
        if (bad) atomicOr(&s_flag[b], FLAG_INVALID_INPUT);
      }
    }
    __syncthreads();
    for (int b = tid; b < B; b += blockDim.x) {
      UttMeta m = s_meta[b];
      if (s_flag[b] || m.L + s_order[b] > m.T) m.feasible = 0;
      m.lab_off = b * p.label_stride;
      m.J = groups_of(m.L);
      m.W = em_width_of(m.L);
      m.scratch_off = 0;
      m.em_off = 0;
      m.sym_off = b * p.max_label_len;
      m.pad_ = 0;
      s_meta[b] = m;
      s_key[b] = work_key(m);
    }
    __syncthreads();
    // compact scratch / emission regions (exclusive prefix sums over the mini-batch, one warp): the workspace is
    // sized for the worst case, but the part that is touched stays as dense as in the host-planned call
    if (warp == 0) {
      long long base_s = 0, base_e = 0;
      for (int b0 = 0; b0 < B; b0 += 32) {
        const int b = b0 + lane;
        const long long ns = b < B ? (long long)(s_meta[b].T + 1) * s_meta[b].J : 0;
        const long long ne = b < B ? (long long)s_meta[b].T * s_meta[b].W : 0;
        long long is = ns, ie = ne;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const long long vs = __shfl_up_sync(0xffffffffu, is, o), ve = __shfl_up_sync(0xffffffffu, ie, o);
          if (lane >= o) { is += vs; ie += ve; }
        }
        if (b < B) { s_meta[b].scratch_off = base_s + is - ns; s_meta[b].em_off = base_e + ie - ne; }
        base_s += __shfl_sync(0xffffffffu, is, 31);
        base_e += __shfl_sync(0xffffffffu, ie, 31);
      }
    }
    for (int b = tid; b < B; b += blockDim.x) {
      const long long kb = s_key[b];
      int rank = 0;
#pragma unroll 4
      for (int j = 0; j < B; ++j) {
        const long long kj = s_key[j];
        rank += (kj > kb || (kj == kb && j < b)) ? 1 : 0;
      }
      s_order[rank] = b;
    }
    __syncthreads();
    pdl_wait_primary();                                   // the previous call no longer reads the tables
    for (int b = tid; b < B; b += blockDim.x) {
      meta[b] = s_meta[b];
      flags[b] = s_flag[b];
      order[b] = s_order[b];
    }
    if (tid == 0) flags[B] = 0;                           // finished-utterance counter of the lattice kernel
    return;
  }
  // large mini-batches: planned in place, after the previous call
  pdl_wait_primary();
  for (int b = warp; b < B; b += n_warps) {
    int f;
    const UttMeta m = plan_utterance(p, b, lane, J_max, W_max, &f);
    if (lane == 0) { meta[b] = m; flags[b] = f; }
  }
  if (tid == 0) flags[B] = 0;
  __threadfence_block();
  __syncthreads();
  for (int b = tid; b < B; b += blockDim.x) {
    const long long kb = work_key(meta[b]);
    int rank = 0;
    for (int j = 0; j < B; ++j) {
      const long long kj = work_key(meta[j]);
      rank += (kj > kb || (kj == kb && j < b)) ? 1 : 0;
    }
    order[rank] = b;
  }
}

}  // namespace

cudaError_t launch_plan(const CallParams& p, UttMeta* meta, int* order, int* flags, cudaStream_t stream) {
  const size_t smem = p.B <= kPlanSmemUtts ? (size_t)p.B * (sizeof(UttMeta) + sizeof(long long) + 2 * sizeof(int)) : 0;   // <= 44 KB
  // Programmatic dependent launch: the lattice kernel of the PREVIOUS call on this stream signals its
  // dependents at its start, so this kernel runs underneath it instead of after it (see plan_kernel).
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(kPlanThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, plan_kernel, p, meta, order, flags);
}

}  // namespace b200ctc
