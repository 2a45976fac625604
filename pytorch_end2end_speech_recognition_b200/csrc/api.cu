// C-ABI entry points (include/b200ctc.h): argument validation, host-side batch planning,
// workspace carving and kernel dispatch.  No exceptions leave this file.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace b200ctc {

namespace {

constexpr size_t kAlign = 256;
constexpr int kStagingSlots = 4;
constexpr int kGatherMinV = 129;  // V above this: the lattice reads gathered emission rows

inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

// The call's host-prepared tables (utterance metadata, launch order, flags, labels: ~160 KB for C3) are
// written into a pinned staging slot and copied to a device slot of the handle on the handle's own COPY
// STREAM, at call time -- not stream-ordered behind the caller's earlier kernels.  The kernels wait for the
// copy's event.  A copy enqueued on the compute stream is issued to the copy engines only when that stream
// reaches it, i.e. after the application has already queued its prefetch of the next mini-batch of logits
// (bench.py's e2e leg, any data loader), and then waits behind 12 MB of it: +0.2 ms per call measured on
// B200 (tools/e2e_timeline.py).  Issued at call time it travels while the previous call's kernels run.
struct WorkspaceLayout {
  size_t blob_bytes;   // meta + order + flags + labels
  size_t off_meta, off_order, off_flags, off_labels;
  size_t off_lse, off_em, off_scratch;
  size_t total;
};

struct BatchTotals {
  long long sum_labels = 0;
  long long em_floats = 0;       // sum_b T_b * W_b
  long long scratch_units = 0;   // sum_b T_b * J_b
};

// scratch units (32 bytes) per frame: the safe lattice stores 4 doubles per group of four states,
// the fast lattice a little more than that for very short label sequences
inline int groups_of(int L) {
  const int j4 = (2 * L + 1 + 3) / 4, j8 = (2 * L + 1 + 7) / 8;
  // fast lattice: NS*4 bytes of mantissas per group of NS states + an int32 exponent row padded to 4 groups
  const int fast8 = (36 * j8 + 12 + 31) / 32, fast4 = (20 * j4 + 12 + 31) / 32;
  int u = j4;                       // safe lattice: 4 doubles per group of four states
  u = u > fast8 ? u : fast8;
  u = u > fast4 ? u : fast4;
  return u;
}
inline int em_width_of(int L) { return (L + 1 + 3) / 4 * 4; }

int totals_from_lens(const int* label_lens, const int* act_lens, int T, int B, BatchTotals* out) {
  BatchTotals t;
  for (int b = 0; b < B; ++b) {
    const int L = label_lens[b], Tb = act_lens[b];
    if (L < 0 || Tb < 0 || Tb > T) return B200CTC_STATUS_INVALID_VALUE;
    t.sum_labels += L;
    t.em_floats += (long long)Tb * em_width_of(L);
    t.scratch_units += (long long)(Tb + 1) * groups_of(L);   // + one dump frame block
  }
  *out = t;
  return B200CTC_STATUS_SUCCESS;
}

WorkspaceLayout make_layout(const BatchTotals& t, int T, int B) {
  WorkspaceLayout w;
  size_t off = 0;
  w.off_meta = off;   off += align_up((size_t)B * sizeof(UttMeta));
  w.off_order = off;  off += align_up((size_t)B * sizeof(int));
  w.off_flags = off;  off += align_up((size_t)(B + 1) * sizeof(int));   // + the finished-utterance counter
  w.off_labels = off; off += align_up((size_t)t.sum_labels * sizeof(int));
  w.blob_bytes = off;
  w.off_lse = off;    off += align_up((size_t)T * B * sizeof(float));
  w.off_em = off;     off += align_up((size_t)t.em_floats * sizeof(float));
  w.off_scratch = off; off += align_up((size_t)t.scratch_units * kGroupBytes);
  w.total = off;
  return w;
}

}  // namespace

}  // namespace b200ctc

using namespace b200ctc;

struct b200ctc_handle {
  int device;
  struct Slot {
    void* host = nullptr;           // pinned staging buffer the host fills
    void* dev = nullptr;            // its device copy, read by the kernels
    size_t capacity = 0;
    cudaEvent_t copied = nullptr;   // the tables have reached `dev` (copy stream)
    cudaEvent_t done = nullptr;     // the call's kernels have finished: host and dev are reusable
    bool in_flight = false;
  } slots[kStagingSlots];
  cudaStream_t copy_stream = nullptr;
  int next_slot = 0;
  bool profiling = false;
  cudaEvent_t prof[4] = {nullptr, nullptr, nullptr, nullptr};  // before softmax / lattice / cost sum, after
  bool prof_valid = false;
  const int* last_flags = nullptr;  // device pointer into the last call's workspace
  int last_B = 0;
};

extern "C" {

int b200ctc_version(void) { return B200CTC_VERSION; }

const char* b200ctc_status_string(int status) {
  switch (status) {
    case B200CTC_STATUS_SUCCESS: return "success";
    case B200CTC_STATUS_INVALID_VALUE: return "invalid value";
    case B200CTC_STATUS_EXECUTION_FAILED: return "CUDA execution failed";
    case B200CTC_STATUS_UNSUPPORTED: return "unsupported problem size";
    case B200CTC_STATUS_WORKSPACE_TOO_SMALL: return "workspace too small";
    default: return "unknown status";
  }
}

int b200ctc_create(b200ctc_handle** handle, int device) {
  if (!handle) return B200CTC_STATUS_INVALID_VALUE;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return B200CTC_STATUS_EXECUTION_FAILED;  // no CPU fallback: a CUDA device is required
  }
  b200ctc_handle* h = new (std::nothrow) b200ctc_handle();
  if (!h) return B200CTC_STATUS_EXECUTION_FAILED;
  h->device = device;
  *handle = h;
  return B200CTC_STATUS_SUCCESS;
}

int b200ctc_destroy(b200ctc_handle* h) {
  if (!h) return B200CTC_STATUS_SUCCESS;
  for (auto& s : h->slots) {
    if (s.done) {
      if (s.in_flight) cudaEventSynchronize(s.done);
      cudaEventDestroy(s.done);
    }
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.host) cudaFreeHost(s.host);
    if (s.dev) cudaFree(s.dev);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (auto& e : h->prof)
    if (e) cudaEventDestroy(e);
  delete h;
  return B200CTC_STATUS_SUCCESS;
}

int b200ctc_get_workspace_size(const int* label_lens, const int* act_lens, int T, int V, int B,
                               size_t* bytes) {
  if (!bytes || T < 0 || V < 1 || B < 0 || (B > 0 && (!label_lens || !act_lens)))
    return B200CTC_STATUS_INVALID_VALUE;
  BatchTotals t;
  int st = totals_from_lens(label_lens, act_lens, T, B, &t);
  if (st != B200CTC_STATUS_SUCCESS) return st;
  *bytes = make_layout(t, T, B).total + kAlign;
  return B200CTC_STATUS_SUCCESS;
}

int b200ctc_loss_and_grad(b200ctc_handle* h, const float* acts, int64_t acts_stride_t,
                          int64_t acts_stride_b, float* grads, const int* flat_labels,
                          const int* label_lens, const int* act_lens, int T, int V, int B, int blank,
                          float* costs, float* loss_sum, void* workspace, size_t workspace_bytes,
                          void* stream_v) {
  if (!h || T < 0 || V < 1 || B < 0 || blank < 0 || blank >= V) return B200CTC_STATUS_INVALID_VALUE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (B == 0) {
    if (loss_sum && cudaMemsetAsync(loss_sum, 0, sizeof(float), stream) != cudaSuccess)
      return B200CTC_STATUS_EXECUTION_FAILED;
    return B200CTC_STATUS_SUCCESS;
  }
  if (!label_lens || !act_lens || !costs || !workspace || (!acts && T > 0))
    return B200CTC_STATUS_INVALID_VALUE;

  // ---- plan on the host ----------------------------------------------------------------------
  BatchTotals tot;
  int st = totals_from_lens(label_lens, act_lens, T, B, &tot);
  if (st != B200CTC_STATUS_SUCCESS) return st;
  if (tot.sum_labels > 0 && !flat_labels) return B200CTC_STATUS_INVALID_VALUE;
  if (tot.sum_labels > 0x7fffffffLL) return B200CTC_STATUS_UNSUPPORTED;
  if ((long long)B * V * 4 > 0x7fffffffLL) return B200CTC_STATUS_UNSUPPORTED;   // frame stride of the gradient rows in bytes (int32 in the lattice)
  const WorkspaceLayout lay = make_layout(tot, T, B);
  unsigned char* ws = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(workspace) + kAlign - 1) / kAlign * kAlign);
  const size_t lost = (size_t)(ws - reinterpret_cast<unsigned char*>(workspace));
  if (workspace_bytes < lay.total + lost) return B200CTC_STATUS_WORKSPACE_TOO_SMALL;

  // staging slot (pinned): reuse only after the copy that last read it has completed
  b200ctc_handle::Slot& slot = h->slots[h->next_slot];
  h->next_slot = (h->next_slot + 1) % kStagingSlots;
  if (slot.in_flight) {
    if (cudaEventSynchronize(slot.done) != cudaSuccess) return B200CTC_STATUS_EXECUTION_FAILED;
    slot.in_flight = false;
  }
  if (slot.capacity < lay.blob_bytes) {                 // grows on demand (rare: synchronous allocation)
    if (slot.host) cudaFreeHost(slot.host);
    if (slot.dev) cudaFree(slot.dev);
    slot.host = slot.dev = nullptr;
    slot.capacity = 0;
    const size_t cap = std::max(lay.blob_bytes * 2, (size_t)1 << 16);
    if (cudaHostAlloc(&slot.host, cap, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&slot.dev, cap) != cudaSuccess) {
      cudaGetLastError();
      return B200CTC_STATUS_EXECUTION_FAILED;
    }
    slot.capacity = cap;
  }
  if ((!slot.done && cudaEventCreateWithFlags(&slot.done, cudaEventDisableTiming) != cudaSuccess) ||
      (!slot.copied && cudaEventCreateWithFlags(&slot.copied, cudaEventDisableTiming) != cudaSuccess) ||
      (!h->copy_stream && cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess))
    return B200CTC_STATUS_EXECUTION_FAILED;

  unsigned char* blob = reinterpret_cast<unsigned char*>(slot.host);
  UttMeta* meta = reinterpret_cast<UttMeta*>(blob + lay.off_meta);
  int* order = reinterpret_cast<int*>(blob + lay.off_order);
  int* flags = reinterpret_cast<int*>(blob + lay.off_flags);
  int* labels = reinterpret_cast<int*>(blob + lay.off_labels);

  long long lab_off = 0, em_off = 0, scratch_off = 0;
  int max_L = 0;
  for (int b = 0; b < B; ++b) {
    const int L = label_lens[b], Tb = act_lens[b];
    const int* lab = flat_labels + lab_off;
    int repeats = 0;
    for (int i = 0; i < L; ++i) {
      const int s = lab[i];
      if (s < 0 || s >= V || s == blank) return B200CTC_STATUS_INVALID_VALUE;
      if (i > 0 && s == lab[i - 1]) ++repeats;
    }
    UttMeta& m = meta[b];
    m.T = Tb;
    m.L = L;
    m.lab_off = (int)lab_off;
    m.feasible = (L + repeats <= Tb) ? 1 : 0;
    m.J = groups_of(L);
    m.W = em_width_of(L);
    m.scratch_off = scratch_off;
    m.em_off = em_off;
    lab_off += L;
    em_off += (long long)Tb * m.W;
    scratch_off += (long long)(Tb + 1) * m.J;
    if (m.feasible) max_L = std::max(max_L, L);
    order[b] = b;
    flags[b] = 0;
  }
  flags[B] = 0;   // finished-utterance counter (the last CTA of the lattice kernel sums the costs)
  if (tot.sum_labels > 0) std::memcpy(labels, flat_labels, (size_t)tot.sum_labels * sizeof(int));
  // longest lattice first: CTAs are dispatched in index order, so the tail of the launch is short
  std::stable_sort(order, order + B, [&](int x, int y) {
    const long long wx = (long long)meta[x].T * (2 * meta[x].L + 1) * meta[x].feasible;
    const long long wy = (long long)meta[y].T * (2 * meta[y].L + 1) * meta[y].feasible;
    return wx > wy;
  });

  // tables: host slot -> device slot on the handle's copy stream, now; the caller's stream waits for the event
  unsigned char* tab = reinterpret_cast<unsigned char*>(slot.dev);
  if (cudaMemcpyAsync(tab, blob, lay.blob_bytes, cudaMemcpyHostToDevice, h->copy_stream) != cudaSuccess ||
      cudaEventRecord(slot.copied, h->copy_stream) != cudaSuccess ||
      cudaStreamWaitEvent(stream, slot.copied, 0) != cudaSuccess) {
    cudaGetLastError();
    return B200CTC_STATUS_EXECUTION_FAILED;
  }

  CallParams p;
  p.acts = acts;
  p.as_t = acts_stride_t;
  p.as_b = acts_stride_b;
  p.grads = grads;
  p.T = T; p.B = B; p.V = V; p.blank = blank;
  p.meta = reinterpret_cast<const UttMeta*>(tab + lay.off_meta);
  p.order = reinterpret_cast<const int*>(tab + lay.off_order);
  p.flags = reinterpret_cast<int*>(tab + lay.off_flags);
  p.done_counter = p.flags + B;
  p.labels = reinterpret_cast<const int*>(tab + lay.off_labels);
  p.lse = reinterpret_cast<float*>(ws + lay.off_lse);
  p.em = reinterpret_cast<float*>(ws + lay.off_em);
  p.scratch = ws + lay.off_scratch;
  p.costs = costs;
  p.loss_sum = loss_sum;
  p.gathered = (V >= kGatherMinV || grads == nullptr) ? 1 : 0;

  h->last_flags = p.flags;
  h->last_B = B;
  const bool prof = h->profiling;
  if (prof) cudaEventRecord(h->prof[0], stream);
  cudaError_t e = launch_softmax_rows(p, stream);
  if (prof) cudaEventRecord(h->prof[1], stream);
  if (e == cudaSuccess) e = launch_lattice(p, max_L, stream);   // its last CTA also writes loss_sum (fixed-order sum)
  // the slot (host and device side) is reusable once the kernels of this call have finished
  if (cudaEventRecord(slot.done, stream) == cudaSuccess) slot.in_flight = true;
  else if (e == cudaSuccess) e = cudaErrorUnknown;
  if (prof) cudaEventRecord(h->prof[2], stream);
  if (prof) {
    cudaEventRecord(h->prof[3], stream);
    h->prof_valid = true;
  }
  if (e != cudaSuccess) {
    std::fprintf(stderr, "b200ctc: CUDA error: %s\n", cudaGetErrorString(e));
    return (e == cudaErrorInvalidValue || e == cudaErrorInvalidConfiguration)
               ? B200CTC_STATUS_UNSUPPORTED
               : B200CTC_STATUS_EXECUTION_FAILED;
  }
  return B200CTC_STATUS_SUCCESS;
}

int b200ctc_set_profiling(b200ctc_handle* h, int enable) {
  if (!h) return B200CTC_STATUS_INVALID_VALUE;
  if (enable) {
    for (auto& e : h->prof)
      if (!e && cudaEventCreate(&e) != cudaSuccess) return B200CTC_STATUS_EXECUTION_FAILED;
  }
  h->profiling = enable != 0;
  h->prof_valid = false;
  return B200CTC_STATUS_SUCCESS;
}

int b200ctc_get_last_kernel_ms(b200ctc_handle* h, float* ms3) {
  if (!h || !ms3 || !h->prof_valid) return B200CTC_STATUS_INVALID_VALUE;
  if (cudaEventSynchronize(h->prof[3]) != cudaSuccess) return B200CTC_STATUS_EXECUTION_FAILED;
  for (int i = 0; i < 3; ++i)
    if (cudaEventElapsedTime(ms3 + i, h->prof[i], h->prof[i + 1]) != cudaSuccess)
      return B200CTC_STATUS_EXECUTION_FAILED;
  return B200CTC_STATUS_SUCCESS;
}

int b200ctc_get_last_fallbacks(b200ctc_handle* h, int* counts2, void* stream_v) {
  if (!h || !counts2 || !h->last_flags) return B200CTC_STATUS_INVALID_VALUE;
  std::vector<int> host((size_t)h->last_B);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (cudaMemcpyAsync(host.data(), h->last_flags, host.size() * sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess) {
    cudaGetLastError();
    return B200CTC_STATUS_EXECUTION_FAILED;
  }
  counts2[0] = counts2[1] = 0;
  for (int f : host) {
    if (f & FLAG_EXTREME_ROW) ++counts2[0];
    if (f & FLAG_PRECISION_LOST) ++counts2[1];
  }
  return B200CTC_STATUS_SUCCESS;
}

int b200ctc_greedy_decode(const float* logits, int64_t stride_b, int64_t stride_t, const int* lens,
                          int T, int V, int B, int blank, int* out_tokens, int* out_lens,
                          void* stream_v) {
  if (T < 0 || V < 1 || B < 0) return B200CTC_STATUS_INVALID_VALUE;
  if (B == 0) return B200CTC_STATUS_SUCCESS;
  if (!lens || !out_lens || (T > 0 && (!logits || !out_tokens))) return B200CTC_STATUS_INVALID_VALUE;
  cudaError_t e = launch_greedy(logits, stride_b, stride_t, lens, T, V, B, blank, out_tokens, out_lens,
                                reinterpret_cast<cudaStream_t>(stream_v));
  if (e != cudaSuccess) {
    std::fprintf(stderr, "b200ctc: CUDA error: %s\n", cudaGetErrorString(e));
    return B200CTC_STATUS_EXECUTION_FAILED;
  }
  return B200CTC_STATUS_SUCCESS;
}

}  // extern "C"
