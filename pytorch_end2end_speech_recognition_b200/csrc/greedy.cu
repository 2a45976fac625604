// K4: batched best-path (greedy) CTC decoder.
//
// Replaces the python loops of models/pytorch_v3/ctc/decoders/greedy_decoder.py:32-45:
//   :32-37  per-frame np.argmax over the vocabulary for t < x_lens[b]   -> frame_argmax kernels
//   :40     collapse repeated labels (itertools.groupby)                 \  collapse kernel
//   :43-44  drop the blank label                                         /
// np.argmax semantics are kept bit-exactly: the first index wins ties, -0.0 == +0.0, and a NaN is
// "larger" than everything (the first NaN wins).
//
// Bound: HBM (one read of the logits, 4 bytes written per frame).
#include <cstdlib>

#include "common.cuh"

namespace b200ctc {

namespace {

// Order-preserving key: (value key << 32) | (~index).  max() over keys = np.argmax.
__device__ __forceinline__ unsigned long long argmax_key(float x, int idx) {
  unsigned int k;
  if (x != x) {
    k = 0xffffffffu;
  } else {
    x += 0.0f;  // -0.0 -> +0.0
    unsigned int bits = __float_as_uint(x);
    k = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
  }
  return ((unsigned long long)k << 32) | (unsigned long long)(0xffffffffu - (unsigned int)idx);
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) {
  return a > b ? a : b;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = umax64(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int key_index(unsigned long long k) {
  return (int)(0xffffffffu - (unsigned int)(k & 0xffffffffull));
}

// small vocabulary: one warp per frame
__global__ void __launch_bounds__(256) frame_argmax_warp_kernel(
    const float* __restrict__ logits, long long stride_b, long long stride_t,
    const int* __restrict__ lens, int T, int V, int B, int* __restrict__ out_tokens) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * T) return;
  const int b = (int)(row / T), t = (int)(row - (long long)b * T);
  if (t >= min(lens[b], T)) return;
  const float* x = logits + b * stride_b + t * stride_t;
  unsigned long long best = 0ull;
  for (int v = lane; v < V; v += 32) best = umax64(best, argmax_key(__ldg(x + v), v));
  best = warp_max_u64(best);
  if (lane == 0) out_tokens[row] = key_index(best);
}

// very small vocabulary (V <= 4 * LPR, LPR = 8 or 16 lanes per frame): a warp decodes 32 / LPR frames at once, so
// that one load instruction of the warp covers several short rows instead of one 120-byte row
template <int LPR>
__global__ void __launch_bounds__(256) frame_argmax_group_kernel(
    const float* __restrict__ logits, long long stride_b, long long stride_t,
    const int* __restrict__ lens, int T, int V, int B, int* __restrict__ out_tokens) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, q = lane & (LPR - 1), sub = lane / LPR;
  long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  const bool in_range = row < (long long)B * T;
  if (!in_range) row = (long long)B * T - 1;                // keep the lane in the shuffles; it stores nothing
  const int b = (int)(row / T), t = (int)(row - (long long)b * T);
  const bool live = in_range && t < min(lens[b], T);
  const float* x = logits + b * stride_b + t * stride_t;
  unsigned long long best = 0ull;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = q + LPR * i;
    if (live && v < V) best = umax64(best, argmax_key(__ldg(x + v), v));
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) best = umax64(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (live && q == 0) out_tokens[row] = key_index(best);
}

// large vocabulary: one CTA per frame, 128-bit loads on the aligned body
__global__ void __launch_bounds__(256) frame_argmax_cta_kernel(
    const float* __restrict__ logits, long long stride_b, long long stride_t,
    const int* __restrict__ lens, int T, int V, int B, int* __restrict__ out_tokens) {
  __shared__ unsigned long long red[8];
  const long long row = blockIdx.x;
  const int b = (int)(row / T), t = (int)(row - (long long)b * T);
  if (t >= min(lens[b], T)) return;
  const float* x = logits + b * stride_b + t * stride_t;
  const int tid = threadIdx.x;
  int head = (int)(((16 - ((uintptr_t)x & 15)) & 15) >> 2);
  if (head > V) head = V;
  unsigned long long best = 0ull;
  for (int v = tid; v < head; v += 256) best = umax64(best, argmax_key(__ldg(x + v), v));
  const int nvec = (V - head) >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  for (int i = tid; i < nvec; i += 256) {
    const float4 q = __ldg(x4 + i);
    const int v = head + (i << 2);
    best = umax64(best, argmax_key(q.x, v));
    best = umax64(best, argmax_key(q.y, v + 1));
    best = umax64(best, argmax_key(q.z, v + 2));
    best = umax64(best, argmax_key(q.w, v + 3));
  }
  for (int v = head + (nvec << 2) + tid; v < V; v += 256) best = umax64(best, argmax_key(__ldg(x + v), v));
  best = warp_max_u64(best);
  if ((tid & 31) == 0) red[tid >> 5] = best;
  __syncthreads();
  if (tid == 0) {
    unsigned long long r = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = umax64(r, red[i]);
    out_tokens[row] = key_index(r);
  }
}

// large vocabulary, streaming: persistent warps, one WARP per frame row (no block reduction, no barrier), eight
// 128-bit loads in flight per lane.  The one-CTA-per-frame kernel above reads C4's logits (V = 3386) at 3.8 TB/s:
// 64 000 short-lived CTAs with three loads per thread each.
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
#ifndef B200CTC_ARGMAX_IN_FLIGHT
#define B200CTC_ARGMAX_IN_FLIGHT 8
#endif
constexpr int kInFlight = B200CTC_ARGMAX_IN_FLIGHT;
__global__ void __launch_bounds__(256) frame_argmax_stream_kernel(
    const float* __restrict__ logits, long long stride_b, long long stride_t,
    const int* __restrict__ lens, int T, int V, int B, int* __restrict__ out_tokens) {
  const int lane = threadIdx.x & 31;
  const long long n_rows = (long long)B * T;
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows; row += n_warps) {
    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
    if (t >= min(lens[b], T)) continue;
    const float* x = logits + b * stride_b + t * stride_t;
    int head = (int)(((16 - ((uintptr_t)x & 15)) & 15) >> 2);
    if (head > V) head = V;
    unsigned long long best = 0ull;
    if (lane < head) best = argmax_key(__ldg(x + lane), lane);
    const int nvec = (V - head) >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x + head);
    int i = lane;
    for (; i + 32 * (kInFlight - 1) < nvec; i += 32 * kInFlight) {   // kInFlight independent loads per lane before any use
      float4 q[kInFlight];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k) q[k] = ldg_stream4(x4 + i + 32 * k);
#pragma unroll
      for (int k = 0; k < kInFlight; ++k) {
        const int v = head + ((i + 32 * k) << 2);
        best = umax64(best, argmax_key(q[k].x, v));
        best = umax64(best, argmax_key(q[k].y, v + 1));
        best = umax64(best, argmax_key(q[k].z, v + 2));
        best = umax64(best, argmax_key(q[k].w, v + 3));
      }
    }
    for (; i < nvec; i += 32) {
      const float4 q = ldg_stream4(x4 + i);
      const int v = head + (i << 2);
      best = umax64(best, argmax_key(q.x, v));
      best = umax64(best, argmax_key(q.y, v + 1));
      best = umax64(best, argmax_key(q.z, v + 2));
      best = umax64(best, argmax_key(q.w, v + 3));
    }
    const int tail = head + (nvec << 2) + lane;
    if (tail < V) best = umax64(best, argmax_key(__ldg(x + tail), tail));
    best = warp_max_u64(best);
    if (lane == 0) out_tokens[row] = key_index(best);
  }
}

// One CTA per utterance: keep frame t iff its symbol differs from frame t-1 and is not blank;
// compact in place with a block-wide prefix sum (writes never overtake reads: compaction only
// moves tokens to lower indices and each chunk is read in full before it is written).
constexpr int kCollapseThreads = 1024;
__global__ void __launch_bounds__(kCollapseThreads) collapse_kernel(
    const int* __restrict__ lens, int T, int blank, int* __restrict__ out_tokens,
    int* __restrict__ out_lens) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* row = out_tokens + (long long)b * T;
  const int n = min(max(lens[b], 0), T);   // a length outside [0, T] is clamped (the reference would raise IndexError)
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n; t0 += kCollapseThreads) {
    const int t = t0 + tid;
    const int base = base_s;  // written by the previous chunk before its closing barrier
    int sym = 0, keep = 0;
    if (t < n) {
      sym = row[t];
      const int prev = (t > 0) ? row[t - 1] : -1;  // frame t-1 is still un-compacted here (see barrier below)
      keep = (sym != blank) && (t == 0 || sym != prev);
    }
    // inclusive scan of keep
    int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();  // all reads of this chunk (and of row[t0-1]) are done
    if (warp == 0) {
      int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int up = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += up;
      }
      warp_tot[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    const int pos = base + warp_tot[warp] + incl - keep;
    // In-place safety: a token moves from frame t to pos <= t.  Frame t0-1 (the next chunk's
    // "prev") is only ever overwritten by itself (pos == t means every earlier frame was kept).
    if (keep) row[pos] = sym;
    if (tid == kCollapseThreads - 1) base_s = pos + keep;
    __syncthreads();
  }
  __syncthreads();
  const int total = (n > 0) ? base_s : 0;
  if (tid == 0) out_lens[b] = total;
  for (int t = total + tid; t < T; t += kCollapseThreads) row[t] = -1;
}

}  // namespace

cudaError_t launch_greedy(const float* logits, long long stride_b, long long stride_t, const int* lens,
                          int T, int V, int B, int blank, int* out_tokens, int* out_lens,
                          cudaStream_t stream) {
  if (B == 0) return cudaSuccess;
  const long long rows = (long long)B * T;
  if (rows > 0) {
    if (V <= 32) {
      frame_argmax_group_kernel<8><<<(unsigned)((rows + 31) / 32), 256, 0, stream>>>(logits, stride_b, stride_t, lens, T, V, B, out_tokens);
    } else if (V <= 64) {
      frame_argmax_group_kernel<16><<<(unsigned)((rows + 15) / 16), 256, 0, stream>>>(logits, stride_b, stride_t, lens, T, V, B, out_tokens);
    } else if (V <= 512) {
      const unsigned grid = (unsigned)((rows + 7) / 8);
      frame_argmax_warp_kernel<<<grid, 256, 0, stream>>>(logits, stride_b, stride_t, lens, T, V, B, out_tokens);
    } else if (std::getenv("B200CTC_GREEDY_CTA_PER_ROW")) {     // the previous kernel, for A/B measurements
      frame_argmax_cta_kernel<<<(unsigned)rows, 256, 0, stream>>>(logits, stride_b, stride_t, lens, T, V, B, out_tokens);
    } else {
      static int n_sm = 0;
      if (n_sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n_sm = 148;
      }
      long long grid = (long long)n_sm * 8;                      // eight resident CTAs of eight warps per SM
      if (grid * 8 > rows) grid = (rows + 7) / 8;
      frame_argmax_stream_kernel<<<(unsigned)grid, 256, 0, stream>>>(logits, stride_b, stride_t, lens, T, V, B, out_tokens);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  collapse_kernel<<<B, kCollapseThreads, 0, stream>>>(lens, T, blank, out_tokens, out_lens);
  return cudaGetLastError();
}

}  // namespace b200ctc
