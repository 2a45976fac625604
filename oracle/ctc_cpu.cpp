// CPU restatement of the reference's CTC path (C++17 + OpenMP) -- TEST / BASELINE INFRASTRUCTURE.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product never does.
//
// Parity status: pinned to the reference's in-tree Chainer CTC through the golden vectors of
// tests/golden/ctc_reference_golden.npz (see oracle/ctc_ref.py); warp-ctc itself cannot be run:
// the reference's PyTorch-path CTC arithmetic is the
// un-vendored, un-pinned warp-ctc (tools/install_warpctc_pytorch.sh:7), whose source is absent from
// /root/reference, so this file restates the published algorithm with the structure of warp-ctc's
// CPU path as far as it is known [recollection]: per (t,b) column a max-subtracted softmax over the
// vocabulary; per utterance a log-space alpha sweep, then a beta sweep fused with the per-symbol
// occupancy and the gradient  softmax - exp(occ - ll) ; `#pragma omp parallel for` over the
// mini-batch.  The in-tree description of the same recursion is
// models/chainer/ctc/ctc_loss_from_chainer.py:185-199 (transition rule), :234-264 (sweeps),
// :283 (cost), :288-304 (gradient and input-length mask).  Calling convention:
// models/pytorch_v3/ctc/ctc.py:35-45 (acts [T,B,V], flat labels, label_lens, act_lens, costs[B]).
//
// Two instantiations: float (what warp-ctc's ProbT=float computes; the timed CPU baseline) and
// double (a second fp64 checker next to the numpy oracle).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

template <typename R>
inline R log_add(R a, R b) {
  const R ninf = -std::numeric_limits<R>::infinity();
  if (a == ninf) return b;
  if (b == ninf) return a;
  return a > b ? a + std::log1p(std::exp(b - a)) : b + std::log1p(std::exp(a - b));
}

// One utterance.  acts/grads point at (t=0, b) with a stride of `stride_t` elements between frames.
template <typename R>
R cost_and_grad_one(const float* acts, float* grads, int64_t stride_t, const int* lab, int L, int T,
                    int V, int blank) {
  const R ninf = -std::numeric_limits<R>::infinity();
  const R pinf = std::numeric_limits<R>::infinity();
  int repeats = 0;
  for (int i = 1; i < L; ++i) repeats += lab[i] == lab[i - 1];
  if (grads)
    for (int t = 0; t < T; ++t) std::fill(grads + t * stride_t, grads + t * stride_t + V, 0.f);
  if (T == 0) return L == 0 ? R(0) : pinf;
  if (L + repeats > T) return pinf;  // no valid alignment: cost +inf, zero gradient (DESIGN.md policy)

  const int S = 2 * L + 1;
  // softmax per frame (max-subtracted), kept as log-probabilities and probabilities
  std::vector<R> lp((size_t)T * V), alpha((size_t)T * S, ninf), beta(S), beta_next(S), occ(V);
  for (int t = 0; t < T; ++t) {
    const float* x = acts + t * stride_t;
    R mx = x[0];
    for (int v = 1; v < V; ++v) mx = std::max<R>(mx, x[v]);
    R sum = 0;
    for (int v = 0; v < V; ++v) sum += std::exp(R(x[v]) - mx);
    const R lse = mx + std::log(sum);
    for (int v = 0; v < V; ++v) lp[(size_t)t * V + v] = R(x[v]) - lse;
  }
  auto sym = [&](int s) { return (s & 1) ? lab[s >> 1] : blank; };

  // alpha sweep
  alpha[0] = lp[sym(0)];
  if (S > 1) alpha[1] = lp[sym(1)];
  for (int t = 1; t < T; ++t) {
    const R* prev = &alpha[(size_t)(t - 1) * S];
    R* cur = &alpha[(size_t)t * S];
    const int lo = std::max(0, S - 2 * (T - t)), hi = std::min(S, 2 * (t + 1));  // reachable band
    for (int s = lo; s < hi; ++s) {
      R a = prev[s];
      if (s >= 1) a = log_add(a, prev[s - 1]);
      if (s >= 2 && (s & 1) && sym(s) != sym(s - 2)) a = log_add(a, prev[s - 2]);
      cur[s] = a == ninf ? ninf : a + lp[(size_t)t * V + sym(s)];
    }
  }
  R ll = alpha[(size_t)(T - 1) * S + S - 1];
  if (S > 1) ll = log_add(ll, alpha[(size_t)(T - 1) * S + S - 2]);
  if (ll == ninf) return pinf;
  if (!grads) return -ll;

  // beta sweep fused with occupancy and gradient
  std::fill(beta_next.begin(), beta_next.end(), ninf);
  for (int t = T - 1; t >= 0; --t) {
    std::fill(beta.begin(), beta.end(), ninf);
    const int lo = std::max(0, S - 2 * (T - t)), hi = std::min(S, 2 * (t + 1));
    for (int s = lo; s < hi; ++s) {
      R b;
      if (t == T - 1) {
        b = (s >= S - 2) ? R(0) : ninf;
      } else {
        b = beta_next[s];
        if (s + 1 < S) b = log_add(b, beta_next[s + 1]);
        if (s + 2 < S && (s & 1) && sym(s) != sym(s + 2)) b = log_add(b, beta_next[s + 2]);
      }
      beta[s] = b == ninf ? ninf : b + lp[(size_t)t * V + sym(s)];
    }
    std::fill(occ.begin(), occ.end(), ninf);
    const R* al = &alpha[(size_t)t * S];
    for (int s = lo; s < hi; ++s) occ[sym(s)] = log_add(occ[sym(s)], al[s] + beta[s]);
    float* g = grads + t * stride_t;
    for (int v = 0; v < V; ++v) {
      const R l = lp[(size_t)t * V + v];
      const R y = std::exp(l);
      g[v] = (float)(occ[v] == ninf ? y : y - std::exp(occ[v] - l - ll));
    }
    std::swap(beta, beta_next);
  }
  return -ll;
}

template <typename R>
int run_batch(const float* acts, float* grads, const int* flat_labels, const int* label_lens,
              const int* act_lens, int T, int B, int V, int blank, float* costs, int num_threads) {
  std::vector<int64_t> off(B + 1, 0);
  for (int b = 0; b < B; ++b) off[b + 1] = off[b] + label_lens[b];
  const int64_t stride_t = (int64_t)B * V;
  if (grads) {
    // rows t >= act_lens[b] stay zero (the reference wrapper pre-zeros grads, ctc.py:36)
#pragma omp parallel for num_threads(num_threads) schedule(static)
    for (int64_t i = 0; i < (int64_t)T * B * V; ++i) grads[i] = 0.f;
  }
#pragma omp parallel for num_threads(num_threads) schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    costs[b] = (float)cost_and_grad_one<R>(acts + (int64_t)b * V, grads ? grads + (int64_t)b * V : nullptr,
                                           stride_t, flat_labels + off[b], label_lens[b], act_lens[b], V,
                                           blank);
  }
  return 0;
}

}  // namespace

extern "C" {

int oracle_ctc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// acts/grads: contiguous [T,B,V] float32 on the host; grads may be null (cost only).
int oracle_ctc_cpu_f32(const float* acts, float* grads, const int* flat_labels, const int* label_lens,
                       const int* act_lens, int T, int B, int V, int blank, float* costs, int num_threads) {
  return run_batch<float>(acts, grads, flat_labels, label_lens, act_lens, T, B, V, blank, costs, num_threads);
}

int oracle_ctc_cpu_f64(const float* acts, float* grads, const int* flat_labels, const int* label_lens,
                       const int* act_lens, int T, int B, int V, int blank, float* costs, int num_threads) {
  return run_batch<double>(acts, grads, flat_labels, label_lens, act_lens, T, B, V, blank, costs, num_threads);
}

}  // extern "C"
