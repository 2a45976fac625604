// Fast lattice: block-exponent fp32 alpha/beta recursion, forward and backward sweeps running
// concurrently in one CTA and meeting in the middle.
//
// Arithmetic.  The recursion of SURVEY Appendix A is evaluated in the LINEAR domain:
//     alpha_t(s) = y_t(l'_s) * (alpha_{t-1}(s) + alpha_{t-1}(s-1) + [skip] alpha_{t-1}(s-2))
// Each lane owns four consecutive lattice states as fp32 mantissas plus ONE shared int32
// power-of-two exponent that is renormalised after every frame, so the representable range is
// unbounded while the inner loop is pure FADD/FMUL + integer exponent arithmetic: no exp/log at
// all (the softmax probabilities y come from K1).  Relative rounding error is ~6e-8 per operation
// independent of |log alpha| -- this is what keeps T=1500 utterances inside the 1e-4 gradient
// tolerance where an fp32 log-space recursion does not ("numerics").
// The one weakness -- a state more than ~2^-110 below its group's largest state loses bits -- is
// harmless unless that state could carry posterior mass; phase 2 bounds that mass for every group
// and frame, and if the bound is not negligible (FLAG_PRECISION_LOST) the utterance is redone by
// the fp64 safe lattice in the same CTA.
//
// Schedule.  One CTA per utterance.  Warps [0,NW) sweep forward (alpha, t = 0,1,..), warps
// [NWMAX, NWMAX+NW) sweep backward (beta, t = T-1,T-2,..; beta is the same recursion on the
// reversed label sequence).  Phase 1: each side covers half of the frames and stores its
// pre-emission values to the scratch.  Phase 2 (after one CTA barrier): each side continues through
// the other half, multiplies its fresh values with the stored ones of the opposite side --
// posterior(t,s) = alpha_t(s) * beta'_t(s) / P -- and subtracts the per-symbol occupancy from the
// gradient row (which K1 filled with the softmax) with one RED per (frame, symbol).  Sequential
// depth is T frames instead of 2T and only half of alpha and beta ever goes through HBM.
//
// Lattice layout.  Lane l of warp w holds positions base_w + 4l .. +3 (a "group"), base_w =
// w*(128-2K): consecutive warp windows overlap by a halo of 2K positions.  Dependencies only point
// downwards (s-1, s-2), so a warp can run K frames without talking to its neighbour while the
// garbage creeping up from its window bottom stays inside the halo; every K frames ("chunk") the
// warps exchange halos through shared memory -- ONE block barrier per K frames.  Neighbour states
// inside a warp travel by __shfl_up.  Emission rows and the opposite side's stored rows for chunk
// c+1 are prefetched with cp.async while chunk c computes.
#pragma once

#include "lattice_common.cuh"
#include "lattice_safe.cuh"

namespace b200ctc {

constexpr int kEZero = -(1 << 28);  // exponent of an all-zero group

__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 2^d for d <= 0 (0 when d < -126)
__device__ __forceinline__ float pow2_neg(int d) { return __int_as_float(max(d + 127, 0) << 23); }
// 2^d clamped to [2^-127 -> 0, 2^127]
__device__ __forceinline__ float pow2_clamped(int d) { return __int_as_float(min(max(d + 127, 0), 254) << 23); }

struct FastGeom {
  int T, L, S, J, J4, P;   // P = 4J positions; J4 = J rounded up to 4 (exponent row stride)
  int NW;                  // active warps per side
  int RW;                  // staged emission row stride (floats, multiple of 4)
};

template <int K>
__host__ __device__ inline int fast_warps_needed(int L) {
  const int P = 4 * ((2 * L + 1 + 3) / 4);
  const int own = 128 - 2 * K;
  return P <= 128 ? 1 : 1 + (P - 128 + own - 1) / own;
}

// shared memory carve-up -------------------------------------------------------------------------
struct FastSideSmem {
  float* rows;    // [2][K][RW]   staged emission rows
  float4* oth_m;  // [2][K][J]    opposite side's stored mantissas
  int* oth_e;     // [2][K][J4]   opposite side's stored exponents
  float* post;    // [2][K][P]    posteriors of the frames of a chunk
  float4* halo_m; // [2][NWMAX][K/2] halo groups
  int* halo_e;    // [2][NWMAX][K/2]
  float* red_m;   // [NWMAX]
  int* red_e;     // [NWMAX]
  int* ctl;       // [4] per-side control words
};

template <int K, int NWMAX>
__host__ __device__ inline size_t fast_side_bytes(int J, int RW) {
  const size_t J4 = (size_t)(J + 3) / 4 * 4, P = 4 * (size_t)J;
  size_t b = 0;
  b += 2 * K * (size_t)RW * 4;
  b += 2 * K * (size_t)J * 16;
  b += 2 * K * J4 * 4;
  b += 2 * K * P * 4;
  b += 2 * NWMAX * (K / 2) * 16 + 2 * NWMAX * (K / 2) * 4;
  b += NWMAX * 8 + 16;
  return (b + 15) / 16 * 16;
}
template <int K, int NWMAX>
__host__ __device__ inline size_t fast_smem_bytes(int L, int RW) {
  const int J = (2 * L + 1 + 3) / 4;
  size_t shared = (size_t)(4 * L + 8) * 4 + (size_t)(L + 4) * 4 + 64;  // lab, sorted, seg_start, seg_sym, rep, ctl
  shared = (shared + 15) / 16 * 16;
  return shared + 2 * fast_side_bytes<K, NWMAX>(J, RW) + 16;
}

template <int K, int NWMAX>
__device__ __forceinline__ FastSideSmem carve_fast_side(unsigned char* base, int J, int RW) {
  const int J4 = (J + 3) / 4 * 4, P = 4 * J;
  FastSideSmem s;
  unsigned char* p = base;
  s.oth_m = reinterpret_cast<float4*>(p);  p += 2 * K * (size_t)J * 16;
  s.halo_m = reinterpret_cast<float4*>(p); p += 2 * NWMAX * (K / 2) * 16;
  s.rows = reinterpret_cast<float*>(p);    p += 2 * K * (size_t)RW * 4;
  s.post = reinterpret_cast<float*>(p);    p += 2 * K * (size_t)P * 4;
  s.oth_e = reinterpret_cast<int*>(p);     p += 2 * K * (size_t)J4 * 4;
  s.halo_e = reinterpret_cast<int*>(p);    p += 2 * NWMAX * (K / 2) * 4;
  s.red_m = reinterpret_cast<float*>(p);   p += NWMAX * 4;
  s.red_e = reinterpret_cast<int*>(p);     p += NWMAX * 4;
  s.ctl = reinterpret_cast<int*>(p);
  return s;
}

// per-lane constants -------------------------------------------------------------------------------
struct LaneConst {
  int idx[4];       // index of each state's symbol in the staged emission row
  int s0, sdir;     // lattice state of slot i is s0 + i*sdir   (forward: +1, backward: -1)
  int valid;        // bit i: slot i is a real lattice state
  int skip;         // bit i: the s-2 -> s (backward: s+2 -> s) transition is allowed
  int live_from, live_until;  // frames during which ALL four states are reachable and can still finish
  bool owned;       // this lane's group belongs to the warp (not to the halo)
  int group;        // global position group  (pos0 / 4)
  int sgroup;       // group index in lattice-state order used by the scratch rows
};

struct LaneState {
  float v0, v1, v2, v3;
  int e;
};

// One frame of the recursion for one lane.  Outputs the pre-emission sums (acc*, exponent E) and the
// new emission-weighted values w* at the same exponent; updates st with the renormalised state.
__device__ __forceinline__ void lattice_frame(LaneState& st, const LaneConst& lc, const float* __restrict__ row,
                                              int lane, float& acc0, float& acc1, float& acc2, float& acc3,
                                              float& w0, float& w1, float& w2, float& w3, int& E) {
  // (the largest of w0..w3 decides the new exponent)
  float n1 = __shfl_up_sync(0xffffffffu, st.v3, 1);
  float n2 = __shfl_up_sync(0xffffffffu, st.v2, 1);
  int ne = __shfl_up_sync(0xffffffffu, st.e, 1);
  if (lane == 0) { n1 = 0.f; n2 = 0.f; ne = kEZero; }
  E = max(st.e, ne);
  const float so = pow2_neg(st.e - E), sn = pow2_neg(ne - E);
  const float a0 = st.v0 * so, a1 = st.v1 * so, a2 = st.v2 * so, a3 = st.v3 * so;
  const float b1 = n1 * sn, b2 = n2 * sn;
  acc0 = a0 + b1 + ((lc.skip & 1) ? b2 : 0.f);
  acc1 = a1 + a0 + ((lc.skip & 2) ? b1 : 0.f);
  acc2 = a2 + a1 + ((lc.skip & 4) ? a0 : 0.f);
  acc3 = a3 + a2 + ((lc.skip & 8) ? a1 : 0.f);
  const float y0 = (lc.valid & 1) ? row[lc.idx[0]] : 0.f;
  const float y1 = (lc.valid & 2) ? row[lc.idx[1]] : 0.f;
  const float y2 = (lc.valid & 4) ? row[lc.idx[2]] : 0.f;
  const float y3 = (lc.valid & 8) ? row[lc.idx[3]] : 0.f;
  w0 = acc0 * y0; w1 = acc1 * y1; w2 = acc2 * y2; w3 = acc3 * y3;
  const float mx = fmaxf(fmaxf(w0, w1), fmaxf(w2, w3));
  if (mx > 0.f) {
    const int eb = __float_as_int(mx) >> 23;                 // biased exponent (mx > 0)
    const float sc = __int_as_float((254 - eb) << 23);        // 2^(127-eb): largest mantissa -> [1,2)
    st.v0 = w0 * sc; st.v1 = w1 * sc; st.v2 = w2 * sc; st.v3 = w3 * sc;
    st.e = E + eb - 127;
  } else {
    st.v0 = st.v1 = st.v2 = st.v3 = 0.f;
    st.e = kEZero;
  }
}

struct FastCtx {
  const CallParams* p;
  UttMeta m;
  FastGeom g;
  int b, side, w, lane, tid_side, nt_side;
  FastSideSmem sm;
  const int* lab;
  SymbolIndex ix;
  float4* scr_m;   // [T][J]
  int* scr_e;      // [T][J4]
  int row_vec;     // floats per cp.async for emission rows (1, 2 or 4)
  int row_len;     // floats to copy per row
  const float* row_src; long long row_stride;  // emission row source: element (t) at row_src + t*row_stride
};

// frame index of step n of this side
__device__ __forceinline__ int frame_of(const FastCtx& c, int n) { return c.side ? c.g.T - 1 - n : n; }

template <int K>
__device__ __forceinline__ void stage_rows(const FastCtx& c, int buf, int n0, int kc) {
  float* dst = c.sm.rows + (size_t)buf * K * c.g.RW;
  const int per_row = c.row_len / c.row_vec;
  for (int i = c.tid_side; i < kc * per_row; i += c.nt_side) {
    const int j = i / per_row, e = (i - j * per_row) * c.row_vec;
    const float* src = c.row_src + (long long)frame_of(c, n0 + j) * c.row_stride + e;
    float* d = dst + j * c.g.RW + e;
    if (c.row_vec == 4) cp_async_16(d, src);
    else if (c.row_vec == 2) cp_async_8(d, src);
    else cp_async_4(d, src);
  }
}

template <int K>
__device__ __forceinline__ void stage_other(const FastCtx& c, int buf, int n0, int kc) {
  float4* dm = c.sm.oth_m + (size_t)buf * K * c.g.J;
  int* de = c.sm.oth_e + (size_t)buf * K * c.g.J4;
  const int J = c.g.J, J4 = c.g.J4, E4 = J4 / 4;
  for (int i = c.tid_side; i < kc * J; i += c.nt_side) {
    const int j = i / J, g = i - j * J;
    cp_async_16(dm + j * J + g, c.scr_m + (long long)frame_of(c, n0 + j) * J + g);
  }
  for (int i = c.tid_side; i < kc * E4; i += c.nt_side) {
    const int j = i / E4, g = (i - j * E4) * 4;
    cp_async_16(de + j * J4 + g, c.scr_e + (long long)frame_of(c, n0 + j) * J4 + g);
  }
}

// Occupancy update for the frames of one finished phase-2 chunk: one deterministic sum and one RED
// per (frame, symbol).  Executed by all threads of the side.
template <int K>
__device__ __forceinline__ void reduce_chunk(const FastCtx& c, int pbuf, int n0, int kc) {
  const CallParams& p = *c.p;
  const int P = c.g.P, NW = c.g.NW;
  const float* post = c.sm.post + (size_t)pbuf * K * P;
  // blank: positions whose lattice state is even.  forward: q even; backward: q odd (4J-1-q even).
  for (int j = c.w; j < kc; j += NW) {
    const float4* row4 = reinterpret_cast<const float4*>(post + (size_t)j * P);
    float acc = 0.f;
    for (int gq = c.lane; gq < c.g.J; gq += 32) {
      const float4 q = row4[gq];
      acc += c.side ? (q.y + q.w) : (q.x + q.z);
    }
    acc = warp_sum(acc);
    if (c.lane == 0) {
      float* grow = p.grads + ((long long)frame_of(c, n0 + j) * p.B + c.b) * p.V;
      atomicAdd(grow + p.blank, -acc);
    }
  }
  const int n_seg = *c.ix.n_seg;
  for (int i = c.tid_side; i < kc * n_seg; i += c.nt_side) {
    const int j = i / n_seg, u = i - j * n_seg;
    const float* row = post + (size_t)j * P;
    float acc = 0.f;
    for (int k = c.ix.seg_start[u]; k < c.ix.seg_start[u + 1]; ++k) {
      const int s = 2 * c.ix.sorted[k] + 1;
      acc += row[c.side ? (P - 1 - s) : s];
    }
    float* grow = p.grads + ((long long)frame_of(c, n0 + j) * p.B + c.b) * p.V;
    atomicAdd(grow + c.ix.seg_sym[u], -acc);
  }
}

// Everything a side's warps carry through the sweep.
struct SweepState {
  LaneState st;
  LaneConst lc;
  int win_s_lo, win_s_hi;   // lattice-state range of the warp window (band skip)
  float inv_mP; int eP;     // total probability P = mP * 2^eP (phase 2)
  bool lost;
};

// One chunk (kc <= K frames starting at step n0, staged in buffer `buf`).
template <int K, bool PH2>
__device__ __forceinline__ void run_chunk(const FastCtx& c, SweepState& ss, int buf, int n0, int kc,
                                          bool write_post) {
  const int T = c.g.T, S = c.g.S, J = c.g.J, J4 = c.g.J4, P = c.g.P, RW = c.g.RW;
  const LaneConst& lc = ss.lc;
  const float* rows = c.sm.rows + (size_t)buf * K * RW;
  for (int j = 0; j < kc; ++j) {
    const int t = frame_of(c, n0 + j);
    const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
    const bool in_band = !(ss.win_s_hi < lo_t || ss.win_s_lo >= hi_t);   // warp-uniform
    float* post_row = c.sm.post + ((size_t)buf * K + j) * P;
    if (in_band) {
      float a0, a1, a2, a3, w0, w1, w2, w3; int E;
      lattice_frame(ss.st, lc, rows + j * RW, c.lane, a0, a1, a2, a3, w0, w1, w2, w3, E);
      if (lc.owned && lc.group < J) {
        if (!PH2) {
          const float4 o = c.side ? make_float4(a3, a2, a1, a0) : make_float4(a0, a1, a2, a3);
          c.scr_m[(long long)t * J + lc.sgroup] = o;
          c.scr_e[(long long)t * J4 + lc.sgroup] = E;
        } else {
          float4 om = c.sm.oth_m[((size_t)buf * K + j) * J + lc.sgroup];
          const int oe = c.sm.oth_e[((size_t)buf * K + j) * J4 + lc.sgroup];
          if (c.side) { float tx = om.x; om.x = om.w; om.w = tx; tx = om.y; om.y = om.z; om.z = tx; }
          // posterior = w * om * 2^dexp / mP.  Both mantissas may be far below 1 (their group's
          // maximum is elsewhere), so dexp can legitimately exceed 127: apply it in two halves.
          const int dexp = E + oe - ss.eP;
          const int dhalf = dexp >> 1;
          const float sa = pow2_clamped(min(dhalf, 120)) * ss.inv_mP;             // inv_mP in (0.5, 1]
          const float sb = pow2_clamped(min(dexp - dhalf, 120));
          const int s_a = lc.s0, d = lc.sdir;
          // states outside the reachable band carry dead (own side) or never-written (other side) values
          const bool b0 = (s_a >= lo_t && s_a < hi_t), b1 = (s_a + d >= lo_t && s_a + d < hi_t);
          const bool b2 = (s_a + 2 * d >= lo_t && s_a + 2 * d < hi_t), b3 = (s_a + 3 * d >= lo_t && s_a + 3 * d < hi_t);
          const float u0 = b0 ? w0 : 0.f, u1 = b1 ? w1 : 0.f, u2 = b2 ? w2 : 0.f, u3 = b3 ? w3 : 0.f;
          const float o0 = b0 ? om.x : 0.f, o1 = b1 ? om.y : 0.f, o2 = b2 ? om.z : 0.f, o3 = b3 ? om.w : 0.f;
          float4 po;
          po.x = (u0 * sa) * (o0 * sb); po.y = (u1 * sa) * (o1 * sb);
          po.z = (u2 * sa) * (o2 * sb); po.w = (u3 * sa) * (o3 * sb);
          // Range check.  A state that sits more than 2^-110 below its group's largest value may
          // have lost bits (on either side).  Its posterior is bounded by
          //   2^-110 * max(own group) * max(other group) * 2^dexp / mP;
          // if that bound is not negligible (> 2^-24) the block-exponent result cannot be trusted.
          // The own maximum runs over ALL four states: dead states (too late to finish) share the
          // exponent.  Evaluated on the exponent fields, so it cannot overflow or underflow.
          const float umax = fmaxf(fmaxf(w0, w1), fmaxf(w2, w3));
          const float omax = fmaxf(fmaxf(o0, o1), fmaxf(o2, o3));
          const int bound = (__float_as_int(umax) >> 23) + (__float_as_int(omax) >> 23) - 254 + dexp;
          ss.lost |= (umax > 0.f) && (omax > 0.f) && (bound > 110 - 24 - 2);
          if (write_post) *reinterpret_cast<float4*>(post_row + 4 * lc.group) = po;
        }
      }
    } else if (PH2 && write_post && lc.owned && lc.group < J) {
      *reinterpret_cast<float4*>(post_row + 4 * lc.group) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Chunk boundary: publish the halo, take the abort snapshot, ONE side barrier, import the halo.
// Returns true when the side must leave the fast path.
template <int K, int NWMAX>
__device__ __forceinline__ bool chunk_boundary(const FastCtx& c, SweepState& ss, int cc, int* abort_flag,
                                               int* abort_seen) {
  constexpr int HG = K / 2;
  const int hb = cc & 1, NW = c.g.NW, w = c.w, lane = c.lane;
  if (w + 1 < NW && lane >= 32 - HG) {
    const int slot = (hb * NWMAX + w) * HG + (lane - (32 - HG));
    c.sm.halo_m[slot] = make_float4(ss.st.v0, ss.st.v1, ss.st.v2, ss.st.v3);
    c.sm.halo_e[slot] = ss.st.e;
  }
  if (__any_sync(0xffffffffu, ss.lost) && lane == 0) *abort_flag = 1;
  if (c.tid_side == 0) abort_seen[c.side * 2 + hb] = *(volatile int*)abort_flag;
  cp_async_wait_all();
  named_bar(1 + c.side, NW * 32);
  if (w > 0 && lane < HG) {
    const int slot = (hb * NWMAX + (w - 1)) * HG + lane;
    const float4 hv = c.sm.halo_m[slot];
    ss.st.v0 = hv.x; ss.st.v1 = hv.y; ss.st.v2 = hv.z; ss.st.v3 = hv.w;
    ss.st.e = c.sm.halo_e[slot];
  }
  return abort_seen[c.side * 2 + hb] != 0;
}

// The whole fast path for one utterance; every thread of the CTA calls it.  On return the shared
// word smem_abort[0] is non-zero when the utterance must be redone by the safe lattice (the
// caller reads it after a __syncthreads()).
template <int K, int NWMAX>
__device__ void lattice_fast_utterance(const CallParams& p, int b, unsigned char* smem, int** smem_abort) {
  constexpr int H = 2 * K;          // halo positions
  constexpr int HG = K / 2;         // halo groups (lanes)
  constexpr int OWN = 128 - H;
  static_assert(K % 2 == 0 && K >= 2 && K <= 16, "K must be even");

  FastCtx c;
  c.p = &p;
  c.b = b;
  c.m = p.meta[b];
  const int T = c.m.T, L = c.m.L;
  c.g.T = T; c.g.L = L; c.g.S = 2 * L + 1; c.g.J = c.m.J; c.g.J4 = (c.m.J + 3) / 4 * 4; c.g.P = 4 * c.m.J;
  c.g.NW = fast_warps_needed<K>(L);
  c.g.RW = p.gathered ? c.m.W : (p.V + 3) / 4 * 4;
  const int J = c.g.J, P = c.g.P, S = c.g.S, NW = c.g.NW;
  const int warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.side = warp / NWMAX;
  c.w = warp - c.side * NWMAX;
  c.tid_side = c.w * 32 + c.lane;
  c.nt_side = NW * 32;
  const int lane = c.lane, side = c.side, w = c.w;

  // ---- shared memory: common part, then one block per side ----
  int* ip = reinterpret_cast<int*>(smem);
  int* abort_flag = ip;          ip += 1;     // set by any thread: leave the fast path
  int* abort_seen = ip;          ip += 7;     // [side][parity] snapshots, [4] midpoint snapshot
  int* lab = ip;                 ip += L;
  c.ix.sorted = ip;              ip += L;
  c.ix.seg_start = ip;           ip += L + 1;
  c.ix.seg_sym = ip;             ip += L + 1;
  int* rep = ip;                 ip += L + 1;
  c.ix.n_seg = ip;               ip += 1;
  size_t common = (size_t)(reinterpret_cast<unsigned char*>(ip) - smem);
  common = (common + 15) / 16 * 16;
  const size_t side_bytes = fast_side_bytes<K, NWMAX>(J, c.g.RW);
  c.sm = carve_fast_side<K, NWMAX>(smem + common + side * side_bytes, J, c.g.RW);
  c.lab = lab;
  *smem_abort = abort_flag;

  // ---- prologue (all threads of the CTA) ----
  for (int i = threadIdx.x; i < L; i += blockDim.x) lab[i] = p.labels[c.m.lab_off + i];
  if (threadIdx.x < 8) abort_flag[threadIdx.x] = 0;  // abort_flag + abort_seen[0..6]
  __syncthreads();
  build_symbol_index(lab, L, c.ix);
  // rep[i] = number of j in [1, i] with lab[j] == lab[j-1]   (one warp, ballot scan)
  if (warp == 0) {
    int running = 0;
    for (int base = 0; base < L; base += 32) {
      const int i = base + lane;
      const bool r = (i >= 1 && i < L) && lab[i] == lab[i - 1];
      const unsigned bal = __ballot_sync(0xffffffffu, r);
      if (i < L) rep[i] = running + __popc(bal & (0xffffffffu >> (31 - lane)));
      running += __popc(bal);
    }
  }
  __syncthreads();
  if (w >= NW) return;   // idle warps wait at the caller's __syncthreads()

  // scratch rows
  unsigned char* scr = p.scratch + c.m.scratch_off * kGroupBytes;
  c.scr_m = reinterpret_cast<float4*>(scr);
  c.scr_e = reinterpret_cast<int*>(scr + (size_t)T * J * 16);

  // emission row source
  if (p.gathered) {
    c.row_src = p.em + c.m.em_off; c.row_stride = c.m.W; c.row_len = c.m.W; c.row_vec = 4;
  } else {
    c.row_src = p.grads + (long long)b * p.V; c.row_stride = (long long)p.B * p.V; c.row_len = p.V;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.grads);
    c.row_vec = (p.V % 4 == 0 && a % 16 == 0) ? 4 : ((p.V % 2 == 0 && a % 8 == 0) ? 2 : 1);
  }

  // ---- per-lane constants ----
  SweepState ss;
  LaneConst& lc = ss.lc;
  const int base_w = w * OWN;
  const int pos0 = base_w + 4 * lane;
  lc.group = pos0 >> 2;
  lc.owned = (w == 0) || (lane >= HG);
  lc.s0 = side ? (P - 1 - pos0) : pos0;
  lc.sdir = side ? -1 : 1;
  lc.sgroup = side ? (J - 1 - lc.group) : lc.group;
  lc.valid = 0; lc.skip = 0;
  {
    const int total_rep = (L > 0) ? rep[L - 1] : 0;
    int from = 0, until = T - 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = pos0 + i;
      const int s = lc.s0 + i * lc.sdir;
      const bool ok = (q < P) && (s >= 0) && (s < S);
      int idx = 0;
      if (ok) {
        lc.valid |= 1 << i;
        const int li = s >> 1;                       // label index of an odd state
        if (s & 1) {
          idx = p.gathered ? li + 1 : lab[li];
          const bool sk = side ? (s + 2 < S && lab[li] != lab[li + 1]) : (s >= 3 && lab[li] != lab[li - 1]);
          if (sk) lc.skip |= 1 << i;
          from = max(from, li + rep[li]);
          until = min(until, T - 1 - ((L - 1 - li) + (total_rep - rep[li])));
        } else {
          idx = p.gathered ? 0 : p.blank;            // blank before label li (li == L: the final blank)
          from = max(from, (li == 0) ? 0 : li + rep[li - 1]);
          until = min(until, T - 1 - ((li == L) ? 0 : (L - li) + (total_rep - rep[li])));
        }
      } else {
        from = T;  // a group with padding positions is never "fully live"
      }
      lc.idx[i] = idx;
    }
    lc.live_from = from; lc.live_until = until;
  }
  {
    const int win_lo_pos = base_w, win_hi_pos = min(base_w + 127, P - 1);
    ss.win_s_lo = side ? (P - 1 - win_hi_pos) : win_lo_pos;
    ss.win_s_hi = side ? (P - 1 - win_lo_pos) : win_hi_pos;
  }
  // ---- initial state: delta on the first lattice state of this side's sweep ----
  ss.st.v0 = ss.st.v1 = ss.st.v2 = ss.st.v3 = 0.f; ss.st.e = kEZero;
  ss.lost = false; ss.inv_mP = 0.f; ss.eP = 0;
  {
    const int q_start = side ? (P - S) : 0;   // backward: 4J - S dummy positions come first
    if (w == 0 && q_start >= pos0 && q_start < pos0 + 4) {
      const int i = q_start - pos0;
      if (i == 0) ss.st.v0 = 1.f; else if (i == 1) ss.st.v1 = 1.f; else if (i == 2) ss.st.v2 = 1.f; else ss.st.v3 = 1.f;
      ss.st.e = 0;
    }
  }

  const int M_side = side ? (T / 2) : (T - T / 2);      // frames this side covers in phase 1
  const int nc1 = (M_side + K - 1) / K;
  const int nc2 = (T - M_side + K - 1) / K;
  const int n_chunks = nc1 + nc2;
  const int bar_id = 1 + side;
  const int n_side_threads = NW * 32;
  auto chunk_n0 = [&](int cc) { return cc < nc1 ? cc * K : M_side + (cc - nc1) * K; };
  auto chunk_kc = [&](int cc) { return cc < nc1 ? min(K, M_side - cc * K) : min(K, T - (M_side + (cc - nc1) * K)); };

  // emission rows of the first chunk
  if (n_chunks > 0) stage_rows<K>(c, 0, chunk_n0(0), chunk_kc(0));
  cp_async_commit();
  cp_async_wait_all();
  named_bar(bar_id, n_side_threads);

  bool aborted = false;

  // ================================ phase 1 ================================
  for (int cc = 0; cc < nc1 && !aborted; ++cc) {
    if (cc + 1 < n_chunks) stage_rows<K>(c, (cc + 1) & 1, chunk_n0(cc + 1), chunk_kc(cc + 1));
    cp_async_commit();
    run_chunk<K, false>(c, ss, cc & 1, chunk_n0(cc), chunk_kc(cc), false);
    aborted = chunk_boundary<K, NWMAX>(c, ss, cc, abort_flag, abort_seen);
  }

  // ================================ midpoint ================================
  // Both sides always meet here exactly once (even when one of them has already given up).
  if (threadIdx.x == 0) abort_seen[4] = *(volatile int*)abort_flag;
  named_bar(3, 2 * n_side_threads);
  if (abort_seen[4]) aborted = true;
  if (aborted || nc2 == 0) return;

  // the opposite side's rows for the first phase-2 chunk could not be prefetched earlier
  stage_other<K>(c, nc1 & 1, chunk_n0(nc1), chunk_kc(nc1));
  cp_async_commit();
  cp_async_wait_all();
  named_bar(bar_id, n_side_threads);

  // ---- total probability P = sum_s alpha_t(s) beta'_t(s) at the first phase-2 frame (state copy) ----
  {
    const int buf = nc1 & 1, n0 = chunk_n0(nc1);
    LaneState tmp = ss.st;
    float part = 0.f; int pe = kEZero;
    const int t = frame_of(c, n0);
    const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
    if (!(ss.win_s_hi < lo_t || ss.win_s_lo >= hi_t)) {
      float a0, a1, a2, a3, w0, w1, w2, w3; int E;
      lattice_frame(tmp, lc, c.sm.rows + (size_t)buf * K * c.g.RW, lane, a0, a1, a2, a3, w0, w1, w2, w3, E);
      if (lc.owned && lc.group < J) {
        float4 om = c.sm.oth_m[(size_t)buf * K * J + lc.sgroup];
        const int oe = c.sm.oth_e[(size_t)buf * K * c.g.J4 + lc.sgroup];
        if (side) { float tx = om.x; om.x = om.w; om.w = tx; tx = om.y; om.y = om.z; om.z = tx; }
        const int s_a = lc.s0, d = lc.sdir;
        float sum = 0.f;
        if (s_a >= lo_t && s_a < hi_t) sum += w0 * om.x;
        if (s_a + d >= lo_t && s_a + d < hi_t) sum += w1 * om.y;
        if (s_a + 2 * d >= lo_t && s_a + 2 * d < hi_t) sum += w2 * om.z;
        if (s_a + 3 * d >= lo_t && s_a + 3 * d < hi_t) sum += w3 * om.w;
        if (sum > 0.f) { part = sum; pe = E + oe; }
      }
    }
    int emax = pe;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    float scaled = part * pow2_neg(pe - emax);
    scaled = warp_sum(scaled);
    if (lane == 0) { c.sm.red_m[w] = scaled; c.sm.red_e[w] = emax; }
    named_bar(bar_id, n_side_threads);
    int Emax = kEZero;
    for (int i = 0; i < NW; ++i) Emax = max(Emax, c.sm.red_e[i]);
    float tot = 0.f;
    for (int i = 0; i < NW; ++i) tot += c.sm.red_m[i] * pow2_neg(c.sm.red_e[i] - Emax);
    if (!(tot > 0.f) || !(tot < INFINITY) || Emax <= kEZero / 2) {
      // zero / underflowed / garbage total probability: the safe lattice decides
      if (c.tid_side == 0) *abort_flag = 1;
      aborted = true;                          // every thread of the side computed the same `tot`
    } else {
      // normalise P = tot * 2^Emax to a mantissa in [1,2)
      const int eb = (__float_as_int(tot) >> 23) - 127;
      const float mP = tot * pow2_clamped(-eb);
      ss.inv_mP = 1.0f / mP; ss.eP = Emax + eb;
      if (side == 1 && c.tid_side == 0)
        p.costs[b] = (float)(-((double)Emax + log2((double)tot)) * 0.69314718055994530942);
    }
  }
  if (aborted) return;
  // cost-only calls still walk phase 2 (for the range check) but neither store posteriors nor update rows
  const bool write_post = p.grads != nullptr;

  // ================================ phase 2 ================================
  for (int cc = nc1; cc < n_chunks && !aborted; ++cc) {
    if (cc + 1 < n_chunks) {
      stage_rows<K>(c, (cc + 1) & 1, chunk_n0(cc + 1), chunk_kc(cc + 1));
      stage_other<K>(c, (cc + 1) & 1, chunk_n0(cc + 1), chunk_kc(cc + 1));
    }
    cp_async_commit();
    run_chunk<K, true>(c, ss, cc & 1, chunk_n0(cc), chunk_kc(cc), write_post);
    aborted = chunk_boundary<K, NWMAX>(c, ss, cc, abort_flag, abort_seen);
    if (!aborted && write_post) reduce_chunk<K>(c, cc & 1, chunk_n0(cc), chunk_kc(cc));
  }
}

}  // namespace b200ctc
