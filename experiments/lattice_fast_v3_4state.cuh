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
// inside a warp travel by __shfl_up.  Emission rows and the opposite side's stored groups for
// chunk c+1 are prefetched with cp.async while chunk c computes.
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

template <int K>
__host__ __device__ inline int fast_warps_needed(int L) {
  const int P = 4 * ((2 * L + 1 + 3) / 4);
  const int own = 128 - 2 * K;
  return P <= 128 ? 1 : 1 + (P - 128 + own - 1) / own;
}

// ---------------------------------------------------------------------------------------------
// shared memory
// ---------------------------------------------------------------------------------------------
struct FastSideSmem {
  float* rows;    // [2][K][RWS]     staged emission rows (+ a zero slot at index RW)
  float4* oth_m;  // [2][K][NT]      opposite side's stored mantissas, one slot per thread
  int* oth_e;     // [2][K][NT]      opposite side's stored exponents
  float* post;    // [2][K][P]       posteriors of the frames of a chunk
  float4* halo_m; // [2][NWMAX][K/2] halo groups
  int* halo_e;    // [2][NWMAX][K/2]
  float* red_m;   // [NWMAX]
  int* red_e;     // [NWMAX]
  int* pos;       // [L]             post-row position of the k-th label in symbol order
};

template <int K, int NWMAX>
__host__ __device__ inline size_t fast_side_bytes(int L, int RW) {
  const size_t J = (size_t)(2 * L + 1 + 3) / 4, P = 4 * J, NT = NWMAX * 32;
  size_t b = 0;
  b += 2 * K * NT * 16;                   // oth_m
  b += 2 * NWMAX * (K / 2) * 16;          // halo_m
  b += 2 * K * (size_t)(RW + 4) * 4;      // rows
  b += 2 * K * P * 4;                     // post
  b += 2 * K * NT * 4;                    // oth_e
  b += 2 * NWMAX * (K / 2) * 4;           // halo_e
  b += NWMAX * 8;                         // red
  b += (size_t)(L + 4) * 4;               // pos
  return (b + 15) / 16 * 16;
}
template <int K, int NWMAX>
__host__ __device__ inline size_t fast_smem_bytes(int L, int RW) {
  size_t common = (size_t)(8 + 4 * L + 8) * 4;  // control words, lab, sorted, seg_start, seg_sym
  common = (common + 15) / 16 * 16;
  return common + 2 * fast_side_bytes<K, NWMAX>(L, RW) + 16;
}

template <int K, int NWMAX>
__device__ __forceinline__ FastSideSmem carve_fast_side(unsigned char* base, int L, int RW) {
  const size_t J = (size_t)(2 * L + 1 + 3) / 4, P = 4 * J, NT = NWMAX * 32;
  FastSideSmem s;
  unsigned char* p = base;
  s.oth_m = reinterpret_cast<float4*>(p);  p += 2 * K * NT * 16;
  s.halo_m = reinterpret_cast<float4*>(p); p += 2 * NWMAX * (K / 2) * 16;
  s.rows = reinterpret_cast<float*>(p);    p += 2 * K * (size_t)(RW + 4) * 4;
  s.post = reinterpret_cast<float*>(p);    p += 2 * K * P * 4;
  s.oth_e = reinterpret_cast<int*>(p);     p += 2 * K * NT * 4;
  s.halo_e = reinterpret_cast<int*>(p);    p += 2 * NWMAX * (K / 2) * 4;
  s.red_m = reinterpret_cast<float*>(p);   p += NWMAX * 4;
  s.red_e = reinterpret_cast<int*>(p);     p += NWMAX * 4;
  s.pos = reinterpret_cast<int*>(p);
  return s;
}

// ---------------------------------------------------------------------------------------------
// per-lane state
// ---------------------------------------------------------------------------------------------
struct LaneConst {
  int idx0, idx1, idx2, idx3;  // index of each state's symbol in the staged emission row (zero slot if invalid)
  float k0, k1, k2, k3;        // 1.0 where the skip transition into the state is allowed, else 0.0
  int s_lo, s_hi;              // lattice-state range of the group (s_hi = s_lo + 3)
  bool owned;                  // this lane's group belongs to the warp (not to the halo) and exists
  int group;                   // global position group (pos0 / 4)
};

struct LaneState {
  float v0, v1, v2, v3;
  int e;
};

// One frame of the recursion for one lane.  Outputs the pre-emission sums (acc*, exponent E) and the
// new emission-weighted values w* at the same exponent; updates st with the renormalised state.
__device__ __forceinline__ void lattice_frame(LaneState& st, const LaneConst& lc, const float* __restrict__ row,
                                              bool lane0, float& acc0, float& acc1, float& acc2, float& acc3,
                                              float& w0, float& w1, float& w2, float& w3, int& E) {
  const float n1 = __shfl_up_sync(0xffffffffu, st.v3, 1);
  const float n2 = __shfl_up_sync(0xffffffffu, st.v2, 1);
  int ne = __shfl_up_sync(0xffffffffu, st.e, 1);
  if (lane0) ne = kEZero;                      // nothing below the window: scales n1, n2 to zero
  E = max(st.e, ne);
  const float so = pow2_neg(st.e - E), sn = pow2_neg(ne - E);
  const float a0 = st.v0 * so, a1 = st.v1 * so, a2 = st.v2 * so, a3 = st.v3 * so;
  const float b1 = n1 * sn, b2 = n2 * sn;
  acc0 = fmaf(lc.k0, b2, a0 + b1);
  acc1 = fmaf(lc.k1, b1, a1 + a0);
  acc2 = fmaf(lc.k2, a0, a2 + a1);
  acc3 = fmaf(lc.k3, a1, a3 + a2);
  w0 = acc0 * row[lc.idx0];
  w1 = acc1 * row[lc.idx1];
  w2 = acc2 * row[lc.idx2];
  w3 = acc3 * row[lc.idx3];
  const float mx = fmaxf(fmaxf(w0, w1), fmaxf(w2, w3));
  // renormalise: largest mantissa -> [1,2).  mx == 0 (or NaN from garbage): the group is empty.
  const int eb = __float_as_int(mx) >> 23;                       // biased exponent
  const bool nz = mx > 0.f;
  const float sc = nz ? __int_as_float((254 - eb) << 23) : 0.f;  // 2^(127-eb)
  st.v0 = w0 * sc; st.v1 = w1 * sc; st.v2 = w2 * sc; st.v3 = w3 * sc;
  st.e = nz ? E + eb - 127 : kEZero;
}

template <int SIDE>
struct FastCtx {
  const CallParams* p;
  int b, T, L, S, J, P, NW, RW, RWS;
  int w, lane, tid_side, nt_side;
  FastSideSmem sm;
  SymbolIndex ix;
  float4* scr_m;   // [T][J]   stored pre-emission mantissas, in the READER's group order
  int* scr_e;      // [T][J]
  int row_vec, per_row;                       // emission rows: floats per cp.async, copies per row
  const float* row_src; long long row_stride; // element (t) at row_src + t*row_stride
  __device__ __forceinline__ int frame_of(int n) const { return SIDE ? T - 1 - n : n; }
};

// Stage the emission rows of the kc frames starting at step n0: one warp per frame.
template <int K, int SIDE>
__device__ __forceinline__ void stage_rows(const FastCtx<SIDE>& c, int buf, int n0, int kc) {
  for (int j = c.w; j < kc; j += c.NW) {
    const float* src = c.row_src + (long long)c.frame_of(n0 + j) * c.row_stride;
    float* dst = c.sm.rows + (size_t)(buf * K + j) * c.RWS;
    if (c.row_vec == 4) {
      for (int e = c.lane; e < c.per_row; e += 32) cp_async_16(dst + 4 * e, src + 4 * e);
    } else if (c.row_vec == 2) {
      for (int e = c.lane; e < c.per_row; e += 32) cp_async_8(dst + 2 * e, src + 2 * e);
    } else {
      for (int e = c.lane; e < c.per_row; e += 32) cp_async_4(dst + e, src + e);
    }
  }
}

// Every thread fetches the opposite side's stored group for ITS OWN group and each frame of the
// chunk into its private shared-memory slot (no cross-thread visibility needed).
template <int K, int SIDE>
__device__ __forceinline__ void stage_other(const FastCtx<SIDE>& c, const LaneConst& lc, int buf, int n0, int kc) {
  if (!lc.owned) return;
  const int NT = blockDim.x >> 1;
  float4* dm = c.sm.oth_m + (size_t)buf * K * NT + c.tid_side;
  int* de = c.sm.oth_e + (size_t)buf * K * NT + c.tid_side;
  const long long step = SIDE ? -(long long)c.J : (long long)c.J;
  long long off = (long long)c.frame_of(n0) * c.J + lc.group;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < kc) {
      cp_async_16(dm + j * NT, c.scr_m + off);
      cp_async_4(de + j * NT, c.scr_e + off);
      off += step;
    }
  }
}

// Occupancy update for the frames of one finished phase-2 chunk: one deterministic sum and one RED
// per (frame, symbol).  Executed by all threads of the side; frames are spread over the warps.
template <int K, int SIDE>
__device__ __forceinline__ void reduce_chunk(const FastCtx<SIDE>& c, int pbuf, int n0, int kc) {
  const CallParams& p = *c.p;
  const int P = c.P, NW = c.NW;
  const float* post = c.sm.post + (size_t)pbuf * K * P;
  const int n_seg = *c.ix.n_seg;
  // work item = (frame j, slice): slice 0 is the blank (a whole warp), slices 1.. cover 32 symbols each
  const int n_slices = 1 + (n_seg + 31) / 32;
  for (int item = c.w; item < kc * n_slices; item += NW) {
    const int j = item / n_slices, slice = item - j * n_slices;
    const float* row = post + (size_t)j * P;
    float* grow = p.grads + ((long long)c.frame_of(n0 + j) * p.B + c.b) * p.V;
    if (slice == 0) {
      // blank: lattice states with even index.  forward: even positions; backward: odd positions.
      const float4* row4 = reinterpret_cast<const float4*>(row);
      float acc = 0.f;
      for (int gq = c.lane; gq < c.J; gq += 32) {
        const float4 q = row4[gq];
        acc += SIDE ? (q.y + q.w) : (q.x + q.z);
      }
      acc = warp_sum(acc);
      if (c.lane == 0) atomicAdd(grow + p.blank, -acc);
    } else {
      const int u = (slice - 1) * 32 + c.lane;
      if (u < n_seg) {
        float acc = 0.f;
        const int k1 = c.ix.seg_start[u + 1];
        for (int k = c.ix.seg_start[u]; k < k1; ++k) acc += row[c.sm.pos[k]];
        atomicAdd(grow + c.ix.seg_sym[u], -acc);
      }
    }
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
template <int K, bool PH2, int SIDE>
__device__ __forceinline__ void run_chunk(const FastCtx<SIDE>& c, SweepState& ss, int buf, int n0, int kc,
                                          bool write_post) {
  const int T = c.T, S = c.S, J = c.J, P = c.P;
  const LaneConst& lc = ss.lc;
  const bool lane0 = c.lane == 0;
  const int NT = blockDim.x >> 1;
  const float* rows = c.sm.rows + (size_t)buf * K * c.RWS;
  const float4* oth_m = c.sm.oth_m + (size_t)buf * K * NT + c.tid_side;
  const int* oth_e = c.sm.oth_e + (size_t)buf * K * NT + c.tid_side;
  float* post = c.sm.post + (size_t)buf * K * P + 4 * lc.group;
  int t = c.frame_of(n0);
  // scratch slot of this lane's group for the opposite side's reader (mirrored group order)
  long long scr_off = (long long)t * J + (J - 1 - lc.group);
  const long long scr_step = SIDE ? -(long long)J : (long long)J;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < kc) {
      const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
      const bool in_band = !(ss.win_s_hi < lo_t || ss.win_s_lo >= hi_t);   // warp-uniform
      if (in_band) {
        float a0, a1, a2, a3, w0, w1, w2, w3; int E;
        lattice_frame(ss.st, lc, rows + j * c.RWS, lane0, a0, a1, a2, a3, w0, w1, w2, w3, E);
        if (lc.owned) {
          if (!PH2) {
            // stored reversed at the mirrored group: exactly the slot order of the reader's lane
            c.scr_m[scr_off] = make_float4(a3, a2, a1, a0);
            c.scr_e[scr_off] = E;
          } else {
            const float4 om = oth_m[j * NT];
            const int oe = oth_e[j * NT];
            // posterior = w * om * 2^dexp / mP.  Both mantissas may be far below 1 (their group's
            // maximum is elsewhere), so dexp can legitimately exceed 127: apply it in two halves.
            const int dexp = E + oe - ss.eP;
            const int dhalf = dexp >> 1;
            const float sa = pow2_clamped(min(dhalf, 120)) * ss.inv_mP;   // inv_mP in (0.5, 1]
            const float sb = pow2_clamped(min(dexp - dhalf, 120));
            float u0 = w0, u1 = w1, u2 = w2, u3 = w3;
            float o0 = om.x, o1 = om.y, o2 = om.z, o3 = om.w;
            // States outside the reachable band carry dead (own side) or never-written (other side)
            // values.  Only the warps at the band edges have such lanes.
            const bool all_in = lc.s_lo >= lo_t && lc.s_hi < hi_t;
            if (!__all_sync(__activemask(), all_in)) {
              const int sA = SIDE ? lc.s_hi : lc.s_lo, d = SIDE ? -1 : 1;   // state of slot 0, direction
              const bool b0 = (sA >= lo_t && sA < hi_t), b1 = (sA + d >= lo_t && sA + d < hi_t);
              const bool b2 = (sA + 2 * d >= lo_t && sA + 2 * d < hi_t), b3 = (sA + 3 * d >= lo_t && sA + 3 * d < hi_t);
              u0 = b0 ? u0 : 0.f; u1 = b1 ? u1 : 0.f; u2 = b2 ? u2 : 0.f; u3 = b3 ? u3 : 0.f;
              o0 = b0 ? o0 : 0.f; o1 = b1 ? o1 : 0.f; o2 = b2 ? o2 : 0.f; o3 = b3 ? o3 : 0.f;
            }
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
            if (write_post) *reinterpret_cast<float4*>(post + j * P) = po;
          }
        }
      } else if (PH2 && write_post && lc.owned) {
        *reinterpret_cast<float4*>(post + j * P) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      t += SIDE ? -1 : 1;
      scr_off += scr_step;
    }
  }
}

// Chunk boundary: publish the halo, take the abort snapshot, ONE side barrier, import the halo.
// Returns true when the side must leave the fast path.
template <int K, int NWMAX, int SIDE>
__device__ __forceinline__ bool chunk_boundary(const FastCtx<SIDE>& c, SweepState& ss, int cc, int* abort_flag,
                                               int* abort_seen) {
  constexpr int HG = K / 2;
  const int hb = cc & 1, NW = c.NW, w = c.w, lane = c.lane;
  if (w + 1 < NW && lane >= 32 - HG) {
    const int slot = (hb * NWMAX + w) * HG + (lane - (32 - HG));
    c.sm.halo_m[slot] = make_float4(ss.st.v0, ss.st.v1, ss.st.v2, ss.st.v3);
    c.sm.halo_e[slot] = ss.st.e;
  }
  if (__any_sync(0xffffffffu, ss.lost) && lane == 0) *abort_flag = 1;
  if (c.tid_side == 0) abort_seen[SIDE * 2 + hb] = *(volatile int*)abort_flag;
  cp_async_wait_all();
  named_bar(1 + SIDE, NW * 32);
  if (w > 0 && lane < HG) {
    const int slot = (hb * NWMAX + (w - 1)) * HG + lane;
    const float4 hv = c.sm.halo_m[slot];
    ss.st.v0 = hv.x; ss.st.v1 = hv.y; ss.st.v2 = hv.z; ss.st.v3 = hv.w;
    ss.st.e = c.sm.halo_e[slot];
  }
  return abort_seen[SIDE * 2 + hb] != 0;
}

struct FastCommon {
  int* abort_flag;   // set by any thread: leave the fast path
  int* abort_seen;   // [side][parity] snapshots, [4] midpoint snapshot
  int* lab;
  SymbolIndex ix;
};

// One side's sweep (all warps w < NW of that side).
template <int K, int NWMAX, int SIDE>
__device__ void fast_side_sweep(const CallParams& p, int b, const UttMeta& m, const FastCommon& cm,
                                unsigned char* side_smem, int w, int lane) {
  constexpr int H = 2 * K;          // halo positions
  constexpr int HG = K / 2;         // halo groups (lanes)
  constexpr int OWN = 128 - H;

  FastCtx<SIDE> c;
  c.p = &p; c.b = b;
  const int T = m.T, L = m.L;
  c.T = T; c.L = L; c.S = 2 * L + 1; c.J = m.J; c.P = 4 * m.J;
  c.NW = fast_warps_needed<K>(L);
  c.RW = p.gathered ? m.W : (p.V + 3) / 4 * 4;
  c.RWS = c.RW + 4;
  c.w = w; c.lane = lane; c.tid_side = w * 32 + lane; c.nt_side = c.NW * 32;
  c.ix = cm.ix;
  const int J = c.J, P = c.P, S = c.S, NW = c.NW;
  c.sm = carve_fast_side<K, NWMAX>(side_smem, L, c.RW);
  const int* lab = cm.lab;

  // scratch rows
  unsigned char* scr = p.scratch + m.scratch_off * kGroupBytes;
  c.scr_m = reinterpret_cast<float4*>(scr);
  c.scr_e = reinterpret_cast<int*>(scr + (size_t)T * J * 16);

  // emission row source
  if (p.gathered) {
    c.row_src = p.em + m.em_off; c.row_stride = m.W; c.row_vec = 4; c.per_row = m.W / 4;
  } else {
    c.row_src = p.grads + (long long)b * p.V; c.row_stride = (long long)p.B * p.V;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.grads);
    c.row_vec = (p.V % 4 == 0 && a % 16 == 0) ? 4 : ((p.V % 2 == 0 && a % 8 == 0) ? 2 : 1);
    c.per_row = p.V / c.row_vec;
  }

  // ---- side prologue: zero slots of the row buffers, label positions in this side's post rows ----
  for (int i = c.tid_side; i < 2 * K; i += c.nt_side) c.sm.rows[(size_t)i * c.RWS + c.RW] = 0.f;
  for (int k = c.tid_side; k < L; k += c.nt_side) {
    const int s = 2 * cm.ix.sorted[k] + 1;
    c.sm.pos[k] = SIDE ? (P - 1 - s) : s;
  }

  // ---- per-lane constants ----
  SweepState ss;
  LaneConst& lc = ss.lc;
  const int base_w = w * OWN;
  const int pos0 = base_w + 4 * lane;
  lc.group = pos0 >> 2;
  lc.owned = ((w == 0) || (lane >= HG)) && (lc.group < J);
  lc.s_lo = SIDE ? (P - 1 - pos0 - 3) : pos0;
  lc.s_hi = lc.s_lo + 3;
  {
    int idx[4]; float kk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = pos0 + i;
      const int s = SIDE ? (P - 1 - q) : q;
      const bool ok = (q < P) && (s >= 0) && (s < S);
      idx[i] = c.RW;   // zero slot
      kk[i] = 0.f;
      if (ok) {
        const int li = s >> 1;                       // label index of an odd state
        if (s & 1) {
          idx[i] = p.gathered ? li + 1 : lab[li];
          const bool sk = SIDE ? (s + 2 < S && lab[li] != lab[li + 1]) : (s >= 3 && lab[li] != lab[li - 1]);
          kk[i] = sk ? 1.f : 0.f;
        } else {
          idx[i] = p.gathered ? 0 : p.blank;
        }
      }
    }
    lc.idx0 = idx[0]; lc.idx1 = idx[1]; lc.idx2 = idx[2]; lc.idx3 = idx[3];
    lc.k0 = kk[0]; lc.k1 = kk[1]; lc.k2 = kk[2]; lc.k3 = kk[3];
  }
  {
    const int win_lo_pos = base_w, win_hi_pos = min(base_w + 127, P - 1);
    ss.win_s_lo = SIDE ? (P - 1 - win_hi_pos) : win_lo_pos;
    ss.win_s_hi = SIDE ? (P - 1 - win_lo_pos) : win_hi_pos;
  }
  // ---- initial state: delta on the first lattice state of this side's sweep ----
  ss.st.v0 = ss.st.v1 = ss.st.v2 = ss.st.v3 = 0.f; ss.st.e = kEZero;
  ss.lost = false; ss.inv_mP = 0.f; ss.eP = 0;
  {
    const int q_start = SIDE ? (P - S) : 0;   // backward: 4J - S dummy positions come first
    if (w == 0 && q_start >= pos0 && q_start < pos0 + 4) {
      const int i = q_start - pos0;
      if (i == 0) ss.st.v0 = 1.f; else if (i == 1) ss.st.v1 = 1.f; else if (i == 2) ss.st.v2 = 1.f; else ss.st.v3 = 1.f;
      ss.st.e = 0;
    }
  }

  const int M_side = SIDE ? (T / 2) : (T - T / 2);      // frames this side covers in phase 1
  const int nc1 = (M_side + K - 1) / K;
  const int nc2 = (T - M_side + K - 1) / K;
  const int n_chunks = nc1 + nc2;
  const int bar_id = 1 + SIDE;
  const int n_side_threads = NW * 32;
  auto chunk_n0 = [&](int cc) { return cc < nc1 ? cc * K : M_side + (cc - nc1) * K; };
  auto chunk_kc = [&](int cc) { return cc < nc1 ? min(K, M_side - cc * K) : min(K, T - (M_side + (cc - nc1) * K)); };
  int* abort_flag = cm.abort_flag;
  int* abort_seen = cm.abort_seen;

  // emission rows of the first chunk
  if (n_chunks > 0) stage_rows<K, SIDE>(c, 0, chunk_n0(0), chunk_kc(0));
  cp_async_commit();
  cp_async_wait_all();
  named_bar(bar_id, n_side_threads);

  bool aborted = false;

  // ================================ phase 1 ================================
  for (int cc = 0; cc < nc1 && !aborted; ++cc) {
    if (cc + 1 < n_chunks) stage_rows<K, SIDE>(c, (cc + 1) & 1, chunk_n0(cc + 1), chunk_kc(cc + 1));
    cp_async_commit();
    run_chunk<K, false, SIDE>(c, ss, cc & 1, chunk_n0(cc), chunk_kc(cc), false);
    aborted = chunk_boundary<K, NWMAX, SIDE>(c, ss, cc, abort_flag, abort_seen);
  }

  // ================================ midpoint ================================
  // Both sides always meet here exactly once (even when one of them has already given up).
  if (threadIdx.x == 0) abort_seen[4] = *(volatile int*)abort_flag;
  named_bar(3, 2 * n_side_threads);
  if (abort_seen[4]) aborted = true;
  if (aborted || nc2 == 0) return;

  // the opposite side's groups for the first phase-2 chunk could not be prefetched earlier
  stage_other<K, SIDE>(c, lc, nc1 & 1, chunk_n0(nc1), chunk_kc(nc1));
  cp_async_commit();
  cp_async_wait_all();

  // ---- total probability P = sum_s alpha_t(s) beta'_t(s) at the first phase-2 frame (state copy) ----
  {
    const int buf = nc1 & 1, n0 = chunk_n0(nc1);
    const int NT = blockDim.x >> 1;
    LaneState tmp = ss.st;
    float part = 0.f; int pe = kEZero;
    const int t = c.frame_of(n0);
    const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
    if (!(ss.win_s_hi < lo_t || ss.win_s_lo >= hi_t)) {
      float a0, a1, a2, a3, w0, w1, w2, w3; int E;
      lattice_frame(tmp, lc, c.sm.rows + (size_t)buf * K * c.RWS, lane == 0, a0, a1, a2, a3, w0, w1, w2, w3, E);
      if (lc.owned) {
        const float4 om = c.sm.oth_m[(size_t)buf * K * NT + c.tid_side];
        const int oe = c.sm.oth_e[(size_t)buf * K * NT + c.tid_side];
        const int sA = SIDE ? lc.s_hi : lc.s_lo, d = SIDE ? -1 : 1;
        float sum = 0.f;
        if (sA >= lo_t && sA < hi_t) sum += w0 * om.x;
        if (sA + d >= lo_t && sA + d < hi_t) sum += w1 * om.y;
        if (sA + 2 * d >= lo_t && sA + 2 * d < hi_t) sum += w2 * om.z;
        if (sA + 3 * d >= lo_t && sA + 3 * d < hi_t) sum += w3 * om.w;
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
      if (SIDE == 1 && c.tid_side == 0)
        p.costs[b] = (float)(-((double)Emax + log2((double)tot)) * 0.69314718055994530942);
    }
  }
  if (aborted) return;
  // cost-only calls still walk phase 2 (for the range check) but neither store posteriors nor update rows
  const bool write_post = p.grads != nullptr;

  // ================================ phase 2 ================================
  for (int cc = nc1; cc < n_chunks && !aborted; ++cc) {
    if (cc + 1 < n_chunks) {
      stage_rows<K, SIDE>(c, (cc + 1) & 1, chunk_n0(cc + 1), chunk_kc(cc + 1));
      stage_other<K, SIDE>(c, lc, (cc + 1) & 1, chunk_n0(cc + 1), chunk_kc(cc + 1));
    }
    cp_async_commit();
    run_chunk<K, true, SIDE>(c, ss, cc & 1, chunk_n0(cc), chunk_kc(cc), write_post);
    aborted = chunk_boundary<K, NWMAX, SIDE>(c, ss, cc, abort_flag, abort_seen);
    if (!aborted && write_post) reduce_chunk<K, SIDE>(c, cc & 1, chunk_n0(cc), chunk_kc(cc));
  }
}

// The whole fast path for one utterance; every thread of the CTA calls it.  On return the shared
// word (*smem_abort)[0] is non-zero when the utterance must be redone by the safe lattice (the
// caller reads it after a __syncthreads()).
template <int K, int NWMAX>
__device__ void lattice_fast_utterance(const CallParams& p, int b, unsigned char* smem, int** smem_abort) {
  static_assert(K % 2 == 0 && K >= 2 && K <= 16, "K must be even");
  const UttMeta m = p.meta[b];
  const int L = m.L;
  const int NW = fast_warps_needed<K>(L);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side = warp / NWMAX;
  const int w = warp - side * NWMAX;

  // ---- shared memory: common part, then one block per side ----
  FastCommon cm;
  int* ip = reinterpret_cast<int*>(smem);
  cm.abort_flag = ip;            ip += 1;
  cm.abort_seen = ip;            ip += 7;
  cm.lab = ip;                   ip += L;
  cm.ix.sorted = ip;             ip += L;
  cm.ix.seg_start = ip;          ip += L + 1;
  cm.ix.seg_sym = ip;            ip += L + 1;
  cm.ix.n_seg = ip;              ip += 1;
  size_t common = (size_t)(reinterpret_cast<unsigned char*>(ip) - smem);
  common = (common + 15) / 16 * 16;
  const int RW = p.gathered ? m.W : (p.V + 3) / 4 * 4;
  const size_t side_bytes = fast_side_bytes<K, NWMAX>(L, RW);
  *smem_abort = cm.abort_flag;

  // ---- prologue (all threads of the CTA) ----
  for (int i = threadIdx.x; i < L; i += blockDim.x) cm.lab[i] = p.labels[m.lab_off + i];
  if (threadIdx.x < 8) cm.abort_flag[threadIdx.x] = 0;  // abort_flag + abort_seen[0..6]
  __syncthreads();
  build_symbol_index(cm.lab, L, cm.ix);
  if (w >= NW) return;   // idle warps wait at the caller's __syncthreads()

  if (side == 0) fast_side_sweep<K, NWMAX, 0>(p, b, m, cm, smem + common, w, lane);
  else           fast_side_sweep<K, NWMAX, 1>(p, b, m, cm, smem + common + side_bytes, w, lane);
}

}  // namespace b200ctc
