// Fast lattice: fp64 linear-domain alpha/beta recursion with one power-of-two exponent per warp,
// forward and backward sweeps running concurrently in one CTA, meeting in the middle, and
// RECOMPUTING the half they need from each other instead of streaming it through HBM.
//
// Arithmetic.  The recursion of SURVEY Appendix A is evaluated in the LINEAR domain,
//     alpha_t(s) = y_t(l'_s) * (alpha_{t-1}(s) + alpha_{t-1}(s-1) + [skip] alpha_{t-1}(s-2)),
// on doubles: each lane owns eight consecutive lattice states, all 256 states of a warp share ONE
// int32 exponent that is renormalised once per chunk of K frames.  The inner loop is DADD/DMUL
// only -- no exp/log, no per-state exponent work (the softmax probabilities y come from K1 as
// doubles).  Relative error is ~1e-16 per operation; the gradient error against the fp64 oracle
// is dominated by the fp32 softmax of K1 (~1e-7).  A warp's window spans 2^-1022..2^1023 around
// its largest state; a state that drops out of that range is harmless unless it could carry
// posterior mass -- phase 2 bounds that mass once per chunk and, if the bound is not negligible
// (FLAG_PRECISION_LOST), the utterance is redone by the log-space safe lattice in the same CTA.
//
// Schedule.  One CTA per utterance, warps [0,NW) sweep forward (alpha, t = 0,1,..), warps
// [NWMAX, NWMAX+NW) sweep backward (beta, t = T-1,T-2,..; beta is the same recursion on the
// reversed label sequence).
//   Phase 1: each side advances through its half of the frames and stores only a CHECKPOINT of
//            its window state at every chunk boundary (64 B per lane per K frames).
//   Phase 2: per chunk, each side (1) reloads its checkpoint of the chunk the OTHER side is about
//            to enter and recomputes those K frames, handing the pre-emission values to the other
//            side through shared memory; (2) one CTA barrier; (3) advances its own frontier through
//            K new frames, multiplying with the values the other side just recomputed:
//            posterior(t,s) = alpha_t(s) * beta'_t(s) / P; (4) sums the posteriors per symbol and
//            subtracts the occupancy from the gradient row (filled with the softmax by K1) with one
//            RED per (frame, symbol).
// Sequential depth is T frames, HBM sees only the checkpoints (1/K of the lattice), and the
// 1.5x recursion work is cheaper than the memory traffic it replaces.
//
// Lattice layout.  Lane l of warp w holds positions base_w + 8l .. +7, base_w = w*(256-2K):
// consecutive warp windows overlap by a halo of 2K positions.  Dependencies only point downwards
// (s-1, s-2), so a warp runs K frames without talking to its neighbour while the garbage creeping
// up from its window bottom stays inside the halo; every K frames the warps of a side exchange
// halos through shared memory (one barrier).  Neighbour states inside a warp travel by __shfl_up.
#pragma once

#include "lattice_common.cuh"
#include "lattice_safe.cuh"

namespace b200ctc {

constexpr int kEZero = -(1 << 28);  // exponent of an all-zero window

__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 2^d as a double, d clamped to [-1022 -> 0.0 below, 1023]
__device__ __forceinline__ double pow2d(int d) {
  const int e = min(max(d + 1023, 0), 2046);
  return __hiloint2double(e << 20, 0);
}
// unbiased binary exponent of a positive finite double (garbage for 0: callers test > 0 first)
__device__ __forceinline__ int exponent_of(double x) { return (__double2hiint(x) >> 20) - 1023; }

// ---------------------------------------------------------------------------------------------
// shared memory
// ---------------------------------------------------------------------------------------------
struct DpSideSmem {
  double* xchg;     // [2][K][P]     pre-emission values recomputed for the OTHER side (its position order)
  int* xchg_e;      // [2][NWMAX]    exponent of each of this side's warps for the recomputed chunk
  double* post_lab; // [K][LS]       label-state posteriors of a chunk, SYMBOL-SORTED order
  double* post_blk; // [K][BS]       blank-state posteriors, by blank index
  double* ckpt;     // [2][NT][8]    checkpoint staging, one 64-byte slot per thread
  int* ckpt_e;      // [2][NWMAX]    exponents of the staged checkpoints
  double* halo;     // [2][NWMAX][HL][8]
  int* halo_e;      // [2][NWMAX]
  double* red_m;    // [NWMAX]
  int* red_e;       // [NWMAX]
};

__host__ __device__ inline int dp_lab_stride(int L) { return (L + 1 + 3) / 4 * 4; }  // + dummy slot L
__host__ __device__ inline int dp_blk_stride(int L) { return dp_positions(L) / 2; }

template <int K, int NWMAX>
__host__ __device__ inline size_t dp_side_bytes(int L) {
  const size_t P = dp_positions(L), NT = NWMAX * 32, HL = K / 4;
  size_t b = 0;
  b += 2 * K * P * 8;                                           // xchg
  b += (size_t)K * (dp_lab_stride(L) + dp_blk_stride(L)) * 8;   // post_lab, post_blk
  b += 2 * NT * 64;                                             // ckpt
  b += 2 * NWMAX * HL * 64;                                     // halo
  b += NWMAX * 8;                                               // red_m
  b += (2 * NWMAX + 2 * NWMAX + NWMAX + 2 * NWMAX) * 4;         // xchg_e, halo_e, red_e, ckpt_e
  return (b + 15) / 16 * 16;
}
// rows: [2 buffers][2 chunks][K][WS] doubles shared by both sides
template <int K, int NWMAX>
__host__ __device__ inline size_t dp_smem_bytes(int L, int W) {
  size_t common = (size_t)(8 + 5 * L + 8) * 4;  // control words, lab, sorted, seg_start, seg_sym, rank_of
  common = (common + 15) / 16 * 16;
  const size_t rows = 2 * 2 * K * (size_t)(W + 2) * 8;
  return common + rows + 2 * dp_side_bytes<K, NWMAX>(L) + 16;
}

template <int K, int NWMAX>
__device__ __forceinline__ DpSideSmem carve_dp_side(unsigned char* base, int L) {
  const size_t P = dp_positions(L), NT = NWMAX * 32, HL = K / 4;
  DpSideSmem s;
  unsigned char* p = base;
  s.xchg = reinterpret_cast<double*>(p);     p += 2 * K * P * 8;
  s.post_lab = reinterpret_cast<double*>(p); p += (size_t)K * dp_lab_stride(L) * 8;
  s.post_blk = reinterpret_cast<double*>(p); p += (size_t)K * dp_blk_stride(L) * 8;
  s.ckpt = reinterpret_cast<double*>(p);     p += 2 * NT * 64;
  s.halo = reinterpret_cast<double*>(p);     p += 2 * NWMAX * HL * 64;
  s.red_m = reinterpret_cast<double*>(p);    p += NWMAX * 8;
  s.xchg_e = reinterpret_cast<int*>(p);      p += 2 * NWMAX * 4;
  s.halo_e = reinterpret_cast<int*>(p);      p += 2 * NWMAX * 4;
  s.red_e = reinterpret_cast<int*>(p);       p += NWMAX * 4;
  s.ckpt_e = reinterpret_cast<int*>(p);
  return s;
}

// ---------------------------------------------------------------------------------------------
// per-lane constants and state
// ---------------------------------------------------------------------------------------------
struct DpLane {
  int idx[8];      // index of each state's symbol in the staged emission row (zero slot if invalid)
  int skip;        // bit i: the s-2 -> s transition into slot i is allowed (label slots only)
  int s_lo, s_hi;  // lattice-state range of the lane's eight positions
  bool owned;      // the lane's positions belong to this warp (not to the halo) and exist
  int pos0;        // first position of the lane
  int r[4];        // symbol-sorted slots of the lane's four label states (dummy slot L if none)
  int bi;          // blank index of the lowest of the lane's four blank states (multiple of 4)
  int src_w;       // warp of the OTHER side that owns the mirror image of this lane's positions
};

struct DpState {
  double v[8];
  int ew;          // warp-uniform exponent: true value = v * 2^ew
};

// One frame of the recursion for one lane: acc = pre-emission sums, st.v <- acc * y.
// SIDE 0: even slots are blanks; SIDE 1: odd slots are blanks (blank states never take the skip).
template <int SIDE>
__device__ __forceinline__ void dp_frame(DpState& st, const DpLane& ln, const double* __restrict__ row,
                                         bool lane0, double (&acc)[8]) {
  double n1 = __shfl_up_sync(0xffffffffu, st.v[7], 1);
  double n2 = __shfl_up_sync(0xffffffffu, st.v[6], 1);
  if (lane0) { n1 = 0.0; n2 = 0.0; }   // nothing below the window
  double y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = row[ln.idx[i]];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double m1 = (i >= 1) ? st.v[i - 1] : n1;
    const double m2 = (i >= 2) ? st.v[i - 2] : ((i == 1) ? n1 : n2);
    double a = st.v[i] + m1;
    const bool label_slot = SIDE ? ((i & 1) == 0) : ((i & 1) == 1);
    if (label_slot && (ln.skip & (1 << i))) a += m2;
    acc[i] = a;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) st.v[i] = acc[i] * y[i];
}

template <int SIDE>
struct DpCtx {
  const CallParams* p;
  int b, T, L, S, P, NW, W, WS, LS, BS;
  int w, lane, tid_side, nt_side;
  DpSideSmem sm;       // this side
  DpSideSmem other;    // the opposite side (xchg / xchg_e are read from there)
  double* rows;        // [2][2][K][WS] shared emission rows: [buffer][chunk slot: 0 = fwd side's frames, 1 = bwd side's]
  SymbolIndex ix;
  double* ck_v;        // checkpoints of this side: [chunk][NW][32][8] doubles
  int* ck_e;           // [chunk][NW]
  const double* em;    // emission rows of this utterance: [T][W] doubles
  __device__ __forceinline__ int frame_of(int n) const { return SIDE ? T - 1 - n : n; }
};

// chunk grid of one side, counted in that side's own step order n = 0..T-1.
// Phase 1 covers steps [0, M): a first chunk of (M mod K) steps (if non-zero), then full chunks.
// Phase 2 covers steps [M, T): full chunks, the last one partial.  This makes the phase-2 chunks
// of one side coincide with the phase-1 chunks of the other side.
struct DpGrid {
  int M, r1, nc1, nc2;
  __device__ __forceinline__ int n0(int cc, int K) const {
    if (cc < nc1) return (cc == 0) ? 0 : (r1 ? r1 + (cc - 1) * K : cc * K);
    return M + (cc - nc1) * K;
  }
  __device__ __forceinline__ int kc(int cc, int K, int T) const {
    if (cc < nc1) return (cc == 0 && r1) ? r1 : K;
    return min(K, T - (M + (cc - nc1) * K));
  }
};
__device__ __forceinline__ DpGrid make_grid(int M, int T, int K) {
  DpGrid g;
  g.M = M; g.r1 = M % K;
  g.nc1 = (M + K - 1) / K;
  g.nc2 = (T - M + K - 1) / K;
  return g;
}

// Stage kc emission rows starting at frame t_first (ascending frames) into a row slot; executed by
// the nthreads threads tid = 0.. of the caller's choosing.
template <int K>
__device__ __forceinline__ void dp_stage_rows(double* dst, const double* em, int W, int WS, int t_first, int kc,
                                              int tid, int nthreads) {
  const int per_row = W / 2;   // 16-byte pieces per row
  for (int j = 0; j < kc; ++j) {
    const double* src = em + (long long)(t_first + j) * W;
    double* d = dst + (size_t)j * WS;
    for (int e = tid; e < per_row; e += nthreads) cp_async_16(d + 2 * e, src + 2 * e);
  }
}

struct FastCommon {
  int* abort_flag;   // set by any thread: leave the fast path
  int* abort_seen;   // [0..3] per side/parity snapshots, [4] midpoint, [5] per-iteration CTA snapshot
  int* lab;
  SymbolIndex ix;
};

// Chunk boundary of an advancing side: drop dead states, renormalise the warp exponent, exchange
// halos (one side barrier).  `t_next` is the next frame this side will process.
template <int K, int NWMAX, int SIDE>
__device__ __forceinline__ void dp_boundary(const DpCtx<SIDE>& c, DpState& st, const DpLane& ln, int cc, int t_next) {
  constexpr int HL = K / 4;
  const int hb = cc & 1, NW = c.NW, w = c.w, lane = c.lane;
  // states that can no longer reach the end (forward) / were never reachable from the start
  // (backward) within the remaining frames are dead for good: clear them so that they neither
  // set the exponent nor keep garbage alive.
  {
    const int lo_t = max(0, c.S - 2 * (c.T - t_next) - 2), hi_t = min(c.S, 2 * (t_next + 1) + 2);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s = SIDE ? ln.s_hi - i : ln.s_lo + i;
      const bool dead = SIDE ? (s >= hi_t) : (s < lo_t);
      if (dead) st.v[i] = 0.0;
    }
  }
  // renormalise: warp maximum -> [1,2)
  double mx = st.v[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmax(mx, st.v[i]);
  int hi = __double2hiint(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  if (hi > 0) {   // some state is positive (positive doubles order like their high words)
    const int eb = (hi >> 20) - 1023;
    const double sc = pow2d(-eb);
#pragma unroll
    for (int i = 0; i < 8; ++i) st.v[i] *= sc;
    st.ew += eb;
  } else {
    st.ew = kEZero;
  }
  // halo exchange
  if (w + 1 < NW && lane >= 32 - HL) {
    double* d = c.sm.halo + ((size_t)(hb * NWMAX + w) * HL + (lane - (32 - HL))) * 8;
#pragma unroll
    for (int i = 0; i < 8; i += 2) *reinterpret_cast<double2*>(d + i) = make_double2(st.v[i], st.v[i + 1]);
  }
  if (lane == 0) c.sm.halo_e[hb * NWMAX + w] = st.ew;
  named_bar(1 + SIDE, NW * 32);
  if (w > 0) {
    const int en = c.sm.halo_e[hb * NWMAX + w - 1];
    // common exponent of the window: the larger of the two
    const int E = max(st.ew, en);
    const double so = pow2d(st.ew - E), sn = pow2d(en - E);
    if (lane < HL) {
      const double* d = c.sm.halo + ((size_t)(hb * NWMAX + w - 1) * HL + lane) * 8;
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 h = *reinterpret_cast<const double2*>(d + i);
        st.v[i] = h.x * sn; st.v[i + 1] = h.y * sn;
      }
    } else if (E != st.ew) {
#pragma unroll
      for (int i = 0; i < 8; ++i) st.v[i] *= so;
    }
    st.ew = E;
  }
}

// Store / load the checkpoint of a window (all lanes, halo included).
template <int SIDE>
__device__ __forceinline__ void dp_store_ckpt(const DpCtx<SIDE>& c, const DpState& st, int chunk) {
  double* d = c.ck_v + (((size_t)chunk * c.NW + c.w) * 32 + c.lane) * 8;
#pragma unroll
  for (int i = 0; i < 8; i += 2) *reinterpret_cast<double2*>(d + i) = make_double2(st.v[i], st.v[i + 1]);
  if (c.lane == 0) c.ck_e[chunk * c.NW + c.w] = st.ew;
}
template <int SIDE>
__device__ __forceinline__ void dp_prefetch_ckpt(const DpCtx<SIDE>& c, int buf, int chunk) {
  const int NT = blockDim.x >> 1;
  const double* s = c.ck_v + (((size_t)chunk * c.NW + c.w) * 32 + c.lane) * 8;
  double* d = c.sm.ckpt + ((size_t)buf * NT + c.tid_side) * 8;
#pragma unroll
  for (int i = 0; i < 8; i += 2) cp_async_16(d + i, s + i);
  if (c.lane == 0) cp_async_4(c.sm.ckpt_e + buf * (NT >> 5) + c.w, c.ck_e + chunk * c.NW + c.w);
}

// Occupancy update for the frames of one finished phase-2 chunk: one deterministic sum and one RED
// per (frame, symbol).  Label posteriors lie in symbol-sorted order: every symbol is a contiguous run.
template <int K, int SIDE>
__device__ __forceinline__ void dp_reduce_chunk(const DpCtx<SIDE>& c, int n0, int kc) {
  const CallParams& p = *c.p;
  const int NW = c.NW;
  const int n_seg = *c.ix.n_seg;
  const int n_slices = 1 + (n_seg + 31) / 32;   // slice 0: blank (whole warp); others: 32 symbols each
  for (int item = c.w; item < kc * n_slices; item += NW) {
    const int j = item / n_slices, slice = item - j * n_slices;
    float* grow = p.grads + ((long long)c.frame_of(n0 + j) * p.B + c.b) * p.V;
    if (slice == 0) {
      const double2* row2 = reinterpret_cast<const double2*>(c.sm.post_blk + (size_t)j * c.BS);
      double acc = 0.0;
      for (int g = c.lane; g < (c.BS >> 1); g += 32) {
        const double2 q = row2[g];
        acc += q.x + q.y;
      }
      acc = warp_sum(acc);
      if (c.lane == 0) atomicAdd(grow + p.blank, -(float)acc);
    } else {
      const int u = (slice - 1) * 32 + c.lane;
      if (u < n_seg) {
        const double* row = c.sm.post_lab + (size_t)j * c.LS;
        const int k1 = c.ix.seg_start[u + 1];
        int k = c.ix.seg_start[u];
        double acc = 0.0;
        for (; k + 4 <= k1; k += 4) {
          const double x0 = row[k], x1 = row[k + 1], x2 = row[k + 2], x3 = row[k + 3];
          acc += (x0 + x1) + (x2 + x3);
        }
        for (; k < k1; ++k) acc += row[k];
        atomicAdd(grow + c.ix.seg_sym[u], -(float)acc);
      }
    }
  }
}

// One side's sweep (all warps w < NW of that side).
template <int K, int NWMAX, int SIDE>
__device__ void dp_side_sweep(const CallParams& p, int b, const UttMeta& m, const FastCommon& cm,
                              double* rows, unsigned char* side_smem, unsigned char* other_smem, int w, int lane) {
  constexpr int H = 2 * K;          // halo positions
  constexpr int HL = K / 4;         // halo lanes
  constexpr int OWN = 256 - H;
  static_assert(K % 4 == 0 && K >= 4 && K <= 16, "K must be a multiple of 4");

  DpCtx<SIDE> c;
  c.p = &p; c.b = b;
  const int T = m.T, L = m.L;
  c.T = T; c.L = L; c.S = 2 * L + 1; c.P = dp_positions(L);
  c.NW = dp_warps_needed<K>(L);
  c.W = m.W; c.WS = m.W + 2;
  c.LS = dp_lab_stride(L); c.BS = dp_blk_stride(L);
  c.w = w; c.lane = lane; c.tid_side = w * 32 + lane; c.nt_side = c.NW * 32;
  c.ix = cm.ix;
  c.rows = rows;
  const int P = c.P, S = c.S, NW = c.NW;
  c.sm = carve_dp_side<K, NWMAX>(side_smem, L);
  c.other = carve_dp_side<K, NWMAX>(other_smem, L);
  const int* lab = cm.lab;
  c.em = p.em + m.em_off;

  const DpGrid grid = make_grid(SIDE ? (T / 2) : (T - T / 2), T, K);
  const DpGrid ogrid = make_grid(SIDE ? (T - T / 2) : (T / 2), T, K);   // the other side's grid
  const int nc1 = grid.nc1, nc2 = grid.nc2;

  // checkpoints of this side in the scratch: first all forward-side chunks, then the backward side's
  {
    unsigned char* scr = p.scratch + m.scratch_off * kGroupBytes;
    const size_t ck_bytes_fwd = (size_t)(SIDE ? ogrid.nc1 : grid.nc1) * NW * 32 * 64;
    const size_t ck_bytes_bwd = (size_t)(SIDE ? grid.nc1 : ogrid.nc1) * NW * 32 * 64;
    unsigned char* base_v = scr + (SIDE ? ck_bytes_fwd : 0);
    c.ck_v = reinterpret_cast<double*>(base_v);
    int* e_base = reinterpret_cast<int*>(scr + ck_bytes_fwd + ck_bytes_bwd);
    c.ck_e = e_base + (SIDE ? (SIDE ? ogrid.nc1 : 0) * NW : 0);
  }

  // ---- side prologue: zero slots, padding ----
  for (int i = c.tid_side; i < K * c.BS; i += c.nt_side) c.sm.post_blk[i] = 0.0;
  for (int i = c.tid_side; i < 2 * K; i += c.nt_side) {   // the zero slot of every row this side stages
    const int buf = i / K, j = i - buf * K;
    double* r = c.rows + (size_t)((buf * 2 + SIDE) * K + j) * c.WS;
    r[c.W] = 0.0; r[c.W + 1] = 0.0;
  }

  // ---- per-lane constants ----
  DpLane ln;
  const int base_w = w * OWN;
  ln.pos0 = base_w + 8 * lane;
  ln.owned = ((w == 0) || (lane >= HL)) && (ln.pos0 < P);
  ln.s_lo = SIDE ? (P - 1 - ln.pos0 - 7) : ln.pos0;
  ln.s_hi = ln.s_lo + 7;
  ln.skip = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = ln.pos0 + i;
    const int s = SIDE ? (P - 1 - q) : q;
    const bool ok = (q < P) && (s >= 0) && (s < S);
    int idx = c.W;   // zero slot
    if (ok) {
      const int li = s >> 1;
      if (s & 1) {
        idx = p.gathered ? li + 1 : lab[li];
        const bool sk = SIDE ? (s + 2 < S && lab[li] != lab[li + 1]) : (s >= 3 && lab[li] != lab[li - 1]);
        if (sk) ln.skip |= 1 << i;
      } else {
        idx = p.gathered ? 0 : p.blank;
      }
    }
    ln.idx[i] = idx;
  }
  {
    // label states: forward slots 1,3,5,7 (labels la..la+3); backward slots 0,2,4,6 (labels la, la-1, ..)
    const int la = SIDE ? ((ln.s_hi - 1) >> 1) : (ln.s_lo >> 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int li = SIDE ? la - k : la + k;
      ln.r[k] = (ln.owned && li >= 0 && li < L) ? cm.ix.rank_of[li] : L;
    }
    ln.bi = ln.s_lo >> 1;   // blank indices bi..bi+3 (s_lo is a multiple of 8)
    // the other side's warp that owns the mirror positions P-8-pos0 .. P-1-pos0
    const int mq = P - 8 - ln.pos0;
    ln.src_w = (mq < 256) ? 0 : (mq - 2 * K) / OWN;   // owned ranges: w0 [0,256), w>=1 [w*OWN+2K, (w+1)*OWN+2K)
  }
  int win_s_lo, win_s_hi;
  {
    const int win_lo_pos = base_w, win_hi_pos = min(base_w + 255, P - 1);
    win_s_lo = SIDE ? (P - 1 - win_hi_pos) : win_lo_pos;
    win_s_hi = SIDE ? (P - 1 - win_lo_pos) : win_hi_pos;
  }
  // ---- initial state: delta on the first lattice state of this side's sweep ----
  DpState st;
#pragma unroll
  for (int i = 0; i < 8; ++i) st.v[i] = 0.0;
  st.ew = kEZero;
  {
    const int q_start = SIDE ? (P - S) : 0;   // backward: P - S dummy positions come first
    if (w == 0) {
      st.ew = 0;
      if (q_start >= ln.pos0 && q_start < ln.pos0 + 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (ln.pos0 + i == q_start) st.v[i] = 1.0;
      }
    }
  }

  const int bar_id = 1 + SIDE;
  const int n_side_threads = NW * 32;
  const bool lane0 = lane == 0;
  int* abort_flag = cm.abort_flag;
  int* abort_seen = cm.abort_seen;
  double acc[8];

  // ================================ phase 1 ================================
  // rows for this side's chunk cc are staged into rows[cc & 1][SIDE] one chunk ahead
  if (nc1 > 0) {
    const int n0 = grid.n0(0, K), kc = grid.kc(0, K, T);
    const int t_first = SIDE ? c.frame_of(n0 + kc - 1) : c.frame_of(n0);
    dp_stage_rows<K>(c.rows + (size_t)((0 * 2 + SIDE) * K) * c.WS, c.em, c.W, c.WS, t_first, kc, c.tid_side, c.nt_side);
  }
  cp_async_commit();
  cp_async_wait_all();
  named_bar(bar_id, n_side_threads);

  for (int cc = 0; cc < nc1; ++cc) {
    const int n0 = grid.n0(cc, K), kc = grid.kc(cc, K, T);
    if (cc + 1 < nc1) {
      const int n0n = grid.n0(cc + 1, K), kcn = grid.kc(cc + 1, K, T);
      const int t_first = SIDE ? c.frame_of(n0n + kcn - 1) : c.frame_of(n0n);
      dp_stage_rows<K>(c.rows + (size_t)((((cc + 1) & 1) * 2 + SIDE) * K) * c.WS, c.em, c.W, c.WS, t_first, kcn,
                       c.tid_side, c.nt_side);
    }
    cp_async_commit();
    dp_store_ckpt<SIDE>(c, st, cc);   // state BEFORE the chunk: what the recomputation starts from
    const double* rbase = c.rows + (size_t)(((cc & 1) * 2 + SIDE) * K) * c.WS;
    int t = c.frame_of(n0);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (j < kc) {
        const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
        if (!(win_s_hi < lo_t || win_s_lo >= hi_t)) {   // warp-uniform band test
          const int rj = SIDE ? (kc - 1 - j) : j;        // rows are staged in ascending frame order
          dp_frame<SIDE>(st, ln, rbase + (size_t)rj * c.WS, lane0, acc);
        }
        t += SIDE ? -1 : 1;
      }
    }
    cp_async_wait_all();
    dp_boundary<K, NWMAX, SIDE>(c, st, ln, cc, t);
  }

  // ================================ midpoint ================================
  __threadfence_block();
  named_bar(3, 2 * n_side_threads);

  // ================================ phase 2 ================================
  // Iteration i: this side advances through its phase-2 chunk i (steps grid.M + i*K ..) and
  // recomputes its own phase-1 chunk that covers the frames of the OTHER side's phase-2 chunk i,
  // i.e. its phase-1 chunk (nc1 - 1 - i).  Row slots per iteration buffer: [0] forward side's
  // advance frames, [1] backward side's advance frames (each side's recompute uses the other slot).
  const int n_iter = max(nc2, ogrid.nc2);
  double inv_mP = 0.0; int eP = 0;
  bool have_P = false, lost = false;
  const bool write_post = p.grads != nullptr;

  auto stage_iter = [&](int it, int buf) {
    // this side stages the rows of ITS OWN advance chunk `it` (both sides do the same for theirs)
    if (it < nc2) {
      const int n0 = grid.n0(nc1 + it, K), kc = grid.kc(nc1 + it, K, T);
      const int t_first = SIDE ? c.frame_of(n0 + kc - 1) : c.frame_of(n0);
      dp_stage_rows<K>(c.rows + (size_t)((buf * 2 + SIDE) * K) * c.WS, c.em, c.W, c.WS, t_first, kc, c.tid_side,
                       c.nt_side);
    }
    // and prefetches its checkpoint for the recomputation of iteration `it`
    if (it < ogrid.nc2 && nc1 - 1 - it >= 0) dp_prefetch_ckpt<SIDE>(c, buf, nc1 - 1 - it);
  };
  stage_iter(0, 0);
  cp_async_commit();

  for (int it = 0; it < n_iter; ++it) {
    const int buf = it & 1;
    cp_async_wait_all();
    named_bar(3, 2 * n_side_threads);          // rows + checkpoints of this iteration are in place (both sides)
    if (it + 1 < n_iter) stage_iter(it + 1, buf ^ 1);
    cp_async_commit();

    // ---- (1) recompute own phase-1 chunk rc for the other side ----
    const int rc = nc1 - 1 - it;
    if (it < ogrid.nc2 && rc >= 0) {
      const int NT = blockDim.x >> 1;
      DpState rs;
      {
        const double* s = c.sm.ckpt + ((size_t)buf * NT + c.tid_side) * 8;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const double2 h = *reinterpret_cast<const double2*>(s + i);
          rs.v[i] = h.x; rs.v[i + 1] = h.y;
        }
        rs.ew = c.sm.ckpt_e[buf * (NT >> 5) + w];
      }
      const int n0 = grid.n0(rc, K), kc = grid.kc(rc, K, T);
      // the frames of this chunk are the other side's advance frames: row slot [1 - SIDE]
      const double* rbase = c.rows + (size_t)((buf * 2 + (1 - SIDE)) * K) * c.WS;
      double* xbase = c.sm.xchg + (size_t)buf * K * P;
      int t = c.frame_of(n0);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j < kc) {
          const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
          const bool in_band = !(win_s_hi < lo_t || win_s_lo >= hi_t);
          // other side's row order: its slot stages ascending frames; our frame t sits at
          // index (t - first frame of that slot)
          const int t_slot_first = SIDE ? c.frame_of(n0) - (kc - 1) : c.frame_of(n0);
          const int rj = SIDE ? (t - t_slot_first) : (t - t_slot_first);
          if (in_band) dp_frame<SIDE>(rs, ln, rbase + (size_t)rj * c.WS, lane0, acc);
          if (ln.owned) {
            // frame slot by the OTHER side's step order within its chunk: it walks these frames in
            // the opposite direction, so our j-th frame is its (kc-1-j)-th
            double* xr = xbase + (size_t)(kc - 1 - j) * P + (P - 8 - ln.pos0);   // mirrored, reversed
            if (in_band) {
              *reinterpret_cast<double2*>(xr + 0) = make_double2(acc[7], acc[6]);
              *reinterpret_cast<double2*>(xr + 2) = make_double2(acc[5], acc[4]);
              *reinterpret_cast<double2*>(xr + 4) = make_double2(acc[3], acc[2]);
              *reinterpret_cast<double2*>(xr + 6) = make_double2(acc[1], acc[0]);
            } else {
              const double2 z = make_double2(0.0, 0.0);
              *reinterpret_cast<double2*>(xr + 0) = z; *reinterpret_cast<double2*>(xr + 2) = z;
              *reinterpret_cast<double2*>(xr + 4) = z; *reinterpret_cast<double2*>(xr + 6) = z;
            }
          }
          t += SIDE ? -1 : 1;
        }
      }
      if (lane == 0) c.sm.xchg_e[buf * NWMAX + w] = rs.ew;
    }
    if (threadIdx.x == 0) abort_seen[5] = *(volatile int*)abort_flag;

    // ---- (2) both sides have published their recomputed chunk ----
    named_bar(3, 2 * n_side_threads);
    if (abort_seen[5]) return;

    // ---- (3) advance own frontier through phase-2 chunk `it` ----
    if (it < nc2) {
      const int cc = nc1 + it;
      const int n0 = grid.n0(cc, K), kc = grid.kc(cc, K, T);
      const double* rbase = c.rows + (size_t)((buf * 2 + SIDE) * K) * c.WS;
      const double* xbase = c.other.xchg + (size_t)buf * K * P + ln.pos0;
      const int oe = c.other.xchg_e[buf * NWMAX + ln.src_w];

      if (!have_P) {
        // ---- total probability P = sum_s alpha_t(s) beta'_t(s) at the first phase-2 frame ----
        DpState tmp = st;
        const int t = c.frame_of(n0);
        const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
        double part = 0.0; int pe = kEZero;
        if (!(win_s_hi < lo_t || win_s_lo >= hi_t)) {
          const int rj = SIDE ? (kc - 1) : 0;
          dp_frame<SIDE>(tmp, ln, rbase + (size_t)rj * c.WS, lane0, acc);
          if (ln.owned) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int s = SIDE ? ln.s_hi - i : ln.s_lo + i;
              if (s >= lo_t && s < hi_t) sum += tmp.v[i] * xbase[i];
            }
            if (sum > 0.0) { part = sum; pe = st.ew + oe; }
          }
        }
        int emax = pe;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
        double scaled = part * pow2d(pe - emax);
        scaled = warp_sum(scaled);
        if (lane == 0) { c.sm.red_m[w] = scaled; c.sm.red_e[w] = emax; }
        named_bar(bar_id, n_side_threads);
        int Emax = kEZero;
        for (int i = 0; i < NW; ++i) Emax = max(Emax, c.sm.red_e[i]);
        double tot = 0.0;
        for (int i = 0; i < NW; ++i) tot += c.sm.red_m[i] * pow2d(c.sm.red_e[i] - Emax);
        if (!(tot > 0.0) || !(tot < INFINITY) || Emax <= kEZero / 2) {
          if (c.tid_side == 0) *abort_flag = 1;   // zero / garbage probability: the safe lattice decides
          lost = true;
          inv_mP = 0.0; eP = 0;
        } else {
          const int eb = exponent_of(tot);
          const double mP = tot * pow2d(-eb);   // [1,2)
          inv_mP = 1.0 / mP; eP = Emax + eb;
          if (SIDE == 1 && c.tid_side == 0)
            p.costs[b] = (float)(-((double)eP + log2(mP)) * 0.69314718055994530942);
        }
        have_P = true;
      }

      // posterior scale of this chunk: 2^(ew_own + ew_src - eP) / mP  (ew is constant inside a chunk)
      const double scale = pow2d(st.ew + oe - eP) * inv_mP;
      int t = c.frame_of(n0);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j < kc) {
          const int hi_t = min(S, 2 * (t + 1)), lo_t = max(0, S - 2 * (T - t));
          const bool in_band = !(win_s_hi < lo_t || win_s_lo >= hi_t);
          const int rj = SIDE ? (kc - 1 - j) : j;
          if (in_band) dp_frame<SIDE>(st, ln, rbase + (size_t)rj * c.WS, lane0, acc);
          if (ln.owned && (write_post || j == 0)) {
            double po[8];
            if (in_band) {
              const double* xr = xbase + (size_t)j * P;
              double om[8];
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                const double2 h = *reinterpret_cast<const double2*>(xr + i);
                om[i] = h.x; om[i + 1] = h.y;
              }
              const bool all_in = ln.s_lo >= lo_t && ln.s_hi < hi_t;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int s = SIDE ? ln.s_hi - i : ln.s_lo + i;
                const bool inb = all_in || (s >= lo_t && s < hi_t);
                po[i] = inb ? (st.v[i] * om[i]) * scale : 0.0;
              }
              if (j == 0) {
                // Range check, once per chunk.  A state that fell out of its window's range (taken
                // as 2^-900 below the window's largest state, which is < 2^8 in units of 2^ew) may
                // have been flushed on either side.  Its posterior is bounded by
                //   2^(8-900) * max(own lane, other lane) * 2^(ew_own + ew_other) / P ;
                // if that bound is not negligible (> 2^-24) the result cannot be trusted.
                double um = st.v[0], omx = om[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) { um = fmax(um, st.v[i]); omx = fmax(omx, om[i]); }
                const double mm = fmax(um, omx);
                if (mm > 0.0) lost |= exponent_of(mm) + st.ew + oe - eP > 900 - 24 - 8;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) po[i] = 0.0;
            }
            if (write_post) {
              double* pl = c.sm.post_lab + (size_t)j * c.LS;
              double* pb = c.sm.post_blk + (size_t)j * c.BS + ln.bi;
              if (SIDE == 0) {
                pl[ln.r[0]] = po[1]; pl[ln.r[1]] = po[3]; pl[ln.r[2]] = po[5]; pl[ln.r[3]] = po[7];
                *reinterpret_cast<double2*>(pb + 0) = make_double2(po[0], po[2]);
                *reinterpret_cast<double2*>(pb + 2) = make_double2(po[4], po[6]);
              } else {
                pl[ln.r[0]] = po[0]; pl[ln.r[1]] = po[2]; pl[ln.r[2]] = po[4]; pl[ln.r[3]] = po[6];
                *reinterpret_cast<double2*>(pb + 0) = make_double2(po[7], po[5]);
                *reinterpret_cast<double2*>(pb + 2) = make_double2(po[3], po[1]);
              }
            }
          }
          t += SIDE ? -1 : 1;
        }
      }
      if (__any_sync(0xffffffffu, lost) && lane == 0) *abort_flag = 1;
      // ---- (4) side boundary (halo exchange, renormalisation), then the occupancy update ----
      dp_boundary<K, NWMAX, SIDE>(c, st, ln, cc, t);
      if (write_post && !lost) dp_reduce_chunk<K, SIDE>(c, n0, kc);
    }
  }
}

// The whole fast path for one utterance; every thread of the CTA calls it.  On return the shared
// word (*smem_abort)[0] is non-zero when the utterance must be redone by the safe lattice (the
// caller reads it after a __syncthreads()).
template <int K, int NWMAX>
__device__ void lattice_dp_utterance(const CallParams& p, int b, unsigned char* smem, int** smem_abort) {
  const UttMeta m = p.meta[b];
  const int L = m.L;
  const int NW = dp_warps_needed<K>(L);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side = warp / NWMAX;
  const int w = warp - side * NWMAX;

  // ---- shared memory: common part, shared emission rows, then one block per side ----
  FastCommon cm;
  int* ip = reinterpret_cast<int*>(smem);
  cm.abort_flag = ip;            ip += 1;
  cm.abort_seen = ip;            ip += 7;
  cm.lab = ip;                   ip += L;
  cm.ix.sorted = ip;             ip += L;
  cm.ix.seg_start = ip;          ip += L + 1;
  cm.ix.seg_sym = ip;            ip += L + 1;
  cm.ix.n_seg = ip;              ip += 1;
  cm.ix.rank_of = ip;            ip += L;
  size_t common = (size_t)(reinterpret_cast<unsigned char*>(ip) - smem);
  common = (common + 15) / 16 * 16;
  double* rows = reinterpret_cast<double*>(smem + common);
  const size_t rows_bytes = 2 * 2 * K * (size_t)(m.W + 2) * 8;
  const size_t side_bytes = dp_side_bytes<K, NWMAX>(L);
  unsigned char* side0 = smem + common + rows_bytes;
  unsigned char* side1 = side0 + side_bytes;
  *smem_abort = cm.abort_flag;

  // ---- prologue (all threads of the CTA) ----
  for (int i = threadIdx.x; i < L; i += blockDim.x) cm.lab[i] = p.labels[m.lab_off + i];
  if (threadIdx.x < 8) cm.abort_flag[threadIdx.x] = 0;  // abort_flag + abort_seen[0..6]
  __syncthreads();
  build_symbol_index(cm.lab, L, cm.ix);
  if (w >= NW) return;   // idle warps wait at the caller's __syncthreads()

  if (side == 0) dp_side_sweep<K, NWMAX, 0>(p, b, m, cm, rows, side0, side1, w, lane);
  else           dp_side_sweep<K, NWMAX, 1>(p, b, m, cm, rows, side1, side0, w, lane);
}

}  // namespace b200ctc
