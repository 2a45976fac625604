// Pieces shared by the two lattice kernels (block-exponent fast path and fp64 safe path).
#pragma once

#include "common.cuh"

namespace b200ctc {

// Symbol index: the occupancy of symbol k at frame t is the sum of the posteriors of all lattice
// states that carry k.  Blank sits on the even states; the label states (odd, s = 2i+1) are
// grouped by symbol here so that every (frame, symbol) pair receives ONE deterministic sum and
// one update of the gradient row (chainer ctc :150-157 does the same grouping with a python set).
//   sorted[0..L)      label positions i ordered by (symbol, position)
//   seg_start[0..n]   segment boundaries into sorted[]
//   seg_sym[0..n)     the symbol of each segment
struct SymbolIndex {
  int* sorted;
  int* seg_start;
  int* seg_sym;
  int* n_seg;  // single int in shared memory
};

// All threads of the CTA call this; lab[] must already be in shared memory and visible.
// Ends with a __syncthreads().  Everything is data-parallel: a rank sort on packed (symbol, position)
// keys (ix.seg_sym doubles as the key array until the segments are written), then the segment
// boundaries by a ballot scan over the sorted order.
constexpr int kCountSortMaxV = 64;     // counting-sort path: vocabularies up to 64 symbols ...
constexpr int kCountSortMaxBlk = 16;   // ... and label sequences up to 512 (one thread per label)

__device__ __forceinline__ void build_symbol_index(const int* lab, int L, int V, SymbolIndex ix) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, n_warps = nt >> 5;
  __shared__ int s_warp_tot[32];
  __shared__ int s_base;
  if (V <= kCountSortMaxV && L <= nt && L <= 32 * kCountSortMaxBlk) {
    // Small vocabularies: a stable counting sort, O(L) work.  Thread i owns label i; a warp is a block of 32
    // labels.  s_cnt[blk][sym] first holds the block's count of the symbol, then the number of earlier labels
    // with that symbol in earlier blocks; match_any gives the rank inside the block.
    __shared__ int s_cnt[kCountSortMaxBlk * kCountSortMaxV];
    __shared__ int s_sym_base[kCountSortMaxV];
    const int n_blk = (L + 31) >> 5;
    for (int i = tid; i < n_blk * kCountSortMaxV; i += nt) s_cnt[i] = 0;
    __syncthreads();
    int sym = -1, r = 0;
    if (warp < n_blk) {                                   // whole warps: lanes past L carry the sentinel -1
      sym = tid < L ? lab[tid] : -1;
      const unsigned m = __match_any_sync(0xffffffffu, sym);
      r = __popc(m & ((1u << lane) - 1u));
      if (sym >= 0 && r == 0) s_cnt[warp * kCountSortMaxV + sym] = __popc(m);
    }
    __syncthreads();
    if (tid < kCountSortMaxV) {                           // per symbol: exclusive prefix over the blocks, total
      int acc = 0;
      for (int b = 0; b < n_blk; ++b) {
        const int c = s_cnt[b * kCountSortMaxV + tid];
        s_cnt[b * kCountSortMaxV + tid] = acc;
        acc += c;
      }
      s_sym_base[tid] = acc;                              // the symbol's count, for now
    }
    __syncthreads();
    if (warp == 0) {                                      // exclusive scan over the symbols; segments = symbols present
      int base = 0, nseg = 0;
#pragma unroll
      for (int h = 0; h < kCountSortMaxV / 32; ++h) {
        const int t = 32 * h + lane;
        const int c = s_sym_base[t];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const unsigned present = __ballot_sync(0xffffffffu, c > 0);
        const int start = base + incl - c;
        s_sym_base[t] = start;
        if (c > 0) {
          const int u = nseg + __popc(present & ((1u << lane) - 1u));
          ix.seg_start[u] = start;
          ix.seg_sym[u] = t;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
        nseg += __popc(present);
      }
      if (lane == 0) { ix.seg_start[nseg] = L; *ix.n_seg = nseg; }
    }
    __syncthreads();
    if (sym >= 0) ix.sorted[s_sym_base[sym] + s_cnt[warp * kCountSortMaxV + sym] + r] = tid;
    __syncthreads();
    return;
  }
  if ((long long)V * L < (1ll << 31)) {
    // rank sort on one 32-bit key per label: position i goes to slot #{j : key[j] < key[i]}
    int* key = ix.seg_sym;
    for (int i = tid; i < L; i += nt) key[i] = lab[i] * L + i;
    __syncthreads();
    const int head = min(L, (int)((16u - ((unsigned)__cvta_generic_to_shared(key) & 15u)) & 15u) >> 2);
    for (int i = tid; i < L; i += nt) {
      const int ki = key[i];
      int rank = 0, j = 0;
      for (; j < head; ++j) rank += key[j] < ki;
      for (; j + 4 <= L; j += 4) {
        const int4 q = *reinterpret_cast<const int4*>(key + j);
        rank += (q.x < ki) + (q.y < ki) + (q.z < ki) + (q.w < ki);
      }
      for (; j < L; ++j) rank += key[j] < ki;
      ix.sorted[rank] = i;
    }
  } else {
    for (int i = tid; i < L; i += nt) {
      const int li = lab[i];
      int rank = 0;
      for (int j = 0; j < L; ++j) {
        const int lj = lab[j];
        rank += (lj < li) || (lj == li && j < i);
      }
      ix.sorted[rank] = i;
    }
  }
  if (tid == 0) s_base = 0;
  __syncthreads();
  // segment starts in sorted order, numbered by a CTA-wide ballot scan, one tile of blockDim.x positions at a time
  for (int k0 = 0; k0 < L; k0 += nt) {
    const int k = k0 + tid;
    int sym = 0;
    bool start = false;
    if (k < L) {
      sym = lab[ix.sorted[k]];
      start = k == 0 || sym != lab[ix.sorted[k - 1]];
    }
    const unsigned bal = __ballot_sync(0xffffffffu, start);
    if (lane == 0) s_warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int i = 0; i < warp; ++i) off += s_warp_tot[i];
    if (start) {
      const int u = off + __popc(bal & ((1u << lane) - 1u));
      ix.seg_start[u] = k;
      ix.seg_sym[u] = sym;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = s_base;
      for (int i = 0; i < n_warps; ++i) tot += s_warp_tot[i];
      s_base = tot;
    }
    __syncthreads();
  }
  if (tid == 0) {
    ix.seg_start[s_base] = L;
    *ix.n_seg = s_base;
  }
  __syncthreads();
}

}  // namespace b200ctc
