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
// Ends with a __syncthreads().
__device__ __forceinline__ void build_symbol_index(const int* lab, int L, SymbolIndex ix) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // rank sort: position i goes to slot #{j : (lab[j], j) < (lab[i], i)}
  for (int i = tid; i < L; i += nt) {
    const int li = lab[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const int lj = lab[j];
      rank += (lj < li) || (lj == li && j < i);
    }
    ix.sorted[rank] = i;
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int k = 0; k < L; ++k) {
      const int sym = lab[ix.sorted[k]];
      if (k == 0 || sym != lab[ix.sorted[k - 1]]) {
        ix.seg_start[n] = k;
        ix.seg_sym[n] = sym;
        ++n;
      }
    }
    ix.seg_start[n] = L;
    *ix.n_seg = n;
  }
  __syncthreads();
}

}  // namespace b200ctc
