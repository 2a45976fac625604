// Micro-benchmark: cost of the intra-CTA hand-off primitives the lattice kernel can use between warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/sync_cost tools/ubench/sync_cost.cu
// Modes (cycles per iteration, measured by warp 0):
//   0  empty loop with a dependent FMA chain of 16 (baseline)
//   1  + bar.sync among NW warps
//   2  + mbarrier self hand-off: lane 0 arrives (count 1), the warp waits on the parity (success at once)
//   3  + ping-pong between two warps through two mbarriers (round trip = 2 hand-offs)
//   4  + ping-pong between two warps through volatile shared flags
//   5  + __syncwarp only
#include <cstdio>
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool mb_try(unsigned long long* b, unsigned par) {
  unsigned d;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(d) : "r"(s32(b)), "r"(par) : "memory");
  return d != 0;
}
__device__ __forceinline__ void mb_wait(unsigned long long* b, unsigned par) { while (!mb_try(b, par)) {} }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, int nw) {
  __shared__ unsigned long long mb[4];
  __shared__ volatile int flag[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mb_init(mb, 1); mb_init(mb + 1, 1); flag[0] = 0; flag[1] = 0; }
  __syncthreads();
  float a = 1.0f + lane * 1e-3f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) a = fmaf(a, 1.0001f, 1e-6f);
    if (MODE == 1) asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
    if (MODE == 2) { __syncwarp(); if (lane == 0) mb_arrive(mb); mb_wait(mb, it & 1); }
    if (MODE == 3) {
      if (warp == 0) { __syncwarp(); if (lane == 0) mb_arrive(mb); mb_wait(mb + 1, it & 1); }
      else if (warp == 1) { mb_wait(mb, it & 1); __syncwarp(); if (lane == 0) mb_arrive(mb + 1); }
    }
    if (MODE == 4) {
      if (warp == 0) { __syncwarp(); if (lane == 0) flag[0] = it + 1; while (flag[1] < it + 1) {} }
      else if (warp == 1) { while (flag[0] < it + 1) {} __syncwarp(); if (lane == 0) flag[1] = it + 1; }
    }
    if (MODE == 5) __syncwarp();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a;
}

template <int MODE>
void run(const char* name, int nw, float* out, long long* cyc) {
  const int iters = 20000;
  k<MODE><<<1, nw * 32>>>(out, cyc, iters, nw);
  k<MODE><<<1, nw * 32>>>(out, cyc, iters, nw);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps %2d: %7.1f cycles/iter  (%s)\n", name, nw, (double)h / iters, cudaGetErrorString(e));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  run<0>("baseline (16 dependent FMAs)", 1, out, cyc);
  run<5>("+ __syncwarp", 1, out, cyc);
  for (int nw : {2, 4, 8, 16}) run<1>("+ bar.sync", nw, out, cyc);
  run<2>("+ mbarrier arrive + try_wait (same warp)", 1, out, cyc);
  run<3>("+ mbarrier ping-pong (2 hand-offs)", 2, out, cyc);
  run<4>("+ volatile-flag ping-pong (2 hand-offs)", 2, out, cyc);
  return 0;
}
