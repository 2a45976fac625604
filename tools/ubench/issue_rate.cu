// Micro-benchmark (B200): issue cost of the instruction classes the lattice inner loop is made of,
// at the low occupancies the lattice runs at (1 CTA per SM, 2..4 warps per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_rate issue_rate.cu && ./issue_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmaf_v(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float max2(float a, float b) { float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ int iadd_v(int a, int b) { int r; asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, const int* idx) {
  __shared__ __align__(16) float sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  float f[8]; int n[8];
  for (int i = 0; i < 8; ++i) { f[i] = 1.0f + threadIdx.x * 1e-3f + i; n[i] = threadIdx.x + i; }
  u64 p[4]; for (int i = 0; i < 4; ++i) p[i] = pk(f[2 * i], f[2 * i + 1]);
  const float m = 0.99999f, c = 1e-9f; const u64 m2 = pk(m, m), c2 = pk(c, c);
  const int lane = threadIdx.x & 31;
  int i0 = idx[threadIdx.x] & 1023, i1 = idx[threadIdx.x + 1] & 1023, i2 = idx[threadIdx.x + 2] & 1023, i3 = idx[threadIdx.x + 3] & 1023;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { for (int i = 0; i < 8; ++i) f[i] = fmaf_v(f[i], m, c); }                       // 8 FFMA
    if (MODE == 1) { for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], m2, c2); }                       // 4 FFMA2
    if (MODE == 2) { for (int i = 0; i < 4; ++i) { f[i] = fmaf_v(f[i], m, c); n[i] = iadd_v(n[i], it); f[i + 4] = max2(f[i + 4], f[i]); } }  // 4 FFMA + 4 IADD + 4 FMNMX
    if (MODE == 3) { for (int i = 0; i < 4; ++i) { p[i] = fma2(p[i], m2, c2); n[i] = iadd_v(n[i], it); } }   // 4 FFMA2 + 4 IADD
    if (MODE == 4) { for (int i = 0; i < 4; ++i) f[i] = __shfl_up_sync(0xffffffffu, f[i], 1) ; }    // 4 SHFL
    if (MODE == 5) { f[0] += sh[i0]; f[1] += sh[i1]; f[2] += sh[i2]; f[3] += sh[i3]; i0 = (i0 + 33) & 1023; }  // 4 LDS.32 random + 4 FADD + 2
    if (MODE == 6) { float4 q = *reinterpret_cast<const float4*>(&sh[((lane * 4 + it * 128) & 1023)]); f[0] += q.x; f[1] += q.y; f[2] += q.z; f[3] += q.w; }  // LDS.128
    if (MODE == 7) { asm volatile("bar.sync 1, 128;" ::: "memory"); f[0] = fmaf_v(f[0], m, c); }    // named barrier over 4 warps
    if (MODE == 8) { f[0] = max3(f[0], f[1], f[2]); f[3] = max3(f[3], f[4], f[5]); f[6] = max3(f[6], f[7], f[0]); f[1] = max2(f[3], f[6]); }  // 3 FMNMX3 + 1 FMNMX (dependent)
    if (MODE == 9) { int v = __reduce_max_sync(0xffffffffu, n[0]); n[0] = v + lane; }               // CREDUX + IADD
    if (MODE == 10) { sh[i0] = f[0]; sh[i1] = f[1]; sh[i2] = f[2]; sh[i3] = f[3]; i0 = (i0 + 33) & 1023; }   // 4 STS.32 random
  }
  for (int i = 0; i < 4; ++i) { float a, b; upk(p[i], a, b); f[i] += a + b; }
  float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + n[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sh[(threadIdx.x * 7) & 1023];
}
template <int MODE>
void run(const char* name, int warps, double ops_per_iter, const int* idx) {
  const int sms = 148, iters = 20000;
  float* out; cudaMalloc(&out, sizeof(float) * sms * warps * 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, warps * 32>>>(out, 100, idx);
  cudaEventRecord(e0);
  k<MODE><<<sms, warps * 32>>>(out, iters, idx);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("%-44s warps/SM %2d: %8.3f ms  %7.2f cyc/iter/warp-slot  %6.2f cyc per warp-instr per SMSP\n", name, warps, ms,
         cycles / iters, cycles / (iters * ops_per_iter * warps / 4.0));
  cudaFree(out);
}
int main() {
  int* idx; cudaMalloc(&idx, 4096 * sizeof(int));
  int h[4096]; unsigned s = 12345; for (int i = 0; i < 4096; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) & 1023; }
  cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w : {4, 8, 12, 16}) {
#define R(M, NAME, OPS) run<M>(NAME, w, OPS, idx)
    R(0, "FFMA x8", 8); R(1, "FFMA2 x4 (same flops)", 4); R(2, "FFMA x4 + IADD x4 + FMNMX x4", 12); R(3, "FFMA2 x4 + IADD x4", 8);
    R(4, "SHFL x4 (dependent chain per reg)", 4); R(5, "LDS.32 random x4 + FADD x4", 8); R(6, "LDS.128 + FADD x4", 5);
    R(8, "FMNMX3 x3 + FMNMX (dependent)", 4); R(9, "CREDUX.MAX + IADD", 2); R(10, "STS.32 random x4", 4);
    if (w == 4) R(7, "bar.sync(4 warps) + FFMA", 2);
  }
  return 0;
}
