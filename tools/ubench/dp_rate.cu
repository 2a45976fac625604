// Micro-benchmark: per-SM issue rate of DFMA / DADD / FFMA / F2F(f64->f32) / SHFL on B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = (float)a0, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c); }
    if (MODE == 1) { f0 = fmaf(f0, 0.99999f, 1e-9f); f1 = fmaf(f1, 0.99999f, 1e-9f); f2 = fmaf(f2, 0.99999f, 1e-9f); f3 = fmaf(f3, 0.99999f, 1e-9f); f4 = fmaf(f4, 0.99999f, 1e-9f); f5 = fmaf(f5, 0.99999f, 1e-9f); f6 = fmaf(f6, 0.99999f, 1e-9f); f7 = fmaf(f7, 0.99999f, 1e-9f); }
    if (MODE == 2) { f0 += (float)a0; f1 += (float)a1; f2 += (float)a2; f3 += (float)a3; f4 += (float)a4; f5 += (float)a5; f6 += (float)a6; f7 += (float)a7;
                     a0 += 1e-3; a1 += 1e-3; a2 += 1e-3; a3 += 1e-3; a4 += 1e-3; a5 += 1e-3; a6 += 1e-3; a7 += 1e-3; }   // 8 F2F + 8 DADD + 8 FADD
    if (MODE == 3) { a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c; }
    if (MODE == 4) { a0 = __shfl_up_sync(0xffffffffu, a0, 1) + c; a1 = __shfl_up_sync(0xffffffffu, a1, 1) + c; }  // 4 SHFL + 2 DADD
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}
template <int MODE>
void run(const char* name, int warps_per_sm, double ops_per_iter) {
  int dev_sms = 148, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * dev_sms * warps_per_sm * 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<dev_sms, warps_per_sm * 32>>>(out, 100);
  cudaEventRecord(e0);
  k<MODE><<<dev_sms, warps_per_sm * 32>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr_per_sm = (double)iters * ops_per_iter * warps_per_sm;
  double cycles = ms * 1e-3 * 1.965e9;
  printf("%-28s warps/SM %2d: %.3f ms -> %.2f cycles per warp-instr per SM (%.2f per SMSP)\n", name, warps_per_sm, ms,
         cycles / warp_instr_per_sm, 4 * cycles / warp_instr_per_sm);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("DFMA x8", 4, 8); run<1>("FFMA x8", 4, 8); run<2>("F2F+DADD+FADD x8", 4, 24); run<3>("DADD x8", 4, 8); run<4>("SHFL64 x2 + DADD x2", 4, 6); }
    if (w == 8) { run<0>("DFMA x8", 8, 8); run<1>("FFMA x8", 8, 8); run<2>("F2F+DADD+FADD x8", 8, 24); run<3>("DADD x8", 8, 8); run<4>("SHFL64 x2 + DADD x2", 8, 6); }
    if (w == 16) { run<0>("DFMA x8", 16, 8); run<1>("FFMA x8", 16, 8); run<2>("F2F+DADD+FADD x8", 16, 24); run<3>("DADD x8", 16, 8); run<4>("SHFL64 x2 + DADD x2", 16, 6); }
  }
  return 0;
}
