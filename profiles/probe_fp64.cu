// Micro-benchmark: throughput of DADD / DMUL / DFMA / F2F.F64.F32 / F2F.F32.F64 / I2F.F64 per SM
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/probe_fp64 profiles/probe_fp64.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int OP>
__global__ void k(double* out, float* fout, int iters, double seed) {
  double a[8]; float f[8]; unsigned u[8];
  for (int j = 0; j < 8; ++j) { a[j] = seed + threadIdx.x * 1e-3 + j; f[j] = (float)a[j]; u[j] = threadIdx.x + j; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) a[j] = a[j] + seed;
      if (OP == 1) a[j] = a[j] * seed;
      if (OP == 2) a[j] = fma(a[j], seed, seed);
      if (OP == 3) { a[j] = (double)f[j]; f[j] = f[j] + 1.0f; }              // F2F.F64.F32 + FADD
      if (OP == 4) { f[j] = (float)a[j]; a[j] = __hiloint2double(__double2hiint(a[j]) ^ it, __float_as_int(f[j])); }  // F2F.F32.F64
      if (OP == 5) { a[j] = (double)u[j]; u[j] += __double2loint(a[j]); }     // I2F.F64.U32
      if (OP == 6) f[j] = fmaf(f[j], 1.0001f, 0.5f);                          // FFMA reference
      if (OP == 7) { u[j] = __float2uint_rn(f[j]); f[j] = f[j] + (float)(u[j] & 1); }  // F2I + I2F.F32
    }
  }
  double s = 0; float fs = 0;
  for (int j = 0; j < 8; ++j) { s += a[j]; fs += f[j] + u[j]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; fout[blockIdx.x * blockDim.x + threadIdx.x] = fs;
}
int main() {
  double* out; float* fout; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&fout, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"DADD", "DMUL", "DFMA", "F2F.F64.F32(+FADD)", "F2F.F32.F64(+misc)", "I2F.F64.U32(+misc)", "FFMA", "F2I+I2F.F32(+FADD)"};
  const int iters = 4096;
  for (int op = 0; op < 8; ++op) {
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      switch (op) {
        case 0: k<0><<<148, 1024>>>(out, fout, iters, 1.000001); break;
        case 1: k<1><<<148, 1024>>>(out, fout, iters, 1.000001); break;
        case 2: k<2><<<148, 1024>>>(out, fout, iters, 1.000001); break;
        case 3: k<3><<<148, 1024>>>(out, fout, iters, 1.000001); break;
        case 4: k<4><<<148, 1024>>>(out, fout, iters, 1.000001); break;
        case 5: k<5><<<148, 1024>>>(out, fout, iters, 1.000001); break;
        case 6: k<6><<<148, 1024>>>(out, fout, iters, 1.000001); break;
        case 7: k<7><<<148, 1024>>>(out, fout, iters, 1.000001); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double ops = 148.0 * 1024 * iters * 8;
    printf("%-22s %.3f ms  %.1f lane-ops/clk/SM (at 1.9 GHz)  %.2f Tops/s\n", names[op], best, ops / (best * 1e-3) / 148 / 1.9e9, ops / (best * 1e-3) / 1e12);
  }
  return 0;
}
