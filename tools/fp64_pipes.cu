// Microbenchmark: are the FP64 tensor pipe (DMMA) and the FP64 FMA pipe (DFMA)
// independent on sm_100a?  Times DMMA-only, DFMA-only and an interleaved mix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void __launch_bounds__(256) mix(int iters, double* out) {
  double c[8][2], f[16];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NM) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (i * 2 + j < NF) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i * 2 + j]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}

template <int NM, int NF>
void run(const char* name, int warps_per_block) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8);
  const int iters = 20000;
  dim3 grid(sms * 2), block(warps_per_block * 32);
  mix<NM, NF><<<grid, block>>>(100, out);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mix<NM, NF><<<grid, block>>>(iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps = double(grid.x) * warps_per_block;
  double fl_mma = warps * iters * NM * 512.0 * 2.0, fl_fma = warps * iters * NF * 32.0 * 2.0;
  printf("%-28s warps/blk %2d  %8.3f ms   DMMA %7.2f TF/s   DFMA %7.2f TF/s   sum %7.2f\n", name, warps_per_block, ms,
         fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8}) {
    run<8, 0>("DMMA only (8 indep)", w);
    run<4, 0>("DMMA only (4 indep)", w);
    run<2, 0>("DMMA only (2 indep)", w);
    run<0, 16>("DFMA only (16 indep)", w);
    run<8, 16>("DMMA 8 + DFMA 16", w);
    run<8, 8>("DMMA 8 + DFMA 8", w);
    run<8, 4>("DMMA 8 + DFMA 4", w);
    run<4, 16>("DMMA 4 + DFMA 16", w);
  }
  return 0;
}
