// Microbenchmark: DMMA.8x8x4 issue rate of ONE warp vs several warps per SM sub-partition (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_issue dmma_issue.cu
// One CTA per SM, W warps per CTA (warp w sits on sub-partition w % 4); every warp issues `iters` rounds of
// NI independent DMMAs.  Reports cycles per DMMA per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NI>
__global__ void __launch_bounds__(512) issue(int iters, double* out, long long* clk) {
  double c[NI][2];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NI; ++i) c[i][0] = c[i][1] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NI; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NI; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) clk[0] = t1 - t0;
}

template <int NI>
void run(int warps) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; long long* clk; cudaMalloc(&out, 8); cudaMalloc(&clk, 8);
  const int iters = 4000;
  issue<NI><<<sms, warps * 32>>>(100, out, clk);
  issue<NI><<<sms, warps * 32>>>(iters, out, clk);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double per_warp = double(h) / (double(iters) * NI);
  const double per_smsp = per_warp / (warps >= 4 ? warps / 4.0 : 1.0);
  printf("NI=%d warps/SM=%2d (%.1f per sub-partition): %.1f clk per DMMA per warp, %.1f clk per DMMA per sub-partition\n",
         NI, warps, warps / 4.0, per_warp, per_smsp);
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {1, 4, 8, 12, 16}) { run<8>(w); run<4>(w); run<2>(w); run<1>(w); }
  return 0;
}
