// Microbenchmark: dependent-issue latency and per-SM throughput of scalar FP64 (DFMA) on
// sm_100a as a function of warps per SM and independent chains per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(int iters, double a, double b, double* out, long long* clk) {
  double f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) f[i] = i + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int ILP>
void run(int warps) {
  double* out; long long* clk; cudaMalloc(&out, 8); cudaMalloc(&clk, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000;
  chain<ILP><<<sms, warps * 32>>>(10, 0.999, 0.001, out, clk);
  chain<ILP><<<sms, warps * 32>>>(iters, 0.999, 0.001, out, clk);
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  const double per_dep = double(c) / (iters * 8.0);          // clocks per dependent step of one warp
  const double fma_per_clk_sm = warps * 32.0 * ILP / per_dep;  // lanes retired per clock per SM
  printf("warps/SM %2d  ILP %2d : %6.2f clk per dependent DFMA step, %6.1f DFMA lanes/clk/SM\n", warps, ILP, per_dep,
         fma_per_clk_sm);
  cudaFree(out); cudaFree(clk);
}

// throughput of other FP64-pipe instructions (16 warps/SM, 4 independent chains per warp)
template <int OP>
__global__ void ops(int iters, double a, double b, double* out, long long* clk) {
  double f[4];
  int k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[i] = 1.0 + i + threadIdx.x; k[i] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
        if (OP == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(a));
        if (OP == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(b));
        if (OP == 3) { asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(k[i]) : "d"(f[i])); asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(f[i]) : "r"(k[i] + 1)); }
        if (OP == 4) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b)); asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(a)); }
        if (OP == 5) { float x; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x) : "d"(f[i])); asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(x)); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(f[i]) : "f"(x)); }
      }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += f[i] + k[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int OP>
void run_op(const char* name, int nops) {
  double* out; long long* clk; cudaMalloc(&out, 8); cudaMalloc(&clk, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000, warps = 16;
  ops<OP><<<sms, warps * 32>>>(10, 0.999, 0.001, out, clk);
  ops<OP><<<sms, warps * 32>>>(iters, 0.999, 0.001, out, clk);
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  const double inst = double(iters) * 8 * 4 * nops * warps;  // warp instructions per SM
  printf("%-28s %6.2f clk per warp instruction per SM sub-partition (16 warps, ILP 4)\n", name, double(c) / (inst / 4.0));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {1, 4, 8, 16, 24, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  run_op<0>("DFMA", 1);
  run_op<1>("DMUL", 1);
  run_op<2>("DADD", 1);
  run_op<3>("F2I.F64 + I2F.F64 (pair)", 2);
  run_op<4>("DFMA + DMUL (pair)", 2);
  run_op<5>("F2F.F32.F64+MUFU.RSQ+F2F.F64.F32", 3);
  return 0;
}
