// FP64 tensor-core (DMMA m8n8k4) GEMM building block for the batched stack
// contractions (K5, K7) and other GEMM-shaped steps of the prediction path.
//
//   C[z][m][n] = sum_{k in split z} A[m][k] * B(k, n)         m < M, n < N
//
//   A is row-major [M][lda] (K contiguous).
//   B_NMAJOR == false: B(k, n) = B[n][k], row-major [N][ldb] (K contiguous)   "NT"
//   B_NMAJOR == true : B(k, n) = B[k][n], row-major [K][ldb] (N contiguous)   "NN"
//   split-K: blockIdx.z owns k in [z*kchunk, min(K, (z+1)*kchunk)), kchunk % BK == 0;
//   the partial results are written to separate slabs C + z*c_split_stride and are
//   summed by the caller in a fixed order (bit-reproducible, no atomics).
//
// Tiles: CTA BM x BN x BK(16), 3-stage cp.async pipeline; shared-memory pitches are
// == 4 (mod 16) doubles so that every DMMA fragment load (lane (g,tg) reads row g,
// column tg) hits 32 distinct banks pairs; warp tile (8*MI) x (8*NI) with the
// accumulators in registers.  Operand rows that are not 16-byte aligned (odd lda /
// ldb, e.g. n^4 = 28561 for 13 orbitals) use 8-byte cp.async (VEC == 1).
#pragma once
#include "common.cuh"

namespace evc_gemm {

constexpr int BK = 16;
constexpr int STAGES = 3;

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem, bool pred) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const int src = pred ? BYTES : 0;  // src-size 0 => zero fill
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int BM, int BN, bool B_NMAJOR>
struct Smem {
  static constexpr int A_PITCH = BK + 4;
  static constexpr int B_PITCH = B_NMAJOR ? BN + 4 : BK + 4;
  static constexpr int A_STAGE = BM * A_PITCH;
  static constexpr int B_STAGE = B_NMAJOR ? BK * B_PITCH : BN * B_PITCH;
  static constexpr size_t BYTES = static_cast<size_t>(STAGES) * (A_STAGE + B_STAGE) * sizeof(double);
};

template <int BM, int BN, int WARPS_M, int WARPS_N, bool B_NMAJOR, int VEC>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, (WARPS_M * WARPS_N <= 8) ? 2 : 1)
dgemm_kernel(int M, int N, int K, int kchunk, const double* __restrict__ A, int64_t lda,
             const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc,
             int64_t c_split_stride) {
  using S = Smem<BM, BN, B_NMAJOR>;
  constexpr int NT = WARPS_M * WARPS_N * 32;
  constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
  constexpr int MI = WTM / 8, NI = WTN / 8;
  static_assert(WTM % 8 == 0 && WTN % 8 == 0, "warp tile must be a multiple of 8x8");
  static_assert(!B_NMAJOR || BN % 16 == 0, "N-major B needs BN % 16 == 0 for the bank-conflict-free pitch");
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * S::A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int wm0 = (warp / WARPS_N) * WTM, wn0 = (warp % WARPS_N) * WTN;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int nk = (kend - kbeg + BK - 1) / BK;

  auto load_tile = [&](int stage, int kt) {
    const int k0 = kbeg + kt * BK;
    double* as = As + stage * S::A_STAGE;
    double* bs = Bs + stage * S::B_STAGE;
    constexpr int GA = BM * (BK / VEC);  // granules of VEC doubles
    for (int idx = tid; idx < GA; idx += NT) {
      const int r = idx / (BK / VEC), c = (idx - r * (BK / VEC)) * VEC;
      const bool ok = (m0 + r < M) && (k0 + c < kend);
      const double* src = ok ? A + static_cast<int64_t>(m0 + r) * lda + k0 + c : A;
      cp_async<VEC * 8>(as + r * S::A_PITCH + c, src, ok);
    }
    if constexpr (!B_NMAJOR) {
      constexpr int GB = BN * (BK / VEC);
      for (int idx = tid; idx < GB; idx += NT) {
        const int r = idx / (BK / VEC), c = (idx - r * (BK / VEC)) * VEC;
        const bool ok = (n0 + r < N) && (k0 + c < kend);
        const double* src = ok ? B + static_cast<int64_t>(n0 + r) * ldb + k0 + c : B;
        cp_async<VEC * 8>(bs + r * S::B_PITCH + c, src, ok);
      }
    } else {
      constexpr int GB = BK * (BN / VEC);
      for (int idx = tid; idx < GB; idx += NT) {
        const int r = idx / (BN / VEC), c = (idx - r * (BN / VEC)) * VEC;
        const bool ok = (k0 + r < kend) && (n0 + c < N);  // VEC == 2 requires N even
        const double* src = ok ? B + static_cast<int64_t>(k0 + r) * ldb + n0 + c : B;
        cp_async<VEC * 8>(bs + r * S::B_PITCH + c, src, ok);
      }
    }
  };

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_tile(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < nk) load_tile((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
    cp_async_commit();
    const double* as = As + (kt % STAGES) * S::A_STAGE;
    const double* bs = Bs + (kt % STAGES) * S::B_STAGE;
    // k-steps of this tile that hold data: the last tile of a K extent that is no multiple of BK (K7p: 210 =
    // 13 x 16 + 2) runs one k-step instead of four (the rest of the tile is zero padding)
    const int kvalid = kend - (kbeg + kt * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      if (kk >= kvalid) break;  // block-uniform
      double af[MI], bf[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] = as[(wm0 + i * 8 + g) * S::A_PITCH + kk + tg];
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        if constexpr (!B_NMAJOR) bf[j] = bs[(wn0 + j * 8 + g) * S::B_PITCH + kk + tg];
        else bf[j] = bs[(kk + tg) * S::B_PITCH + wn0 + j * 8 + g];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  double* Cz = C + static_cast<int64_t>(blockIdx.z) * c_split_stride;
  const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cz) & 15) == 0);
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int m = m0 + wm0 + i * 8 + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int n = n0 + wn0 + j * 8 + tg * 2;
      double* dst = Cz + static_cast<int64_t>(m) * ldc + n;
      if (vec_ok && n + 1 < N) {
        *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (n < N) dst[0] = acc[i][j][0];
        if (n + 1 < N) dst[1] = acc[i][j][1];
      }
    }
  }
}

struct Plan {
  int nsplit;
  int kchunk;
};

// split-K: the number of splits that minimises (waves of CTAs) x (k-steps per CTA), two CTAs resident per SM, never
// below 4 k-steps a split.  (Covering the SMs "about twice" left 384 CTAs on 296 slots for the K5p product of the bench:
// a second wave that is 30 % full; three splits = 288 CTAs run as one wave and finish 1.5x sooner.)  A small term per
// split accounts for the reduction of the partial results.
inline Plan plan_split(int tiles, int K, int sm_count, int max_split) {
  Plan p;
  const int ksteps = (K + BK - 1) / BK;
  const int slots = 2 * sm_count;
  int smax = ksteps / 4;
  if (smax > max_split) smax = max_split;
  if (smax < 1) smax = 1;
  int best = 1;
  double best_cost = 1e300;
  for (int s = 1; s <= smax; ++s) {
    const long long ctas = static_cast<long long>(tiles) * s;
    const long long waves = (ctas + slots - 1) / slots;
    const int per = (ksteps + s - 1) / s;
    const double cost = static_cast<double>(waves) * (per + 2) + 0.25 * s;
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  const int per = (ksteps + best - 1) / best;
  p.kchunk = per * BK;
  p.nsplit = (K + p.kchunk - 1) / p.kchunk;
  return p;
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool B_NMAJOR>
int launch(cudaStream_t st, int M, int N, int K, const Plan& pl, const double* A, int64_t lda,
           const double* B, int64_t ldb, double* C, int64_t ldc, int64_t c_split_stride) {
  using S = Smem<BM, BN, B_NMAJOR>;
  const bool vec2 = ((lda & 1) == 0) && ((ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (!B_NMAJOR || (N & 1) == 0) &&
                    (B_NMAJOR || (K & 1) == 0) && ((K & 1) == 0);
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, pl.nsplit);
  dim3 block(WARPS_M * WARPS_N * 32);
  if (vec2) {
    auto kern = dgemm_kernel<BM, BN, WARPS_M, WARPS_N, B_NMAJOR, 2>;
    EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S::BYTES)));
    kern<<<grid, block, S::BYTES, st>>>(M, N, K, pl.kchunk, A, lda, B, ldb, C, ldc, c_split_stride);
  } else {
    auto kern = dgemm_kernel<BM, BN, WARPS_M, WARPS_N, B_NMAJOR, 1>;
    EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S::BYTES)));
    kern<<<grid, block, S::BYTES, st>>>(M, N, K, pl.kchunk, A, lda, B, ldb, C, ldc, c_split_stride);
  }
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // namespace evc_gemm
