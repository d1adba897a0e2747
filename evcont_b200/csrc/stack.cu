// K5 / K7: the two streaming passes over the stored t-RDM stack.
//
//   K5  H[g, ab]    = one_rdm[ab] . h1[g] + (1/2) two_rdm[ab] . h2[g]
//       (evcont/ab_initio_eigenvector_continuation.py:38-71, four layouts)
//   K7  gamma[g]    = sum_ab c_a c_b one_rdm[ab]
//       Gamma[g]    = sum_ab w_ab   two_rdm[ab]     (w: tril-weighted for the
//       data-symmetric layouts; exchange symmetry restored afterwards)
//       (evcont/ab_initio_gradients_loewdin.py:343-361)
//
// The stack is by far the largest object on the prediction path (32 MB at
// H10/N=20, 2.6 GB at H30/N=20, 12 GB at Zundel/N=100) and each element is used
// once per geometry, so both passes are HBM (or L2) streaming reductions:
// vectorised 16-byte loads, a register tile of GB geometries per stack element,
// fixed-order reductions (bit-reproducible).
#include "common.cuh"
#include "dgemm.cuh"

namespace {

constexpr int kGB = 4;             // geometries per register tile
constexpr int64_t kChunk = 16384;  // stack elements per CTA in K5

__host__ __device__ inline int64_t exch_len(int n) {
  const int64_t n2 = static_cast<int64_t>(n) * n;
  return n2 * (n2 + 1) / 2;
}

inline bool layout_is_tril(int layout) { return layout == EVC_LAYOUT_TRIL || layout == EVC_LAYOUT_TRIL_EXCH; }
inline bool layout_is_exch(int layout) { return layout == EVC_LAYOUT_FULL_EXCH || layout == EVC_LAYOUT_TRIL_EXCH; }
inline bool layout_ok(int layout) {
  return layout == EVC_LAYOUT_FULL || layout == EVC_LAYOUT_TRIL || layout == EVC_LAYOUT_FULL_EXCH ||
         layout == EVC_LAYOUT_TRIL_EXCH;
}

// compress_electron_exchange_symmetry(h2, diag_multiplier=0.5)
// (evcont/electron_integral_utils.py:38-66): row-major lower triangle of the
// (n^2 x n^2) matrix, diagonal halved.
__global__ void pack_exchange_kernel(int n2, int64_t Lc, const double* __restrict__ h2,
                                     double* __restrict__ hc) {
  const int g = blockIdx.y;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= Lc) return;
  // t = x(x+1)/2 + y, y <= x
  int64_t x = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (x * (x + 1) / 2 > t) --x;
  while ((x + 1) * (x + 2) / 2 <= t) ++x;
  const int64_t y = t - x * (x + 1) / 2;
  const double v = h2[static_cast<int64_t>(g) * n2 * n2 + x * n2 + y];
  hc[static_cast<int64_t>(g) * Lc + t] = (x == y) ? 0.5 * v : v;
}

__device__ __forceinline__ double block_reduce_sum(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double tot = 0.0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) tot += scratch[w];
  return tot;
}

// partial[g][p][chunk] = sum_{l in chunk} R2[p][l] * hv[g][l]
template <bool VEC2>
__global__ void __launch_bounds__(256)
stack_dot_kernel(const double* __restrict__ R2, int64_t L, int P, const double* __restrict__ hv,
                 int G, int nchunk, double* __restrict__ partial) {
  __shared__ double scratch[8];
  const int p = blockIdx.x, ch = blockIdx.y, g0 = blockIdx.z * kGB;
  const int64_t lo = static_cast<int64_t>(ch) * kChunk;
  const int64_t hi = min(L, lo + kChunk);
  const double* row = R2 + static_cast<int64_t>(p) * L;
  const int ng = min(kGB, G - g0);
  double acc[kGB];
#pragma unroll
  for (int q = 0; q < kGB; ++q) acc[q] = 0.0;
  if (VEC2) {
    for (int64_t l = lo + 2 * threadIdx.x; l < hi; l += 512) {
      const double2 r = __ldg(reinterpret_cast<const double2*>(row + l));
#pragma unroll
      for (int q = 0; q < kGB; ++q) {
        if (q < ng) {
          const double2 h = __ldg(reinterpret_cast<const double2*>(hv + static_cast<int64_t>(g0 + q) * L + l));
          acc[q] += r.x * h.x;
          acc[q] += r.y * h.y;
        }
      }
    }
  } else {
    for (int64_t l = lo + threadIdx.x; l < hi; l += 256) {
      const double r = __ldg(row + l);
#pragma unroll
      for (int q = 0; q < kGB; ++q)
        if (q < ng) acc[q] += r * __ldg(hv + static_cast<int64_t>(g0 + q) * L + l);
    }
  }
#pragma unroll
  for (int q = 0; q < kGB; ++q) {
    const double tot = block_reduce_sum(acc[q], scratch);
    if (threadIdx.x == 0 && q < ng)
      partial[(static_cast<int64_t>(g0 + q) * P + p) * nchunk + ch] = tot;
  }
}

__device__ __forceinline__ void tril_unrank(int64_t t, int& a, int& b) {
  int64_t x = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (x * (x + 1) / 2 > t) --x;
  while ((x + 1) * (x + 2) / 2 <= t) ++x;
  a = static_cast<int>(x);
  b = static_cast<int>(t - x * (x + 1) / 2);
}

// H[g][a][b] = one_rdm[a][b] . h1[g] + scale * sum_chunks partial
// partial element (g, p, c) lives at partial[g*sg + p*sp + c*sc]
__global__ void assemble_H_kernel(int N, int n2, int tril, double scale, int P, int nchunk,
                                  int64_t sg, int64_t sp, int64_t sc,
                                  const double* __restrict__ one_rdm, const double* __restrict__ h1,
                                  const double* __restrict__ partial, double* __restrict__ H) {
  const int g = blockIdx.y;
  const int ab = blockIdx.x * blockDim.x + threadIdx.x;
  if (ab >= N * N) return;
  const int a = ab / N, b = ab - a * N;
  const double* r1 = one_rdm + static_cast<int64_t>(ab) * n2;
  const double* hg = h1 + static_cast<int64_t>(g) * n2;
  double one = 0.0;
  for (int k = 0; k < n2; ++k) one += r1[k] * hg[k];
  double two = 0.0;
  int p = -1;
  if (!tril) p = ab;
  else if (a >= b) p = a * (a + 1) / 2 + b;
  if (p >= 0) {
    const double* pp = partial + g * sg + p * sp;
    for (int c = 0; c < nchunk; ++c) two += pp[c * sc];
  }
  H[static_cast<int64_t>(g) * N * N + ab] = one + scale * two;
}

// w[g][P] from the ground-state vector: c_a c_b (full) or tril-weighted
__global__ void pair_weights_kernel(int N, int tril, int P, const double* __restrict__ C,
                                    int64_t c_stride, double* __restrict__ w) {
  const int g = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double* c = C + static_cast<int64_t>(g) * c_stride;
  int a, b;
  double v;
  if (tril) {
    tril_unrank(p, a, b);
    v = (a == b) ? c[a] * c[a] : 2.0 * c[a] * c[b];
  } else {
    a = p / N; b = p - a * N;
    v = c[a] * c[b];
  }
  w[static_cast<int64_t>(g) * P + p] = v;
}

// gamma[g][pq] = sum_ab c_a c_b one_rdm[ab][pq]
__global__ void gamma1_kernel(int N, int n2, const double* __restrict__ one_rdm,
                              const double* __restrict__ C, int64_t c_stride,
                              double* __restrict__ gamma) {
  const int g = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  const double* c = C + static_cast<int64_t>(g) * c_stride;
  double acc = 0.0;
  for (int a = 0; a < N; ++a) {
    const double ca = c[a];
    for (int b = 0; b < N; ++b)
      acc += ca * c[b] * one_rdm[(static_cast<int64_t>(a) * N + b) * n2 + k];
  }
  gamma[static_cast<int64_t>(g) * n2 + k] = acc;
}

// part[split][g][l] = sum_{p in split} w[g][p] R2[p][l]
template <bool VEC2>
__global__ void __launch_bounds__(256)
stack_axpy_kernel(const double* __restrict__ R2, int64_t L, int P, const double* __restrict__ w,
                  int G, int nsplit, double* __restrict__ part) {
  const int split = blockIdx.y, g0 = blockIdx.z * kGB;
  const int p0 = static_cast<int>(static_cast<int64_t>(P) * split / nsplit);
  const int p1 = static_cast<int>(static_cast<int64_t>(P) * (split + 1) / nsplit);
  const int ng = min(kGB, G - g0);
  const double* wg = w + static_cast<int64_t>(g0) * P;
  if (VEC2) {
    const int64_t l = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) * 2;
    if (l >= L) return;
    double2 acc[kGB];
#pragma unroll
    for (int q = 0; q < kGB; ++q) acc[q] = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const double2 r = __ldg(reinterpret_cast<const double2*>(R2 + static_cast<int64_t>(p) * L + l));
#pragma unroll
      for (int q = 0; q < kGB; ++q) {
        if (q < ng) {
          const double wv = __ldg(wg + static_cast<int64_t>(q) * P + p);
          acc[q].x += wv * r.x;
          acc[q].y += wv * r.y;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kGB; ++q)
      if (q < ng)
        *reinterpret_cast<double2*>(part + (static_cast<int64_t>(split) * G + g0 + q) * L + l) = acc[q];
  } else {
    const int64_t l = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (l >= L) return;
    double acc[kGB];
#pragma unroll
    for (int q = 0; q < kGB; ++q) acc[q] = 0.0;
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const double r = __ldg(R2 + static_cast<int64_t>(p) * L + l);
#pragma unroll
      for (int q = 0; q < kGB; ++q)
        if (q < ng) acc[q] += __ldg(wg + static_cast<int64_t>(q) * P + p) * r;
    }
#pragma unroll
    for (int q = 0; q < kGB; ++q)
      if (q < ng) part[(static_cast<int64_t>(split) * G + g0 + q) * L + l] = acc[q];
  }
}

// Gamma[g][x][y] = sum_split part[split][g][idx(x,y)];  idx = x*n2+y (plain) or the
// lower-triangle index of (max, min) (restore_electron_exchange_symmetry,
// evcont/electron_integral_utils.py:69-88)
__global__ void gamma2_finalize_kernel(int n2, int exch, int64_t L, int G, int nsplit,
                                       const double* __restrict__ part, double* __restrict__ Gamma) {
  const int g = blockIdx.y;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n4 = static_cast<int64_t>(n2) * n2;
  if (k >= n4) return;
  int64_t idx = k;
  if (exch) {
    const int64_t x = k / n2, y = k - x * n2;
    const int64_t hi = x > y ? x : y, lo = x > y ? y : x;
    idx = hi * (hi + 1) / 2 + lo;
  }
  double acc = 0.0;
  for (int s = 0; s < nsplit; ++s) acc += part[(static_cast<int64_t>(s) * G + g) * L + idx];
  Gamma[static_cast<int64_t>(g) * n4 + k] = acc;
}

// Batches of more than kGemvMaxBatch geometries go through the DMMA GEMM kernels
// (dgemm.cuh); smaller ones through the streaming GEMV-style kernels above, which
// read the stack exactly once at HBM/L2 speed.
constexpr int kGemvMaxBatch = 4;
constexpr int kPlanSms = 148;  // B200; the plan must not depend on the ctx (workspace sizing)
constexpr int kK5BM = 128, kK5BN = 80;

evc_gemm::Plan k5_plan(int G, int P, int64_t L) {
  const int tiles = ((G + kK5BM - 1) / kK5BM) * ((P + kK5BN - 1) / kK5BN);
  return evc_gemm::plan_split(tiles, static_cast<int>(L), kPlanSms, 64);
}

int axpy_nsplit(int64_t L, int P, int G) {
  const int64_t blocks = ((L + 511) / 512) * ((G + kGB - 1) / kGB);
  int64_t s = (2 * 148 + blocks - 1) / blocks;
  if (s > 16) s = 16;
  if (s > P) s = P;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

}  // namespace

extern "C" {

int evc_subspace_workspace_bytes(int layout, int N, int n, int nbatch, size_t* bytes) {
  EVC_REQUIRE(bytes && layout_ok(layout), "evc_subspace_workspace_bytes: bad layout %d", layout);
  const int64_t n4 = static_cast<int64_t>(n) * n * n * n;
  const int64_t L = layout_is_exch(layout) ? exch_len(n) : n4;
  const int64_t P = layout_is_tril(layout) ? static_cast<int64_t>(N) * (N + 1) / 2 : static_cast<int64_t>(N) * N;
  int64_t nchunk = (L + kChunk - 1) / kChunk;
  if (nbatch > kGemvMaxBatch) nchunk = k5_plan(nbatch, static_cast<int>(P), L).nsplit;
  size_t tot = evc_align_up(static_cast<size_t>(nbatch) * P * nchunk * 8, 256);
  if (layout_is_exch(layout)) tot += evc_align_up(static_cast<size_t>(nbatch) * L * 8, 256);
  *bytes = tot;
  return 0;
}

int evc_subspace_H(evc_ctx* ctx, int layout, int N, int n, const double* one_rdm,
                   const double* two_rdm, int nbatch, const double* h1, const double* h2, double* H,
                   void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && one_rdm && two_rdm && h1 && h2 && H && workspace, "evc_subspace_H: NULL argument");
  EVC_REQUIRE(layout_ok(layout), "evc_subspace_H: two_RDM layout %d not one of 6/5/3/2", layout);
  EVC_REQUIRE(N >= 1 && n >= 1, "evc_subspace_H: empty problem");
  if (nbatch <= 0) return 0;
  const int n2 = n * n;
  const int64_t n4 = static_cast<int64_t>(n2) * n2;
  const bool exch = layout_is_exch(layout), tril = layout_is_tril(layout);
  const int64_t L = exch ? exch_len(n) : n4;
  const int P = tril ? N * (N + 1) / 2 : N * N;
  const bool gemm = nbatch > kGemvMaxBatch;
  EVC_REQUIRE(L < (int64_t(1) << 31), "evc_subspace_H: n=%d too large", n);
  const evc_gemm::Plan k5 = k5_plan(nbatch, P, L);
  const int nchunk = gemm ? k5.nsplit : static_cast<int>((L + kChunk - 1) / kChunk);
  evc_arena ar(workspace, workspace_bytes);
  double* partial = ar.take<double>(static_cast<size_t>(nbatch) * P * nchunk);
  double* hc = exch ? ar.take<double>(static_cast<size_t>(nbatch) * L) : nullptr;
  EVC_REQUIRE(partial && (!exch || hc), "evc_subspace_H: workspace too small (%zu bytes)", workspace_bytes);
  const double* hv = h2;
  if (exch) {
    dim3 grid(static_cast<unsigned>((L + 255) / 256), nbatch);
    pack_exchange_kernel<<<grid, 256, 0, ctx->stream>>>(n2, L, h2, hc);
    EVC_CHECK_LAUNCH();
    hv = hc;
  }
  if (gemm) {
    // partial[z][g][p] = sum_{l in split z} hv[g][l] R2[p][l]   (DMMA, split-K)
    int rc = evc_gemm::launch<kK5BM, kK5BN, 4, 2, false>(ctx->stream, nbatch, P, static_cast<int>(L), k5, hv, L,
                                                          two_rdm, L, partial, P,
                                                          static_cast<int64_t>(nbatch) * P);
    if (rc) return rc;
  } else {
    dim3 grid(P, nchunk, (nbatch + kGB - 1) / kGB);
    EVC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "evc_subspace_H: batch/stack too large for one launch");
    const bool vec2 = (L % 2 == 0) && ((reinterpret_cast<uintptr_t>(two_rdm) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(hv) & 15) == 0);
    if (vec2) stack_dot_kernel<true><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, hv, nbatch, nchunk, partial);
    else stack_dot_kernel<false><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, hv, nbatch, nchunk, partial);
    EVC_CHECK_LAUNCH();
  }
  {
    dim3 grid((N * N + 127) / 128, nbatch);
    const int64_t sg = gemm ? P : static_cast<int64_t>(P) * nchunk;
    const int64_t sp = gemm ? 1 : nchunk;
    const int64_t sc = gemm ? static_cast<int64_t>(nbatch) * P : 1;
    assemble_H_kernel<<<grid, 128, 0, ctx->stream>>>(N, n2, tril ? 1 : 0, exch ? 1.0 : 0.5, P, nchunk, sg, sp, sc,
                                                     one_rdm, h1, partial, H);
    EVC_CHECK_LAUNCH();
  }
  return 0;
}

int evc_predict_workspace_bytes(int layout, int N, int n, int nbatch, size_t* bytes) {
  EVC_REQUIRE(bytes && layout_ok(layout), "evc_predict_workspace_bytes: bad layout %d", layout);
  const int64_t n4 = static_cast<int64_t>(n) * n * n * n;
  const int64_t L = layout_is_exch(layout) ? exch_len(n) : n4;
  const int P = layout_is_tril(layout) ? N * (N + 1) / 2 : N * N;
  const int nsplit = nbatch > kGemvMaxBatch ? 1 : axpy_nsplit(L, P, nbatch);
  *bytes = evc_align_up(static_cast<size_t>(nbatch) * P * 8, 256) +
           evc_align_up(static_cast<size_t>(nsplit) * nbatch * L * 8, 256);
  return 0;
}

int evc_predict_rdm(evc_ctx* ctx, int layout, int N, int n, const double* one_rdm,
                    const double* two_rdm, int nbatch, const double* C, int64_t c_stride,
                    double* gamma, double* Gamma, void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && one_rdm && two_rdm && C && gamma && Gamma && workspace, "evc_predict_rdm: NULL argument");
  EVC_REQUIRE(layout_ok(layout), "evc_predict_rdm: two_RDM layout %d not one of 6/5/3/2", layout);
  EVC_REQUIRE(c_stride >= N, "evc_predict_rdm: c_stride < N");
  if (nbatch <= 0) return 0;
  const int n2 = n * n;
  const int64_t n4 = static_cast<int64_t>(n2) * n2;
  const bool exch = layout_is_exch(layout), tril = layout_is_tril(layout);
  const int64_t L = exch ? exch_len(n) : n4;
  const int P = tril ? N * (N + 1) / 2 : N * N;
  const bool gemm = nbatch > kGemvMaxBatch;
  EVC_REQUIRE(L < (int64_t(1) << 31), "evc_predict_rdm: n=%d too large", n);
  const int nsplit = gemm ? 1 : axpy_nsplit(L, P, nbatch);
  evc_arena ar(workspace, workspace_bytes);
  double* w = ar.take<double>(static_cast<size_t>(nbatch) * P);
  double* part = ar.take<double>(static_cast<size_t>(nsplit) * nbatch * L);
  EVC_REQUIRE(w && part, "evc_predict_rdm: workspace too small (%zu bytes)", workspace_bytes);
  {
    dim3 grid((P + 127) / 128, nbatch);
    pair_weights_kernel<<<grid, 128, 0, ctx->stream>>>(N, tril ? 1 : 0, P, C, c_stride, w);
    EVC_CHECK_LAUNCH();
  }
  {
    dim3 grid((n2 + 127) / 128, nbatch);
    gamma1_kernel<<<grid, 128, 0, ctx->stream>>>(N, n2, one_rdm, C, c_stride, gamma);
    EVC_CHECK_LAUNCH();
  }
  if (gemm) {
    // Gamma[g][l] = sum_p w[g][p] R2[p][l]   (DMMA; written in place unless the
    // exchange symmetry still has to be restored)
    evc_gemm::Plan pl;
    pl.nsplit = 1;
    pl.kchunk = (P + evc_gemm::BK - 1) / evc_gemm::BK * evc_gemm::BK;
    double* dst = exch ? part : Gamma;
    int rc = evc_gemm::launch<64, 128, 2, 4, true>(ctx->stream, nbatch, static_cast<int>(L), P, pl, w, P, two_rdm, L,
                                                   dst, L, 0);
    if (rc) return rc;
  } else {
    const bool vec2 = (L % 2 == 0) && ((reinterpret_cast<uintptr_t>(two_rdm) & 15) == 0);
    const int64_t per_block = vec2 ? 512 : 256;
    dim3 grid(static_cast<unsigned>((L + per_block - 1) / per_block), nsplit, (nbatch + kGB - 1) / kGB);
    EVC_REQUIRE(grid.z <= 65535, "evc_predict_rdm: batch too large for one launch");
    if (vec2) stack_axpy_kernel<true><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, w, nbatch, nsplit, part);
    else stack_axpy_kernel<false><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, w, nbatch, nsplit, part);
    EVC_CHECK_LAUNCH();
  }
  if (!gemm || exch) {
    dim3 grid(static_cast<unsigned>((n4 + 255) / 256), nbatch);
    gamma2_finalize_kernel<<<grid, 256, 0, ctx->stream>>>(n2, exch ? 1 : 0, L, nbatch, nsplit, part, Gamma);
    EVC_CHECK_LAUNCH();
  }
  return 0;
}

}  // extern "C"
