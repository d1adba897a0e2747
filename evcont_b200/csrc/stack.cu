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

// K5 streaming form (few geometries, stack larger than any tensor-core tile can
// amortise): partial[g][p][chunk] = sum_{l in chunk} R2[p][l] * hv[g][l].
// A CTA owns RB consecutive pair rows x one chunk of kChunk stack columns and GB
// geometries.  Per iteration a thread issues RB independent 16-byte loads of the
// stack (2 iterations unrolled => 2*RB loads in flight per thread, ~64 KB per CTA)
// and re-uses the hv vector it holds in registers for all RB rows, so the stack is
// read exactly once from HBM and hv is re-read only P/RB times from L2.
template <int RB, int GB, bool VEC2>
__global__ void __launch_bounds__(256)
stack_dot_kernel(const double* __restrict__ R2, int64_t L, int P, const double* __restrict__ hv,
                 int G, int nchunk, double* __restrict__ partial) {
  __shared__ double red[8][RB * GB];
  const int p0 = blockIdx.x * RB, ch = blockIdx.y, g0 = blockIdx.z * GB;
  const int64_t lo = static_cast<int64_t>(ch) * kChunk;
  const int64_t hi = min(L, lo + kChunk);
  const double* row[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) row[r] = R2 + static_cast<int64_t>(min(p0 + r, P - 1)) * L;
  const double* hg[GB];
#pragma unroll
  for (int q = 0; q < GB; ++q) hg[q] = hv + static_cast<int64_t>(min(g0 + q, G - 1)) * L;
  double acc[RB][GB];
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int q = 0; q < GB; ++q) acc[r][q] = 0.0;
  if (VEC2) {
#pragma unroll 2
    for (int64_t l = lo + 2 * threadIdx.x; l < hi; l += 512) {
      double2 rv[RB], h[GB];
#pragma unroll
      for (int r = 0; r < RB; ++r) rv[r] = __ldg(reinterpret_cast<const double2*>(row[r] + l));
#pragma unroll
      for (int q = 0; q < GB; ++q) h[q] = __ldg(reinterpret_cast<const double2*>(hg[q] + l));
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int q = 0; q < GB; ++q) {
          acc[r][q] = fma(rv[r].x, h[q].x, acc[r][q]);
          acc[r][q] = fma(rv[r].y, h[q].y, acc[r][q]);
        }
    }
  } else {
#pragma unroll 2
    for (int64_t l = lo + threadIdx.x; l < hi; l += 256) {
      double rv[RB], h[GB];
#pragma unroll
      for (int r = 0; r < RB; ++r) rv[r] = __ldg(row[r] + l);
#pragma unroll
      for (int q = 0; q < GB; ++q) h[q] = __ldg(hg[q] + l);
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int q = 0; q < GB; ++q) acc[r][q] = fma(rv[r], h[q], acc[r][q]);
    }
  }
  // fixed-order reduction: lanes (butterfly), then warps 0..7
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int q = 0; q < GB; ++q) {
      double v = acc[r][q];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][r * GB + q] = v;
    }
  __syncthreads();
  if (threadIdx.x < RB * GB) {
    const int r = threadIdx.x / GB, q = threadIdx.x - r * GB;
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w][threadIdx.x];
    if (p0 + r < P && g0 + q < G)
      partial[(static_cast<int64_t>(g0 + q) * P + p0 + r) * nchunk + ch] = tot;
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
// partial element (g, p, c) lives at partial[g*sg + p*sp + c*sc].  One WARP per (a, b): the lanes
// run over the n^2 one-body elements (coalesced) and over the chunks; fixed butterfly reduction.
__global__ void assemble_H_kernel(int N, int n2, int tril, double scale, int P, int nchunk,
                                  int64_t sg, int64_t sp, int64_t sc,
                                  const double* __restrict__ one_rdm, const double* __restrict__ h1,
                                  const double* __restrict__ partial, double* __restrict__ H) {
  const int g = blockIdx.y, lane = threadIdx.x & 31;
  const int ab = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ab >= N * N) return;
  const int a = ab / N, b = ab - a * N;
  const double* r1 = one_rdm + static_cast<int64_t>(ab) * n2;
  const double* hg = h1 + static_cast<int64_t>(g) * n2;
  double one = 0.0;
  for (int k = lane; k < n2; k += 32) one = fma(r1[k], hg[k], one);
  double two = 0.0;
  int p = -1;
  if (!tril) p = ab;
  else if (a >= b) p = a * (a + 1) / 2 + b;
  if (p >= 0) {
    const double* pp = partial + g * sg + p * sp;
    for (int c = lane; c < nchunk; c += 32) two += pp[c * sc];
  }
  double v = one + scale * two;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) H[static_cast<int64_t>(g) * N * N + ab] = v;
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

// gamma[g][pq] = sum_ab c_a c_b one_rdm[ab][pq].  One CTA per (32 columns, geometry): the 16 warps
// walk the N^2 rows with stride 16 (coalesced 256-byte reads, 8 in flight per warp) and are combined
// in a fixed order, so the one-body stack (62.7 MB at N = 100, n = 28) streams instead of being walked
// by n^2 threads one row at a time.
constexpr int kG1Warps = 16;
__global__ void __launch_bounds__(kG1Warps * 32)
gamma1_kernel(int N, int n2, const double* __restrict__ one_rdm, const double* __restrict__ C, int64_t c_stride,
              double* __restrict__ gamma) {
  __shared__ double red[kG1Warps][33];
  extern __shared__ double cs[];  // c[N]
  const int g = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 32 + lane;
  const double* c = C + static_cast<int64_t>(g) * c_stride;
  for (int a = threadIdx.x; a < N; a += kG1Warps * 32) cs[a] = c[a];
  __syncthreads();
  const int P = N * N;
  const bool live = k < n2;
  const double* col = one_rdm + (live ? k : 0);
  double acc = 0.0;
  int ab = warp;
  for (; ab + 7 * kG1Warps < P; ab += 8 * kG1Warps) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(col + static_cast<int64_t>(ab + u * kG1Warps) * n2);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = ab + u * kG1Warps, a = r / N, b = r - a * N;
      acc = fma(cs[a] * cs[b], v[u], acc);
    }
  }
  for (; ab < P; ab += kG1Warps) {
    const int a = ab / N, b = ab - a * N;
    acc = fma(cs[a] * cs[b], __ldg(col + static_cast<int64_t>(ab) * n2), acc);
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && live) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kG1Warps; ++w) t += red[w][lane];
    gamma[static_cast<int64_t>(g) * n2 + k] = t;
  }
}

// K7 streaming form: part[split][g][l] = sum_{p in split} w[g][p] R2[p][l].
// A thread owns two adjacent stack columns and walks the pair rows of its split with
// 8 independent 16-byte loads in flight; the weights of the split are staged in
// shared memory (read as broadcasts).  Grid sized by axpy_nsplit() to >= 16 CTAs/SM.
constexpr int kAxpyWTile = 512;  // weights staged per pass: GB * 512 doubles <= 32 KB

template <int GB, bool VEC2>
__global__ void __launch_bounds__(256)
stack_axpy_kernel(const double* __restrict__ R2, int64_t L, int P, const double* __restrict__ w,
                  int G, int nsplit, double* __restrict__ part) {
  __shared__ double ws[GB][kAxpyWTile];
  const int split = blockIdx.y, g0 = blockIdx.z * GB;
  const int p0 = static_cast<int>(static_cast<int64_t>(P) * split / nsplit);
  const int p1 = static_cast<int>(static_cast<int64_t>(P) * (split + 1) / nsplit);
  constexpr int W = VEC2 ? 2 : 1;
  const int64_t l = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) * W;
  const bool live = l < L;
  const double* col = R2 + (live ? l : 0);
  double acc[GB][W];
#pragma unroll
  for (int q = 0; q < GB; ++q)
#pragma unroll
    for (int e = 0; e < W; ++e) acc[q][e] = 0.0;
  for (int pt = p0; pt < p1; pt += kAxpyWTile) {
    const int cnt = min(kAxpyWTile, p1 - pt);
    __syncthreads();
    for (int k = threadIdx.x; k < GB * cnt; k += 256) {
      const int q = k / cnt, j = k - q * cnt;
      ws[q][j] = (g0 + q < G) ? __ldg(w + static_cast<int64_t>(g0 + q) * P + pt + j) : 0.0;
    }
    __syncthreads();
    if (!live) continue;
    int j = 0;
    for (; j + 8 <= cnt; j += 8) {
      if (VEC2) {
        double2 rv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          rv[u] = __ldg(reinterpret_cast<const double2*>(col + static_cast<int64_t>(pt + j + u) * L));
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int q = 0; q < GB; ++q) {
            const double wv = ws[q][j + u];
            acc[q][0] = fma(wv, rv[u].x, acc[q][0]);
            acc[q][W - 1] = fma(wv, rv[u].y, acc[q][W - 1]);
          }
      } else {
        double rv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) rv[u] = __ldg(col + static_cast<int64_t>(pt + j + u) * L);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int q = 0; q < GB; ++q) acc[q][0] = fma(ws[q][j + u], rv[u], acc[q][0]);
      }
    }
    for (; j < cnt; ++j) {
      if (VEC2) {
        const double2 rv = __ldg(reinterpret_cast<const double2*>(col + static_cast<int64_t>(pt + j) * L));
#pragma unroll
        for (int q = 0; q < GB; ++q) {
          acc[q][0] = fma(ws[q][j], rv.x, acc[q][0]);
          acc[q][W - 1] = fma(ws[q][j], rv.y, acc[q][W - 1]);
        }
      } else {
        const double rv = __ldg(col + static_cast<int64_t>(pt + j) * L);
#pragma unroll
        for (int q = 0; q < GB; ++q) acc[q][0] = fma(ws[q][j], rv, acc[q][0]);
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int q = 0; q < GB; ++q) {
    if (g0 + q >= G) continue;
    double* dst = part + (static_cast<int64_t>(split) * G + g0 + q) * L + l;
    if (VEC2) *reinterpret_cast<double2*>(dst) = make_double2(acc[q][0], acc[q][W - 1]);
    else dst[0] = acc[q][0];
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
constexpr int kGemvMaxBatch = 16;

// geometries per register tile of the streaming kernels
inline int stream_gb(int G) { return G <= 1 ? 1 : G <= 2 ? 2 : G <= 4 ? 4 : 8; }
inline int stream_rb(int gb) { return gb <= 2 ? 8 : 4; }
constexpr int kPlanSms = 148;  // B200; the plan must not depend on the ctx (workspace sizing)
constexpr int kK5BM = 128, kK5BN = 80;

evc_gemm::Plan k5_plan(int G, int P, int64_t L) {
  const int tiles = ((G + kK5BM - 1) / kK5BM) * ((P + kK5BN - 1) / kK5BN);
  return evc_gemm::plan_split(tiles, static_cast<int>(L), kPlanSms, 64);
}

int axpy_nsplit(int64_t L, int P, int G) {
  const int gb = stream_gb(G);
  const int64_t blocks = ((L + 511) / 512) * ((G + gb - 1) / gb);
  int64_t s = (16 * 148 + blocks - 1) / blocks;
  if (s > 32) s = 32;
  if (s > P) s = P;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}


// out[g][k] = sum_{z < nz} part[z * sz + g * sg + k * sk]     (fixed order)
__global__ void reduce_slabs_kernel(int64_t cols, int nz, int64_t sz, int64_t sg, int64_t sk,
                                    const double* __restrict__ part, double* __restrict__ out) {
  const int g = blockIdx.y;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= cols) return;
  const double* p = part + g * sg + k * sk;
  double acc = 0.0;
  for (int z = 0; z < nz; ++z) acc += p[z * sz];
  out[static_cast<int64_t>(g) * cols + k] = acc;
}

constexpr int kRdBM = 128, kRdBN = 80;   // rows_dot GEMM tile (G x P)
constexpr int kRaBM = 64, kRaBN = 128;   // rows_axpy GEMM tile (G x L)

// column tile of the rows_dot GEMM: 80 (4 x 2 warps) or 72 (8 x 1 warps), whichever pads the P pair columns less
// (N = 20 training states: 210 columns run as 3 x 72 = 216 instead of 3 x 80 = 240)
inline int rows_dot_bn(int P) {
  const int p80 = (P + 79) / 80 * 80, p72 = (P + 71) / 72 * 72;
  return p72 < p80 ? 72 : kRdBN;
}

evc_gemm::Plan rows_dot_plan(int G, int P, int64_t L) {
  const int bn = rows_dot_bn(P);
  const int tiles = ((G + kRdBM - 1) / kRdBM) * ((P + bn - 1) / bn);
  return evc_gemm::plan_split(tiles, static_cast<int>(L), kPlanSms, 128);
}

}  // namespace

// ---- generic row contractions over a row-major matrix rows[P][L] (L % 2 == 0) ----
// rows_dot : out[g][p] = sum_l rows[p][l] * hv[g][l]      (hv: [G][L])
// rows_axpy: out[g][l] = sum_p w[g][p] * rows[p][l]       (w: [G][P])
// Small batches stream the rows once at HBM/L2 speed; larger ones run on the FP64
// tensor cores (DMMA).  Used by the packed prediction step (packed.cu).
size_t evc_rows_dot_ws_bytes(int64_t L, int P, int G) {
  if (G > kGemvMaxBatch) {
    const evc_gemm::Plan pl = rows_dot_plan(G, P, L);
    return evc_align_up(static_cast<size_t>(pl.nsplit) * G * P * 8, 256);
  }
  const int64_t nchunk = (L + kChunk - 1) / kChunk;
  return evc_align_up(static_cast<size_t>(G) * P * nchunk * 8, 256);
}

int evc_rows_dot(evc_ctx* ctx, const double* rows, int64_t L, int P, const double* hv, int G, double* out,
                 void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(L < (int64_t(1) << 31) && L >= 1, "rows_dot: bad row length %lld", (long long)L);
  EVC_REQUIRE(workspace_bytes >= evc_rows_dot_ws_bytes(L, P, G), "rows_dot: workspace too small");
  double* partial = static_cast<double*>(workspace);
  if (G > kGemvMaxBatch) {
    const evc_gemm::Plan pl = rows_dot_plan(G, P, L);
    int rc = rows_dot_bn(P) == 72
                 ? evc_gemm::launch<kRdBM, 72, 8, 1, false>(ctx->stream, G, P, static_cast<int>(L), pl, hv, L, rows, L,
                                                            partial, P, static_cast<int64_t>(G) * P)
                 : evc_gemm::launch<kRdBM, kRdBN, 4, 2, false>(ctx->stream, G, P, static_cast<int>(L), pl, hv, L, rows, L,
                                                               partial, P, static_cast<int64_t>(G) * P);
    if (rc) return rc;
    dim3 grid((P + 127) / 128, G);
    reduce_slabs_kernel<<<grid, 128, 0, ctx->stream>>>(P, pl.nsplit, static_cast<int64_t>(G) * P, P, 1, partial, out);
    EVC_CHECK_LAUNCH();
    return 0;
  }
  const int nchunk = static_cast<int>((L + kChunk - 1) / kChunk);
  const int gb = stream_gb(G), rb = stream_rb(gb);
  dim3 grid((P + rb - 1) / rb, nchunk, (G + gb - 1) / gb);
  EVC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "rows_dot: batch/stack too large for one launch");
  const bool vec2 = (L % 2 == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(hv) & 15) == 0);
#define EVC_DOT(RB, GB)                                                                                  \
  do {                                                                                                   \
    if (vec2) stack_dot_kernel<RB, GB, true><<<grid, 256, 0, ctx->stream>>>(rows, L, P, hv, G, nchunk, partial); \
    else stack_dot_kernel<RB, GB, false><<<grid, 256, 0, ctx->stream>>>(rows, L, P, hv, G, nchunk, partial);     \
  } while (0)
  switch (gb) {
    case 1: EVC_DOT(8, 1); break;
    case 2: EVC_DOT(8, 2); break;
    case 4: EVC_DOT(4, 4); break;
    default: EVC_DOT(4, 8); break;
  }
#undef EVC_DOT
  EVC_CHECK_LAUNCH();
  {
    dim3 g2((P + 127) / 128, G);
    reduce_slabs_kernel<<<g2, 128, 0, ctx->stream>>>(P, nchunk, 1, static_cast<int64_t>(P) * nchunk, nchunk, partial, out);
    EVC_CHECK_LAUNCH();
  }
  return 0;
}

size_t evc_rows_axpy_ws_bytes(int64_t L, int P, int G) {
  if (G > kGemvMaxBatch) return 256;
  return evc_align_up(static_cast<size_t>(axpy_nsplit(L, P, G)) * G * L * 8, 256);
}

int evc_rows_axpy(evc_ctx* ctx, const double* rows, int64_t L, int P, const double* w, int G, double* out,
                  void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(L < (int64_t(1) << 31) && L >= 1, "rows_axpy: bad row length %lld", (long long)L);
  EVC_REQUIRE(workspace_bytes >= evc_rows_axpy_ws_bytes(L, P, G), "rows_axpy: workspace too small");
  if (G > kGemvMaxBatch) {
    evc_gemm::Plan pl;
    pl.nsplit = 1;
    pl.kchunk = (P + evc_gemm::BK - 1) / evc_gemm::BK * evc_gemm::BK;
    return evc_gemm::launch<kRaBM, kRaBN, 2, 4, true>(ctx->stream, G, static_cast<int>(L), P, pl, w, P, rows, L, out, L, 0);
  }
  double* part = static_cast<double*>(workspace);
  const int nsplit = axpy_nsplit(L, P, G);
  const bool vec2 = (L % 2 == 0) && (reinterpret_cast<uintptr_t>(rows) & 15) == 0;
  const int64_t per_block = vec2 ? 512 : 256;
  const int gb = stream_gb(G);
  dim3 grid(static_cast<unsigned>((L + per_block - 1) / per_block), nsplit, (G + gb - 1) / gb);
  EVC_REQUIRE(grid.z <= 65535, "rows_axpy: batch too large for one launch");
#define EVC_AXPY(GB)                                                                                  \
  do {                                                                                                \
    if (vec2) stack_axpy_kernel<GB, true><<<grid, 256, 0, ctx->stream>>>(rows, L, P, w, G, nsplit, part); \
    else stack_axpy_kernel<GB, false><<<grid, 256, 0, ctx->stream>>>(rows, L, P, w, G, nsplit, part);     \
  } while (0)
  switch (gb) {
    case 1: EVC_AXPY(1); break;
    case 2: EVC_AXPY(2); break;
    case 4: EVC_AXPY(4); break;
    default: EVC_AXPY(8); break;
  }
#undef EVC_AXPY
  EVC_CHECK_LAUNCH();
  {
    dim3 g2(static_cast<unsigned>((L + 255) / 256), G);
    reduce_slabs_kernel<<<g2, 256, 0, ctx->stream>>>(L, nsplit, static_cast<int64_t>(G) * L, L, 1, part, out);
    EVC_CHECK_LAUNCH();
  }
  return 0;
}

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
    const int gb = stream_gb(nbatch), rb = stream_rb(gb);
    dim3 grid((P + rb - 1) / rb, nchunk, (nbatch + gb - 1) / gb);
    EVC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "evc_subspace_H: batch/stack too large for one launch");
    const bool vec2 = (L % 2 == 0) && ((reinterpret_cast<uintptr_t>(two_rdm) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(hv) & 15) == 0);
#define EVC_DOT(RB, GB)                                                                                         \
  do {                                                                                                          \
    if (vec2) stack_dot_kernel<RB, GB, true><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, hv, nbatch, nchunk, partial); \
    else stack_dot_kernel<RB, GB, false><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, hv, nbatch, nchunk, partial);     \
  } while (0)
    switch (gb) {
      case 1: EVC_DOT(8, 1); break;
      case 2: EVC_DOT(8, 2); break;
      case 4: EVC_DOT(4, 4); break;
      default: EVC_DOT(4, 8); break;
    }
#undef EVC_DOT
    EVC_CHECK_LAUNCH();
  }
  {
    dim3 grid((N * N + 3) / 4, nbatch);  // 4 warps per CTA, one warp per (a, b)
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
    dim3 grid((n2 + 31) / 32, nbatch);
    gamma1_kernel<<<grid, kG1Warps * 32, static_cast<size_t>(N) * sizeof(double), ctx->stream>>>(N, n2, one_rdm, C,
                                                                                              c_stride, gamma);
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
    const int gb = stream_gb(nbatch);
    dim3 grid(static_cast<unsigned>((L + per_block - 1) / per_block), nsplit, (nbatch + gb - 1) / gb);
    EVC_REQUIRE(grid.z <= 65535, "evc_predict_rdm: batch too large for one launch");
#define EVC_AXPY(GB)                                                                                           \
  do {                                                                                                         \
    if (vec2) stack_axpy_kernel<GB, true><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, w, nbatch, nsplit, part); \
    else stack_axpy_kernel<GB, false><<<grid, 256, 0, ctx->stream>>>(two_rdm, L, P, w, nbatch, nsplit, part);     \
  } while (0)
    switch (gb) {
      case 1: EVC_AXPY(1); break;
      case 2: EVC_AXPY(2); break;
      case 4: EVC_AXPY(4); break;
      default: EVC_AXPY(8); break;
    }
#undef EVC_AXPY
    EVC_CHECK_LAUNCH();
  }
  if (!gemm || exch) {
    dim3 grid(static_cast<unsigned>((n4 + 255) / 256), nbatch);
    gamma2_finalize_kernel<<<grid, 256, 0, ctx->stream>>>(n2, exch ? 1 : 0, L, nbatch, nsplit, part, Gamma);
    EVC_CHECK_LAUNCH();
  }
  return 0;
}

// ---- e3: the row operations of K5 / K7 on a SLAB of training pairs (pair-sharded stack) ----------------
int evc_exchange_compress(evc_ctx* ctx, int n, int nbatch, const double* h2, double* h2c) {
  EVC_REQUIRE(ctx && h2 && h2c && n >= 1, "evc_exchange_compress: bad argument");
  if (nbatch <= 0) return 0;
  const int64_t L = exch_len(n);
  dim3 grid(static_cast<unsigned>((L + 255) / 256), nbatch);
  pack_exchange_kernel<<<grid, 256, 0, ctx->stream>>>(n * n, L, h2, h2c);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_exchange_restore(evc_ctx* ctx, int n, int nbatch, const double* gc, double* Gamma) {
  EVC_REQUIRE(ctx && gc && Gamma && n >= 1, "evc_exchange_restore: bad argument");
  if (nbatch <= 0) return 0;
  const int64_t n4 = static_cast<int64_t>(n) * n * n * n;
  dim3 grid(static_cast<unsigned>((n4 + 255) / 256), nbatch);
  gamma2_finalize_kernel<<<grid, 256, 0, ctx->stream>>>(n * n, 1, exch_len(n), nbatch, 1, gc, Gamma);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_stack_rows_workspace_bytes(int64_t row_len, int nrows, int nbatch, size_t* bytes) {
  EVC_REQUIRE(bytes && row_len >= 2 && nrows >= 1 && nbatch >= 1, "evc_stack_rows_workspace_bytes: bad argument");
  const size_t a = evc_rows_dot_ws_bytes(row_len, nrows, nbatch), b = evc_rows_axpy_ws_bytes(row_len, nrows, nbatch);
  *bytes = a > b ? a : b;
  return 0;
}

int evc_stack_rows_dot(evc_ctx* ctx, const double* rows, int64_t row_len, int nrows, const double* hv, int nbatch,
                       double* out, void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && rows && hv && out && workspace, "evc_stack_rows_dot: NULL argument");
  if (nbatch <= 0 || nrows <= 0) return 0;
  return evc_rows_dot(ctx, rows, row_len, nrows, hv, nbatch, out, workspace, workspace_bytes);
}

int evc_stack_rows_axpy(evc_ctx* ctx, const double* rows, int64_t row_len, int nrows, const double* w, int nbatch,
                        double* out, void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && rows && w && out && workspace, "evc_stack_rows_axpy: NULL argument");
  if (nbatch <= 0 || nrows <= 0) return 0;
  return evc_rows_axpy(ctx, rows, row_len, nrows, w, nbatch, out, workspace, workspace_bytes);
}

}  // extern "C"
