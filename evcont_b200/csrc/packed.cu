// Packed prediction step: the same arithmetic as K4/K5/K7/K8 (dense.cu, stack.cu,
// grad.cu), restated on the 8-fold permutational symmetry of the AO two-electron
// integrals, (ij|kl) = (ji|kl) = (ij|lk) = (kl|ij), which libcint / PySCF
// integrals always have (evcont/ab_initio_gradients_loewdin.py:283-284, 338-339).
//
// Every contraction of the stack with h2, and every use of the predicted two-body
// RDM in the gradient, only sees the part of the RDM that is totally symmetric
// under that 8-element group.  So the stack is packed ONCE into rows
//
//   R[p][0 : n^2]             = one_rdm block           (full, no symmetry assumed)
//   R[p][n^2 + tri(I, K)]     = 1/2 * sum_{distinct (i'j'k'l') in orbit(ijkl)} two_rdm[i'j'k'l']
//                               I = pair(i >= j), K = pair(k >= l), I >= K
//
// for the N(N+1)/2 state pairs a >= b (the eigensolver reads the lower triangle of
// H only, evcont/ab_initio_eigenvector_continuation.py:75).  Row length
// L8 = n^2 + np(np+1)/2, np = n(n+1)/2: 1640 doubles instead of 10100 at n = 10, and
// 210 rows instead of 400 at N = 20 -- 12x fewer bytes and flops in K5 and K7.
//
//   K4p  per geometry: ERIp (np x np) -> T = ERIp Q -> h2p = Q^T T,  Q = pair transform
//        built from X; two DMMA GEMMs in shared memory.   hvec = [h1 | tril(h2p)]
//   K5p  Hp[g][p]   = hvec[g] . RH[p]            (rows_dot: DMMA GEMM / streaming)
//   K7p  out7[g][:] = sum_p w[g][p] RG[p][:]     (rows_axpy), w = tril weights of c (x) c
//   K8p  per geometry: U0 = T Gm, Y from U0; W = P0 Gm P0^T (three DMMA GEMMs in shared
//        memory); streams int2e_ip1 once against W; one-electron adjoint as in grad.cu.
//
// Algebra (derivation and numpy check: DESIGN.md section 4, tools/packed_proto.py):
//   P0[AB, I] = X_ai X_bj + X_aj X_bi,  Q = diag(1/s_AB) P0,  s = 2 on diagonal pairs
//   Gm[I, K]  = out7[tri(I,K)]   (symmetric np x np; RG rows carry the diagonal pairs doubled)
//   Y[a, i]   = 2 sum_{b j} X_bj s_(ij) U0[(ab), (ij)]
//   -1/2 sum_{m in A} sum_{bcd} (d_x m b|c d) W[(mb), (cd)]   is the ERI-derivative term.
#include <algorithm>

#include "packed.cuh"

// Optional phase timing of the per-geometry kernels (development aid): build with
// -DEVC_PHASE_TIMING, read with evc_debug_phase_clocks().
#ifdef EVC_PHASE_TIMING
__device__ long long g_evc_phase[4][24];
#define EVC_PHASE(slot, idx)                                                             \
  do {                                                                                   \
    __syncthreads();                                                                     \
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 700))                       \
      g_evc_phase[(slot) * 2 + (blockIdx.x ? 1 : 0)][idx] = clock64();                    \
  } while (0)
#define EVC_MARK(slot, idx, cond)                                                        \
  do {                                                                                   \
    if ((cond) && (threadIdx.x & 31) == 0 && (blockIdx.x == 0 || blockIdx.x == 700))      \
      g_evc_phase[(slot) * 2 + (blockIdx.x ? 1 : 0)][idx] = clock64();                    \
  } while (0)
#else
#define EVC_PHASE(slot, idx) do { } while (0)
#define EVC_MARK(slot, idx, cond) do { } while (0)
#endif

using namespace evcp;

namespace {

constexpr int kThreads = 256;      // K4p CTA
constexpr int kGradThreads = 512;  // K8a CTA
constexpr int kNJ = 4;  // 8x8 output tiles per warp work item (A fragment re-used kNJ times)

// ---------------------------------------------------------------------------
// stack packing (once per stack)
// ---------------------------------------------------------------------------
struct StackView {
  const double* two;
  int layout, N, n;
  int64_t n4, Lc;
  __device__ bool has_block(int a, int b) const {
    return (layout == EVC_LAYOUT_FULL || layout == EVC_LAYOUT_FULL_EXCH) || a >= b;
  }
  __device__ double at(int a, int b, int i, int j, int k, int l) const {
    const bool tril = (layout == EVC_LAYOUT_TRIL || layout == EVC_LAYOUT_TRIL_EXCH);
    const bool exch = (layout == EVC_LAYOUT_FULL_EXCH || layout == EVC_LAYOUT_TRIL_EXCH);
    const int64_t blk = tril ? tri_idx(a, b) : static_cast<int64_t>(a) * N + b;
    const int64_t x = i * n + j, y = k * n + l;
    if (exch) {
      const int64_t hi = x > y ? x : y, lo = x > y ? y : x;
      return two[blk * Lc + hi * (hi + 1) / 2 + lo];
    }
    return two[blk * n4 + x * (static_cast<int64_t>(n) * n) + y];
  }
  __device__ double orbit_sum(int a, int b, int i, int j, int k, int l, bool same_pair) const {
    double s = at(a, b, i, j, k, l);
    if (i != j) s += at(a, b, j, i, k, l);
    if (k != l) s += at(a, b, i, j, l, k);
    if (i != j && k != l) s += at(a, b, j, i, l, k);
    if (!same_pair) {
      s += at(a, b, k, l, i, j);
      if (k != l) s += at(a, b, l, k, i, j);
      if (i != j) s += at(a, b, k, l, j, i);
      if (i != j && k != l) s += at(a, b, l, k, j, i);
    }
    return s;
  }
};

__global__ void pack8_stack_kernel(StackView sv, int64_t L8, const double* __restrict__ one_rdm,
                                   double* __restrict__ RH, double* __restrict__ RG) {
  const int p = blockIdx.y;
  const int64_t col = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (col >= L8) return;
  int a, b;
  tril_unrank_i(p, a, b);
  const int n = sv.n, N = sv.N, n2 = n * n, np = npair_of(n);
  double vh = 0.0, vg = 0.0;
  if (col < n2) {
    vh = one_rdm[(static_cast<int64_t>(a) * N + b) * n2 + col];
    vg = (a == b) ? vh : 0.5 * (vh + one_rdm[(static_cast<int64_t>(b) * N + a) * n2 + col]);
  } else if (col - n2 < static_cast<int64_t>(np) * (np + 1) / 2) {
    int I, K, i, j, k, l;
    tril_unrank_i(static_cast<int>(col - n2), I, K);
    tril_unrank_i(I, i, j);
    tril_unrank_i(K, k, l);
    vh = 0.5 * sv.orbit_sum(a, b, i, j, k, l, I == K);
    vg = vh;
    if (a != b && sv.has_block(b, a)) vg = 0.5 * (vh + 0.5 * sv.orbit_sum(b, a, i, j, k, l, I == K));
    if (I == K) vg *= 2.0;  // RG carries Gm directly: diagonal pair entries doubled
  }
  RH[static_cast<int64_t>(p) * L8 + col] = vh;
  RG[static_cast<int64_t>(p) * L8 + col] = vg;
}

// hvec[g] = [h1 | tril(h2 pairs)] from the full OAO integrals (n > kPackedMaxNorb path)
__global__ void hvec_from_full_kernel(int n, int64_t L8, const double* __restrict__ h1,
                                      const double* __restrict__ h2, double* __restrict__ hvec) {
  const int g = blockIdx.y;
  const int64_t col = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (col >= L8) return;
  const int n2 = n * n, np = npair_of(n);
  const int64_t n4 = static_cast<int64_t>(n2) * n2;
  double v = 0.0;
  if (col < n2) {
    v = h1[static_cast<int64_t>(g) * n2 + col];
  } else if (col - n2 < static_cast<int64_t>(np) * (np + 1) / 2) {
    int I, K, i, j, k, l;
    tril_unrank_i(static_cast<int>(col - n2), I, K);
    tril_unrank_i(I, i, j);
    tril_unrank_i(K, k, l);
    v = h2[static_cast<int64_t>(g) * n4 + (static_cast<int64_t>(i) * n + j) * n2 + k * n + l];
  }
  hvec[static_cast<int64_t>(g) * L8 + col] = v;
}

// gamma = out7[0:n^2]; Gamma8[ijkl] = s_I s_K s_IK out7[tri(I,K)] / 4: the totally
// symmetric part of the predicted two-body RDM (all the gradient sees).  One CTA per (i, j) and
// geometry, threads over (k, l): the pair index of (i, j) is uniform and the writes are contiguous.
__global__ void __launch_bounds__(256)
unpack_rdms_kernel(int n, int64_t L8, const double* __restrict__ out7, double* __restrict__ gamma,
                   double* __restrict__ Gamma8) {
  const int g = blockIdx.y, ij = blockIdx.x;
  const int i = ij / n, j = ij - i * n;
  const int n2 = n * n;
  const double* o = out7 + static_cast<int64_t>(g) * L8;
  const int I = i >= j ? tri_idx(i, j) : tri_idx(j, i);
  const double si = (i == j) ? 0.5 : 0.25;  // s_I / 4 (the I == K factor is in RG)
  double* dst = Gamma8 + (static_cast<int64_t>(g) * n2 + ij) * n2;
  for (int t = threadIdx.x; t < n2; t += 256) {
    const int k = t / n, l = t - k * n;
    const int K = k >= l ? tri_idx(k, l) : tri_idx(l, k);
    const int hi = I > K ? I : K, lo = I > K ? K : I;
    dst[t] = (k == l ? 2.0 : 1.0) * si * __ldg(o + n2 + tri_idx(hi, lo));
  }
  if (threadIdx.x == 0) gamma[static_cast<int64_t>(g) * n2 + ij] = o[ij];
}

// w[g][p] = c_a^2 (a == b), 2 c_a c_b (a > b)      (ab_initio_gradients_loewdin.py:345-353)
__global__ void tril_weights_kernel(int N, int P, const double* __restrict__ C, int64_t c_stride,
                                    double* __restrict__ w) {
  const int g = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double* c = C + static_cast<int64_t>(g) * c_stride;
  int a, b;
  tril_unrank_i(p, a, b);
  w[static_cast<int64_t>(g) * P + p] = (a == b) ? c[a] * c[a] : 2.0 * c[a] * c[b];
}

// ---------------------------------------------------------------------------
// shared-memory geometry of the per-geometry kernels
// ---------------------------------------------------------------------------
// Work items of a CTA-level GEMM, C(M8*8 x N8*8) = A * B on the FP64 tensor cores:
// item = (row tile, group of kNJ column tiles).  DMMA issue is per SM sub-partition
// (warp % 4), so the host deals the items to warps such that the four sub-partitions
// carry the same number of tiles (longest-processing-time greedy); the two
// least-loaded warps also run the one-electron chain of K8a.
constexpr int kMaxWarps = 16, kMaxSlots = 5;
struct ItemMap {
  signed char it[kMaxWarps][kMaxSlots];
  unsigned char w1, w2;
};

template <int MAXI, typename LA, typename LB>
__device__ __forceinline__ void gemm_acc(double (&acc)[MAXI][kNJ][2], const ItemMap& map, int M8, int N8, int K4,
                                         bool lower, LA la, LB lb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int NG = (N8 + kNJ - 1) / kNJ;
#pragma unroll
  for (int r = 0; r < MAXI; ++r) {
#pragma unroll
    for (int j = 0; j < kNJ; ++j) acc[r][j][0] = acc[r][j][1] = 0.0;
    const int it = map.it[warp][r];
    if (it < 0) continue;
    const int mt = it / NG, nt0 = (it - mt * NG) * kNJ;
    int ntend = min(N8, nt0 + kNJ);
    if (lower) ntend = min(ntend, mt + 1);
    if (nt0 >= ntend) continue;
    // fragments of step k0 + 4 are fetched while the DMMAs of step k0 issue
    double a = la(mt * 8 + g, tg), b[kNJ];
#pragma unroll
    for (int j = 0; j < kNJ; ++j) b[j] = (nt0 + j < ntend) ? lb(tg, (nt0 + j) * 8 + g) : 0.0;
    for (int k0 = 0; k0 < K4; k0 += 4) {
      const int k1 = (k0 + 4 < K4) ? k0 + 4 : k0;
      const double an = la(mt * 8 + g, k1 + tg);
      double bn[kNJ];
#pragma unroll
      for (int j = 0; j < kNJ; ++j) bn[j] = (nt0 + j < ntend) ? lb(k1 + tg, (nt0 + j) * 8 + g) : 0.0;
#pragma unroll
      for (int j = 0; j < kNJ; ++j)
        if (nt0 + j < ntend) dmma8x8x4(acc[r][j][0], acc[r][j][1], a, b[j]);
      a = an;
#pragma unroll
      for (int j = 0; j < kNJ; ++j) b[j] = bn[j];
    }
  }
}

// st(row, col, v0, v1): accumulator pair for (row, col) and (row, col + 1)
template <int MAXI, typename ST>
__device__ __forceinline__ void gemm_store(const double (&acc)[MAXI][kNJ][2], const ItemMap& map, int M8, int N8,
                                           bool lower, ST st) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int NG = (N8 + kNJ - 1) / kNJ;
#pragma unroll
  for (int r = 0; r < MAXI; ++r) {
    const int it = map.it[warp][r];
    if (it < 0) continue;
    const int mt = it / NG, nt0 = (it - mt * NG) * kNJ;
    int ntend = min(N8, nt0 + kNJ);
    if (lower) ntend = min(ntend, mt + 1);
#pragma unroll
    for (int j = 0; j < kNJ; ++j)
      if (nt0 + j < ntend) st(mt * 8 + g, (nt0 + j) * 8 + tg * 2, acc[r][j][0], acc[r][j][1]);
  }
}

// pair tables: pij[I] = i | j << 8 ; pidx[i*n + j] = pair index of (i, j)
template <int NT>
__device__ __forceinline__ void build_pair_tables(int n, unsigned short* pij, unsigned short* pidx) {
  for (int k = threadIdx.x; k < n * n; k += NT) {
    const int i = k / n, j = k - i * n;
    const int I = i >= j ? tri_idx(i, j) : tri_idx(j, i);
    pidx[k] = static_cast<unsigned short>(I);
    if (i >= j) pij[I] = static_cast<unsigned short>(i | (j << 8));
  }
}

__host__ __device__ inline size_t table_bytes(int n) {
  return ((static_cast<size_t>(npair_of(n)) + n * n) * sizeof(unsigned short) + 15) / 16 * 16;
}

// ---------------------------------------------------------------------------
// K4p: one CTA per geometry
// ---------------------------------------------------------------------------
template <int MAXI, int NC>
__global__ void __launch_bounds__(kThreads)
packed_ao2oao_kernel(const ItemMap mfull, const ItemMap mlow, int n_rt, int64_t L8, const double* __restrict__ x, const double* __restrict__ hcore,
                     const double* __restrict__ erip, double* __restrict__ hvec, double* __restrict__ Tout) {
  extern __shared__ __align__(16) double sm[];
  const int n = NC > 0 ? NC : n_rt;  // compile-time orbital count for the common sizes
  const PGeom pg = pgeom(n);
  const int np = pg.np, pA = pg.pA, pB = pg.pB, ld = n + 1, n2 = n * n;
  const size_t szE = pg.szA > pg.szB ? pg.szA : pg.szB;
  double* Eb = sm;                 // ERIp, A-type [rows8][pA]; later T, B-type [K4][pB]
  double* Qb = Eb + szE;           // Q, B-type [K4][pB]
  double* Xs = Qb + pg.szB;        // [n][ld]
  double* Hs = Xs + n * ld;
  double* Ts = Hs + n * ld;
  unsigned short* pij = reinterpret_cast<unsigned short*>(Ts + n * ld);
  unsigned short* pidx = pij + np;
  const int g = blockIdx.x, tid = threadIdx.x;
  const int64_t o2 = static_cast<int64_t>(g) * n2;
  const double* eg = erip + static_cast<int64_t>(g) * np * pA;

  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kThreads / 32;
  EVC_PHASE(0, 0);
  // ---- asynchronous copies of this geometry's inputs first ----
  // ERIp[AB][CD] = (ab|cd), a >= b, c >= d, arrives with the shared-memory row pitch: one contiguous block
  for (int k = 2 * tid; k < np * pA; k += 2 * kThreads) cp_async16(Eb + k, eg + k);
  build_pair_tables<kThreads>(n, pij, pidx);  // integer only (no FP64 sqrt: that pipe is the DMMA pipe)
  for (int k = tid; k < n2; k += kThreads) {
    const int i = k / n, j = k - i * n;
    cp_async8(Xs + i * ld + j, x + o2 + k);
    cp_async8(Hs + i * ld + j, hcore + o2 + k);
  }
  cp_async_commit_all();
  {
    // zero padding of Q (B-type); its data region is filled below
    const int padq = pB - np;
    for (int k = tid; k < np * padq; k += kThreads) {
      const int r = k / padq, c = np + (k - r * padq);
      Qb[r * pB + c] = 0.0;
    }
    for (int k = np * pB + tid; k < static_cast<int>(pg.szB); k += kThreads) Qb[k] = 0.0;
  }
  cp_async_wait_all();
  __syncthreads();
  {
    // zero padding of the ERIp image (the array in HBM carries none)
    const int padc = pA - np;
    for (int k = tid; k < np * padc; k += kThreads) {
      const int r = k / padc, c = np + (k - r * padc);
      Eb[r * pA + c] = 0.0;
    }
    for (int k = np * pA + tid; k < static_cast<int>(szE); k += kThreads) Eb[k] = 0.0;
  }
  __syncthreads();
  EVC_PHASE(0, 1);
  // Q[CD][K] = (X_ck X_dl + X_dk X_cl) / s_CD
  for (int CD = warp; CD < np; CD += NW) {
    const int c = pij[CD] & 0xff, d = pij[CD] >> 8;
    const double sc = (c == d) ? 0.5 : 1.0;
    for (int K = lane; K < np; K += 32) {
      const int kk = pij[K] & 0xff, l = pij[K] >> 8;
      Qb[CD * pB + K] = sc * (Xs[c * ld + kk] * Xs[d * ld + l] + Xs[d * ld + kk] * Xs[c * ld + l]);
    }
  }
  // h1 = X^T (hcore X): first half
  for (int k = tid; k < n2; k += kThreads) {
    const int i = k / n, j = k - i * n;
    double acc1 = 0.0;
#pragma unroll
    for (int r = 0; r < (NC > 0 ? NC : n); ++r) acc1 += Hs[i * ld + r] * Xs[r * ld + j];
    Ts[i * ld + j] = acc1;
  }
  __syncthreads();
  EVC_PHASE(0, 2);
  double* hv = hvec + static_cast<int64_t>(g) * L8;
  for (int k = tid; k < n2; k += kThreads) {
    const int i = k / n, j = k - i * n;
    double acc1 = 0.0;
#pragma unroll
    for (int r = 0; r < (NC > 0 ? NC : n); ++r) acc1 += Xs[r * ld + i] * Ts[r * ld + j];
    hv[k] = acc1;
  }
  for (int64_t k = n2 + static_cast<int64_t>(np) * (np + 1) / 2 + tid; k < L8; k += kThreads) hv[k] = 0.0;

  double acc[MAXI][kNJ][2];
  // T = ERIp Q
  gemm_acc<MAXI>(acc, mfull, pg.M8, pg.M8, pg.K4, false,
                 [&](int m, int k) { return Eb[m * pA + k]; },
                 [&](int k, int c) { return Qb[k * pB + c]; });
  __syncthreads();  // every warp is done reading ERIp: T may overwrite it
  EVC_PHASE(0, 3);
  double* Tg = Tout + static_cast<int64_t>(g) * np * pA;  // [np][pA], like the ERIp arrays
  gemm_store<MAXI>(acc, mfull, pg.M8, pg.M8, false, [&](int m, int c, double v0, double v1) {
    if (m < pg.K4) {
      Eb[m * pB + c] = v0;
      Eb[m * pB + c + 1] = v1;
    }
    if (m < np) {
      if (c < np) Tg[m * pA + c] = v0;
      if (c + 1 < np) Tg[m * pA + c + 1] = v1;
    }
  });
  __syncthreads();
  EVC_PHASE(0, 4);
  // h2p = Q^T T, lower triangle only
  gemm_acc<MAXI>(acc, mlow, pg.M8, pg.M8, pg.K4, true,
                 [&](int m, int k) { return Qb[k * pB + m]; },
                 [&](int k, int c) { return Eb[k * pB + c]; });
  gemm_store<MAXI>(acc, mlow, pg.M8, pg.M8, true, [&](int m, int c, double v0, double v1) {
    if (m < np) {
      if (c <= m) hv[n2 + tri_idx(m, c)] = v0;
      if (c + 1 <= m) hv[n2 + tri_idx(m, c + 1)] = v1;
    }
  });
  EVC_PHASE(0, 5);
}

size_t ao2oao_smem_bytes(int n) {
  const PGeom pg = pgeom(n);
  return ((pg.szA > pg.szB ? pg.szA : pg.szB) + pg.szB + 3 * static_cast<size_t>(n) * (n + 1)) * sizeof(double) +
         table_bytes(n);
}

// ---------------------------------------------------------------------------
// K8p: one CTA per geometry
// ---------------------------------------------------------------------------
template <int MAXI, int NC>
__global__ void __launch_bounds__(kGradThreads)
packed_grad_kernel(const ItemMap mfull, const ItemMap mlow, int n_rt, int64_t L8, const double* __restrict__ x, const double* __restrict__ evals,
                   const double* __restrict__ evecs, const double* __restrict__ hcore,
                   const double* __restrict__ Tin, const double* __restrict__ out7,
                   double* __restrict__ Wout, double* __restrict__ OmSout, double* __restrict__ PaoOut) {
  extern __shared__ __align__(16) double sm[];
  const int n = NC > 0 ? NC : n_rt;
  const PGeom pg = pgeom(n);
  const int np = pg.np, pA = pg.pA, pB = pg.pB, ld = n | 1, n2 = n * n;  // odd pitch for the n x n matrices
  double* B1 = sm;               // T (A-type) -> P0 (A-type)
  double* B2 = B1 + pg.szA;      // Gm (A-type) -> W (A-type pitch)
  double* B3 = B2 + pg.szA;      // U0 (pitch pB) -> R (B-type), [K4][pB]
  double* V = B3 + pg.szB;       // small matrices, [n][ld] each
  double* X = V + n * ld;
  double* Hc = X + n * ld;
  double* Gm1 = Hc + n * ld;     // gamma
  double* Z = Gm1 + n * ld;
  double* A = Z + n * ld;
  double* Bm = A + n * ld;
  double* PaoS = Bm + n * ld;    // X gamma X^T
  double* Qh = PaoS + n * ld;    // hcore X (gamma + gamma^T)
  double* Gs = Qh + n * ld;      // divided differences of s^-1/2
  double* rs = Gs + n * ld;      // sqrt(s) or 0
  double* sv = rs + n;           // s
  double* Yp = B1;               // [n2][n] partial sums of Y over b (T is dead by then)
  unsigned short* pij = reinterpret_cast<unsigned short*>(sv + n);
  unsigned short* pidx = pij + np;
  const int g = blockIdx.x, tid = threadIdx.x;
  const int64_t o2 = static_cast<int64_t>(g) * n2;
  const double* o7 = out7 + static_cast<int64_t>(g) * L8;
  constexpr int NT = kGradThreads;

  const int warp = tid >> 5, lane = tid & 31;
  EVC_PHASE(1, 0);
  // ---- all global inputs of this geometry as asynchronous copies, issued first ----
  {
    const double* Tg = Tin + static_cast<int64_t>(g) * np * pA;
    for (int k = 2 * tid; k < np * pA; k += 2 * NT) cp_async16(B1 + k, Tg + k);                    // T ([np][pA])
    for (int r = warp; r < np; r += NT / 32) {
      const int rr = r * (r + 1) / 2;
      for (int c = lane; c < np; c += 32)
        cp_async8(B2 + r * pA + c, o7 + n2 + (r >= c ? rr + c : c * (c + 1) / 2 + r));             // Gm = sym(out7)
    }
    for (int k = tid; k < n2; k += NT) {
      const int i = k / n, j = k - i * n;
      cp_async8(V + i * ld + j, evecs + o2 + k);
      cp_async8(X + i * ld + j, x + o2 + k);
      cp_async8(Hc + i * ld + j, hcore + o2 + k);
      cp_async8(Gm1 + i * ld + j, o7 + k);
    }
    cp_async_commit_all();
  }
  build_pair_tables<NT>(n, pij, pidx);
  for (int k = tid; k < n; k += NT) {
    const double s = evals[static_cast<int64_t>(g) * n + k];
    sv[k] = s;
    rs[k] = s > 1.0e-15 ? sqrt(s) : 0.0;
  }
  cp_async_wait_all();
  __syncthreads();
  {
    // zero padding: columns [np, pA) of the data rows, and the rows [np, rows8) (the T array in HBM has none)
    const int padc = pA - np;
    for (int k = tid; k < np * padc; k += NT) {
      const int r = k / padc, c = np + (k - r * padc);
      B1[r * pA + c] = 0.0;
      B2[r * pA + c] = 0.0;
    }
    for (int k = np * pA + tid; k < pg.rows8 * pA; k += NT) { B1[k] = 0.0; B2[k] = 0.0; }
  }
  __syncthreads();
  EVC_PHASE(1, 2);

  // single-warp n x n product for the one-electron chain on the tensor cores (the FP64
  // FMA pipe is the DMMA datapath: scalar DFMAs would queue behind the GEMM warps), run
  // by a warp that has no (or little) GEMM work
  auto warp_mm = [&](double* C, const double* P, bool tp, const double* Q, bool tq) {
    const int gq = lane >> 2, tq4 = lane & 3;
    const int n8 = (n + 7) >> 3, k4 = (n + 3) & ~3;
    for (int mt = 0; mt < n8; ++mt)
      for (int nt = 0; nt < n8; ++nt) {
        double c0 = 0.0, c1 = 0.0;
        const int row = mt * 8 + gq, col = nt * 8 + gq;
        for (int k0 = 0; k0 < k4; k0 += 4) {
          const int kk = k0 + tq4;
          const double av = (row < n && kk < n) ? (tp ? P[kk * ld + row] : P[row * ld + kk]) : 0.0;
          const double bv = (col < n && kk < n) ? (tq ? Q[col * ld + kk] : Q[kk * ld + col]) : 0.0;
          dmma8x8x4(c0, c1, av, bv);
        }
        const int cc = nt * 8 + tq4 * 2;
        if (row < n) {
          if (cc < n) C[row * ld + cc] = c0;
          if (cc + 1 < n) C[row * ld + cc + 1] = c1;
        }
      }
    __syncwarp();
  };
  const int kW1 = mfull.w1, kW2 = mfull.w2;  // the chain warps (least GEMM work)

  double acc[MAXI][kNJ][2];
  // U0 = T Gm  (Gm symmetric: B(k, c) = Gm[c][k])
  gemm_acc<MAXI>(acc, mfull, pg.M8, pg.M8, pg.K4, false,
                     [&](int m, int k) { return B1[m * pA + k]; },
                     [&](int k, int c) { return B2[c * pA + k]; });
  EVC_MARK(1, 12, warp == 0);
  EVC_MARK(1, 13, warp == 5);
  if (warp == kW2) {          // Pao = X gamma X^T  -> PaoS
    warp_mm(A, X, false, Gm1, false);
    warp_mm(PaoS, A, false, X, true);
    // G_pq = -1/(sqrt(s_p) sqrt(s_q) (sqrt(s_p) + sqrt(s_q))), exact divided differences
    for (int k = lane; k < n2; k += 32) {
      const int p = k / n, q = k - p * n;
      const double rp = rs[p], rq = rs[q];
      double gpq = 0.0;
      if (rp > 0.0 && rq > 0.0) {
        gpq = -1.0 / (rp * rq * (rp + rq));
      } else if ((rp > 0.0) != (rq > 0.0)) {
        const double sp = sv[p], sq = sv[q];
        if (sp != sq) gpq = ((rp > 0.0 ? 1.0 / rp : 0.0) - (rq > 0.0 ? 1.0 / rq : 0.0)) / (sp - sq);
      }
      Gs[p * ld + q] = gpq;
    }
    EVC_MARK(1, 14, true);
  } else if (warp == kW1) {   // Qh = hcore X (gamma + gamma^T)
    for (int k = lane; k < n2; k += 32) {
      const int i = k / n, j = k - i * n;
      Bm[i * ld + j] = Gm1[i * ld + j] + Gm1[j * ld + i];
    }
    __syncwarp();
    warp_mm(Z, X, false, Bm, false);
    warp_mm(Qh, Hc, false, Z, false);
    EVC_MARK(1, 15, true);
  }
  gemm_store<MAXI>(acc, mfull, pg.M8, pg.M8, false, [&](int m, int c, double v0, double v1) {
    if (m < pg.K4) {
      B3[m * pB + c] = v0;
      B3[m * pB + c + 1] = v1;
    }
  });
  __syncthreads();
  EVC_PHASE(1, 3);
  // Yp[(a,i), b] = sum_j X_bj s_(ij) U0[(ab),(ij)];  Y = 2 sum_b Yp  (Yp overwrites T)
  for (int t = tid; t < n2 * n; t += NT) {
    const int ab = t / n, i = t - ab * n;   // lanes run over i: distinct words of one U0 row
    const int a = ab / n, b = ab - a * n;
    const double* urow = B3 + pidx[ab] * pB;
    double sacc = 0.0;
#pragma unroll
    for (int j = 0; j < (NC > 0 ? NC : n); ++j) {
      const double u = urow[pidx[i * n + j]];
      sacc += X[b * ld + j] * (i == j ? 2.0 * u : u);
    }
    Yp[(a * n + i) * n + b] = sacc;
  }
  __syncthreads();
  // Z = Y/2 + hcore X (gamma + gamma^T)
  for (int k = tid; k < n2; k += NT) {
    const int i = k / n, j = k - i * n;
    double t = Qh[i * ld + j];
#pragma unroll
    for (int b = 0; b < (NC > 0 ? NC : n); ++b) t += Yp[k * n + b];
    Z[i * ld + j] = t;
  }
  __syncthreads();
  EVC_PHASE(1, 4);
  // P0[AB][I] = X_ai X_bj + X_aj X_bi  (overwrites the Y partials)
  for (int AB = warp; AB < np; AB += NT / 32) {
    const int a = pij[AB] & 0xff, b = pij[AB] >> 8;
    for (int I = lane; I < np; I += 32) {
      const int i = pij[I] & 0xff, j = pij[I] >> 8;
      B1[AB * pA + I] = X[a * ld + i] * X[b * ld + j] + X[a * ld + j] * X[b * ld + i];
    }
  }
  {
    // re-zero what the Y partials left in the padding of P0
    const int padc = pA - np;
    for (int k = tid; k < np * padc; k += NT) {
      const int r = k / padc, c = np + (k - r * padc);
      B1[r * pA + c] = 0.0;
    }
    for (int k = np * pA + tid; k < pg.rows8 * pA; k += NT) B1[k] = 0.0;
  }
  __syncthreads();
  EVC_PHASE(1, 5);
  // R = Gm P0^T
  gemm_acc<MAXI>(acc, mfull, pg.M8, pg.M8, pg.K4, false,
                     [&](int m, int k) { return B2[m * pA + k]; },
                     [&](int k, int c) { return B1[c * pA + k]; });
  EVC_MARK(1, 16, warp == 0);
  EVC_MARK(1, 17, warp == 5);
  if (warp == kW1) {  // Bm = G o (V^T Z V)
    warp_mm(A, V, true, Z, false);
    warp_mm(Bm, A, false, V, false);
    for (int k = lane; k < n2; k += 32) {
      const int p = k / n, q = k - p * n;
      Bm[p * ld + q] *= Gs[p * ld + q];
    }
    __syncwarp();
    EVC_MARK(1, 18, true);
  }
  // (every warp passed the barrier above after its last read of U0)
  gemm_store<MAXI>(acc, mfull, pg.M8, pg.M8, false, [&](int m, int c, double v0, double v1) {
    if (m < pg.K4) {
      B3[m * pB + c] = v0;
      B3[m * pB + c + 1] = v1;
    }
  });
  __syncthreads();
  EVC_PHASE(1, 6);
  // W = P0 R  (symmetric: lower triangle only, the reader takes (max, min))
  gemm_acc<MAXI>(acc, mlow, pg.M8, pg.M8, pg.K4, true,
                     [&](int m, int k) { return B1[m * pA + k]; },
                     [&](int k, int c) { return B3[k * pB + c]; });
  if (warp == kW1) {  // Omega = V Bm V^T -> Z
    warp_mm(A, V, false, Bm, false);
    warp_mm(Z, A, false, V, true);
  }
  {
    const int npw = w_pitch_of(n);
    double* Wg = Wout + static_cast<int64_t>(g) * np * npw;
    gemm_store<MAXI>(acc, mlow, pg.M8, pg.M8, true, [&](int m, int c, double v0, double v1) {
      if (m < np && c < np) *reinterpret_cast<double2*>(Wg + m * npw + c) = make_double2(v0, v1);
    });
  }
  __syncthreads();
  EVC_PHASE(1, 7);
  for (int k = tid; k < n2; k += NT) {
    const int i = k / n, j = k - i * n;
    OmSout[o2 + k] = Z[i * ld + j] + Z[j * ld + i];
    PaoOut[o2 + k] = PaoS[i * ld + j];
  }
}

// ---------------------------------------------------------------------------
// K8b: streaming contraction of the derivative integrals, one CTA per (atom, geometry)
//   grad[g][A][x] = - sum_{mu in A, nu} <d_x mu|nu> OmS[mu,nu] + sum_{mu nu} dh[A,x][mu,nu] Pao[mu,nu]
//                   - 1/2 sum_{m in A} sum_{b, c >= d} (2 - d_cd) (d_x m b|c d) W[(mb),(cd)] + grad_nuc[g][A][x]
// on the PACKED derivative integrals eri_ip1p[x][m][b][CD] (int2e_ip1 is symmetric in its last two
// indices, evcont/ab_initio_gradients_loewdin.py:284): for a fixed (x, m) the n * np integrals are one
// contiguous run that is dotted with the n rows (m, b) of W -- no index arithmetic in the loop.
// Everything is read exactly once; the rows are staged with asynchronous 16-byte copies.
// ---------------------------------------------------------------------------
constexpr int kStreamThreads = 32;

// One single-warp CTA per (atom, geometry); no shared-memory staging: every thread issues its loads (coalesced
// 8-byte loads of the three integral runs, the matching entries of W, its share of the core-Hamiltonian
// derivative) ahead of their use, 32 CTAs are resident per SM.  Smaller CTAs were faster every time: 256 threads
// 0.265 ms, 128 threads 0.181 ms, 64 threads 0.161 ms, 32 threads 0.161 ms for 4096 H10-size geometries.  (The staged forms -- cp.async rows + two block barriers per AO, or one CTA per geometry
// with W expanded in shared memory -- spent a third of their time in barriers: profiles/r02r.)
__global__ void __launch_bounds__(kStreamThreads, 32)
grad_stream_kernel(int n, int natm, const int32_t* __restrict__ aoslices, const double* __restrict__ Wg,
                   const double* __restrict__ OmS, const double* __restrict__ Pao,
                   const double* __restrict__ ipovlp, const double* __restrict__ hcore_deriv,
                   const double* __restrict__ ip1p, const double* __restrict__ grad_nuc,
                   double* __restrict__ grad) {
  __shared__ double red[3][kStreamThreads / 32];
  const int np = npair_of(n), n2 = n * n, rl = n * np;  // rl: doubles of one (x, m) run
  const int At = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  const int p0 = __ldg(aoslices + 2 * At), p1 = __ldg(aoslices + 2 * At + 1);
  const int npw = w_pitch_of(n);
  const double* W = Wg + static_cast<int64_t>(g) * np * npw;
  const double* ipg = ip1p + static_cast<int64_t>(g) * 3 * n * rl;
  const int64_t xs = static_cast<int64_t>(n) * rl;  // stride between the x, y, z components
  const double* hd = hcore_deriv + (static_cast<int64_t>(g) * natm + At) * 3 * n2;
  const double* pa = Pao + static_cast<int64_t>(g) * n2;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  // (b, C) of element e = tid + k * kStreamThreads of a run, walked incrementally: e = b * np + C
  const int b_first = tid / np, c_first = tid - b_first * np;
  const int b_step = kStreamThreads / np, c_step = kStreamThreads - b_step * np;
  for (int m = p0; m < p1; ++m) {
    const double* r0 = ipg + static_cast<int64_t>(m) * rl;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    int b = b_first, C = c_first;
#pragma unroll 4
    for (int e = tid; e < rl; e += kStreamThreads) {
      const int r = m >= b ? tri_idx(m, b) : tri_idx(b, m);
      // weight 2 - delta_cd of the packed pair: C is a diagonal pair iff 8 C + 9 is an odd perfect square
      // ((2 c + 3)^2 = 8 c (c + 3) / 2 + 9); integer check on a float estimate
      const int q = static_cast<int>(sqrtf(static_cast<float>(8 * C + 9)) + 0.5f);
      const double fac = (q * q == 8 * C + 9) ? 1.0 : 2.0;
      const double w = __ldg(W + (r > C ? r * npw + C : C * npw + r)) * fac;
      s0 = fma(__ldg(r0 + e), w, s0);
      s1 = fma(__ldg(r0 + xs + e), w, s1);
      s2 = fma(__ldg(r0 + 2 * xs + e), w, s2);
      b += b_step;
      C += c_step;
      if (C >= np) { C -= np; ++b; }
    }
    a0 -= 0.5 * s0;
    a1 -= 0.5 * s1;
    a2 -= 0.5 * s2;
  }
  {  // overlap term: - sum_{mu in A, nu} <d_x mu|nu> OmS[mu,nu]
    const double* ipo = ipovlp + static_cast<int64_t>(g) * 3 * n2;
    const double* om = OmS + static_cast<int64_t>(g) * n2;
    for (int k = p0 * n + tid; k < p1 * n; k += kStreamThreads) {
      const double o = __ldg(om + k);
      a0 = fma(-__ldg(ipo + k), o, a0);
      a1 = fma(-__ldg(ipo + n2 + k), o, a1);
      a2 = fma(-__ldg(ipo + 2 * n2 + k), o, a2);
    }
  }
  // core-Hamiltonian derivative term
#pragma unroll 2
  for (int k = tid; k < n2; k += kStreamThreads) {
    const double p = __ldg(pa + k);
    a0 = fma(__ldg(hd + k), p, a0);
    a1 = fma(__ldg(hd + n2 + k), p, a1);
    a2 = fma(__ldg(hd + 2 * n2 + k), p, a2);
  }
  // fixed-order block reduction
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; }
  __syncthreads();
  if (tid < 3) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kStreamThreads / 32; ++w) t += red[tid][w];
    const int64_t o = (static_cast<int64_t>(g) * natm + At) * 3 + tid;
    grad[o] = t + (grad_nuc ? grad_nuc[o] : 0.0);
  }
}

size_t grad_stream_smem_bytes(int) { return 0; }  // static shared memory only

// ---------------------------------------------------------------------------
// int2e / int2e_ip1 as full tensors -> the packed arrays (callers that hold libcint-style tensors)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ao_pack8_kernel(int n, const double* __restrict__ eri, const double* __restrict__ eri_ip1,
                double* __restrict__ erip, double* __restrict__ ip1p) {
  const int g = blockIdx.y, np = npair_of(n), n2 = n * n;
  const PGeom pg = pgeom(n);
  const int64_t n4 = static_cast<int64_t>(n2) * n2;
  const int64_t ne = static_cast<int64_t>(np) * np, ni = static_cast<int64_t>(3) * n2 * np;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < ne + ni;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (t < ne) {
      if (erip == nullptr) continue;
      const int AB = static_cast<int>(t / np), CD = static_cast<int>(t - static_cast<int64_t>(AB) * np);
      int a, b, c, d;
      tril_unrank_i(AB, a, b);
      tril_unrank_i(CD, c, d);
      erip[static_cast<int64_t>(g) * np * pg.pA + static_cast<int64_t>(AB) * pg.pA + CD] =
          eri[static_cast<int64_t>(g) * n4 + (static_cast<int64_t>(a) * n + b) * n2 + c * n + d];
    } else {
      if (ip1p == nullptr) continue;
      const int64_t u = t - ne;
      const int64_t xmb = u / np;
      const int CD = static_cast<int>(u - xmb * np);
      int c, d;
      tril_unrank_i(CD, c, d);
      ip1p[static_cast<int64_t>(g) * ni + u] = eri_ip1[static_cast<int64_t>(g) * 3 * n4 + xmb * n2 + c * n + d];
    }
  }
}

size_t grad_smem_bytes(int n) {
  const PGeom pg = pgeom(n);
  return (2 * pg.szA + pg.szB + static_cast<size_t>(10) * n * (n | 1) + 2 * n) * sizeof(double) +
         table_bytes(n);
}

int maxi_for(int n, int nthreads) {
  const PGeom pg = pgeom(n);
  const int nitems = pg.M8 * ((pg.M8 + kNJ - 1) / kNJ);
  const int warps = nthreads / 32;
  return (nitems + warps - 1) / warps;
}

// deal the GEMM items to warps so that the four SM sub-partitions (warp % 4) carry equal tile counts
ItemMap build_item_map(int n, int nthreads, bool lower) {
  const PGeom pg = pgeom(n);
  const int M8 = pg.M8, NG = (M8 + kNJ - 1) / kNJ, nitems = M8 * NG, warps = nthreads / 32;
  const int slots = (nitems + warps - 1) / warps;
  ItemMap m;
  for (int w = 0; w < kMaxWarps; ++w)
    for (int r = 0; r < kMaxSlots; ++r) m.it[w][r] = -1;
  int size[256], order[256];
  for (int it = 0; it < nitems; ++it) {
    const int mt = it / NG, nt0 = (it % NG) * kNJ;
    int ntend = nt0 + kNJ < M8 ? nt0 + kNJ : M8;
    if (lower && ntend > mt + 1) ntend = mt + 1;
    size[it] = ntend > nt0 ? ntend - nt0 : 0;
    order[it] = it;
  }
  for (int i = 0; i < nitems; ++i)  // selection sort, largest first (stable)
    for (int j = i + 1; j < nitems; ++j)
      if (size[order[j]] > size[order[i]]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  int wload[kMaxWarps] = {0}, wcount[kMaxWarps] = {0}, sload[4] = {0};
  for (int q = 0; q < nitems; ++q) {
    const int it = order[q];
    int best = -1;
    for (int w = 0; w < warps; ++w) {
      if (wcount[w] >= slots) continue;
      if (best < 0) { best = w; continue; }
      const int sb = sload[best & 3], sw = sload[w & 3];
      if (sw < sb || (sw == sb && wload[w] < wload[best])) best = w;
    }
    m.it[best][wcount[best]++] = static_cast<signed char>(it);
    wload[best] += size[it];
    sload[best & 3] += size[it];
  }
  int w1 = warps - 1, w2 = warps > 1 ? warps - 2 : 0;
  for (int w = warps - 1; w >= 0; --w)
    if (wload[w] < wload[w1]) w1 = w;
  w2 = (w1 == warps - 1) ? (warps > 1 ? warps - 2 : 0) : warps - 1;
  for (int w = warps - 1; w >= 0; --w)
    if (w != w1 && wload[w] < wload[w2]) w2 = w;
  m.w1 = static_cast<unsigned char>(w1);
  m.w2 = static_cast<unsigned char>(w2);
  return m;
}

__global__ void add_enuc_kernel_p(int G, const double* __restrict__ e0, const double* __restrict__ e_nuc,
                                  double* __restrict__ E) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) E[g] = e0[g] + (e_nuc ? e_nuc[g] : 0.0);
}

}  // namespace

// ---- internal entry points (common.cuh) -------------------------------------
int evc_packed_ao2oao_percta(evc_ctx* ctx, int nbatch, int n, const double* x, const double* hcore,
                             const double* erip, double* hvec, double* Tout);

int evc_packed_ao2oao(evc_ctx* ctx, int nbatch, int n, const double* x, const double* hcore,
                      const double* erip, double* hvec, double* Tout) {
  EVC_REQUIRE(n >= 1 && n <= kPackedMaxNorb, "packed_ao2oao: n=%d unsupported", n);
  if (evc_packed_pipe_supported(n)) return evc_packed_ao2oao_pipe(ctx, nbatch, n, x, hcore, erip, hvec, Tout);
  return evc_packed_ao2oao_percta(ctx, nbatch, n, x, hcore, erip, hvec, Tout);
}

int evc_packed_ao2oao_percta(evc_ctx* ctx, int nbatch, int n, const double* x, const double* hcore,
                             const double* erip, double* hvec, double* Tout) {
  const size_t smem = ao2oao_smem_bytes(n);
  EVC_REQUIRE(smem <= ctx->smem_optin, "packed_ao2oao: needs %zu bytes of shared memory", smem);
  const int64_t L8 = packed_len(n);
  const int mi = maxi_for(n, kThreads);
  const ItemMap mf = build_item_map(n, kThreads, false), ml = build_item_map(n, kThreads, true);
#define EVC_CASE(MI, NCV)                                                                            \
  {                                                                                                  \
    auto kern = packed_ao2oao_kernel<MI, NCV>;                                                       \
    EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                        static_cast<int>(smem)));                                    \
    kern<<<nbatch, kThreads, smem, ctx->stream>>>(mf, ml, n, L8, x, hcore, erip, hvec, Tout);                 \
  }
  // compile-time orbital counts for the benchmark systems (H6, H10, H2O/6-31G), generic otherwise
  if (n == 6) EVC_CASE(1, 6) else if (n == 10) EVC_CASE(2, 10) else if (n == 13) EVC_CASE(5, 13)
  else if (mi <= 1) EVC_CASE(1, 0) else if (mi <= 2) EVC_CASE(2, 0) else if (mi <= 3) EVC_CASE(3, 0)
  else if (mi <= 4) EVC_CASE(4, 0) else EVC_CASE(5, 0)
#undef EVC_CASE
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_packed_grad(evc_ctx* ctx, int nbatch, int n, int natm, const int32_t* aoslices, const double* x,
                    const double* evals, const double* evecs, const double* hcore, const double* Tin,
                    const double* out7, const double* ipovlp, const double* hcore_deriv,
                    const double* eri_ip1p, const double* grad_nuc, double* Wg, double* OmS, double* Pao,
                    double* grad) {
  EVC_REQUIRE(n >= 1 && n <= kPackedMaxNorb, "packed_grad: n=%d unsupported", n);
  if (evc_packed_pipe_supported(n)) {
    int rcp = evc_packed_grad_pipe(ctx, nbatch, n, x, evals, evecs, hcore, Tin, out7, Wg, OmS, Pao);
    if (rcp) return rcp;
  } else {
  const size_t smem = grad_smem_bytes(n);
  EVC_REQUIRE(smem <= ctx->smem_optin, "packed_grad: needs %zu bytes of shared memory", smem);
  const int64_t L8 = packed_len(n);
  const int mi = maxi_for(n, kGradThreads);
  const ItemMap mf = build_item_map(n, kGradThreads, false), ml = build_item_map(n, kGradThreads, true);
#define EVC_CASE(MI, NCV)                                                                            \
  {                                                                                                  \
    auto kern = packed_grad_kernel<MI, NCV>;                                                         \
    EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                        static_cast<int>(smem)));                                    \
    kern<<<nbatch, kGradThreads, smem, ctx->stream>>>(mf, ml, n, L8, x, evals, evecs, hcore, Tin, out7, Wg, OmS, Pao); \
  }
  if (n == 6) EVC_CASE(1, 6) else if (n == 10) EVC_CASE(1, 10) else if (n == 13) EVC_CASE(3, 13)
  else if (mi <= 1) EVC_CASE(1, 0) else if (mi <= 2) EVC_CASE(2, 0) else EVC_CASE(3, 0)
#undef EVC_CASE
  EVC_CHECK_LAUNCH();
  }
  {
    int rcm = evc_stage_mark(ctx, EVC_STAGE_GRAD_STREAM);
    if (rcm) return rcm;
    const size_t sm2 = grad_stream_smem_bytes(n);
    dim3 grid(natm, nbatch);
    grad_stream_kernel<<<grid, kStreamThreads, sm2, ctx->stream>>>(n, natm, aoslices, Wg, OmS, Pao, ipovlp,
                                                                    hcore_deriv, eri_ip1p, grad_nuc, grad);
    EVC_CHECK_LAUNCH();
  }
  return 0;
}

int evc_packed_hvec_from_full(evc_ctx* ctx, int nbatch, int n, const double* h1, const double* h2, double* hvec) {
  const int64_t L8 = packed_len(n);
  dim3 grid(static_cast<unsigned>((L8 + 255) / 256), nbatch);
  hvec_from_full_kernel<<<grid, 256, 0, ctx->stream>>>(n, L8, h1, h2, hvec);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_packed_unpack_rdms(evc_ctx* ctx, int nbatch, int n, const double* out7, double* gamma, double* Gamma8) {
  const int64_t L8 = packed_len(n);
  dim3 grid(n * n, nbatch);
  unpack_rdms_kernel<<<grid, 256, 0, ctx->stream>>>(n, L8, out7, gamma, Gamma8);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_packed_pair_weights(evc_ctx* ctx, int nbatch, int N, const double* C, int64_t c_stride, double* w) {
  const int P = N * (N + 1) / 2;
  dim3 grid((P + 127) / 128, nbatch);
  tril_weights_kernel<<<grid, 128, 0, ctx->stream>>>(N, P, C, c_stride, w);
  EVC_CHECK_LAUNCH();
  return 0;
}

// ---- C ABI ---------------------------------------------------------------------
extern "C" {

// development aid: phase clocks of the per-geometry kernels (zeros unless built with
// -DEVC_PHASE_TIMING); out: [4][24] int64
int evc_debug_phase_clocks(long long* out_host) {
#ifdef EVC_PHASE_TIMING
  EVC_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_evc_phase, sizeof(long long) * 4 * 24));
#else
  for (int i = 0; i < 4 * 24; ++i) out_host[i] = 0;
#endif
  return 0;
}

int64_t evc_packed_row_len(int n) { return n >= 1 ? packed_len(n) : -1; }

// development aid: K4p alone (pipelined kernel if use_pipe != 0 and n <= 10, else the one-CTA-per-geometry kernel)
int evc_debug_packed_ao2oao(evc_ctx* ctx, int nbatch, int n, const double* x, const double* hcore, const double* erip,
                            double* hvec, double* Tout, int use_pipe) {
  if (use_pipe && n >= 2 && n <= kPackedPipeMaxNorb) return evc_packed_ao2oao_pipe(ctx, nbatch, n, x, hcore, erip, hvec, Tout);
  return evc_packed_ao2oao_percta(ctx, nbatch, n, x, hcore, erip, hvec, Tout);
}

int evc_erip_pitch(int n) { return n >= 1 ? pgeom(n).pA : -1; }
int64_t evc_erip_len(int n) { return n >= 1 ? erip_len(n) : -1; }
int64_t evc_eri_ip1p_len(int n) { return n >= 1 ? ip1p_len(n) : -1; }

int evc_ao_pack8(evc_ctx* ctx, int nbatch, int n, const double* eri, const double* eri_ip1, double* erip,
                 double* eri_ip1p) {
  EVC_REQUIRE(ctx != nullptr && n >= 1 && n <= 64, "evc_ao_pack8: bad arguments (n=%d)", n);
  EVC_REQUIRE((erip == nullptr) == (eri == nullptr) && (eri_ip1p == nullptr) == (eri_ip1 == nullptr),
              "evc_ao_pack8: each output needs its input tensor");
  if (nbatch <= 0 || (erip == nullptr && eri_ip1p == nullptr)) return 0;
  const int64_t work = static_cast<int64_t>(npair_of(n)) * npair_of(n) + static_cast<int64_t>(3) * n * n * npair_of(n);
  dim3 grid(static_cast<unsigned>(std::min<int64_t>((work + 255) / 256, 64)), nbatch);
  ao_pack8_kernel<<<grid, 256, 0, ctx->stream>>>(n, eri, eri_ip1, erip, eri_ip1p);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_stack_pack8(evc_ctx* ctx, int layout, int N, int n, const double* one_rdm, const double* two_rdm,
                    double* RH, double* RG) {
  EVC_REQUIRE(ctx && one_rdm && two_rdm && RH && RG, "evc_stack_pack8: NULL argument");
  EVC_REQUIRE(layout == EVC_LAYOUT_FULL || layout == EVC_LAYOUT_TRIL || layout == EVC_LAYOUT_FULL_EXCH ||
                  layout == EVC_LAYOUT_TRIL_EXCH,
              "evc_stack_pack8: two_RDM layout %d not one of 6/5/3/2", layout);
  EVC_REQUIRE(N >= 1 && n >= 1 && n <= 255, "evc_stack_pack8: N=%d n=%d unsupported", N, n);
  StackView sv;
  sv.two = two_rdm;
  sv.layout = layout;
  sv.N = N;
  sv.n = n;
  sv.n4 = static_cast<int64_t>(n) * n * n * n;
  sv.Lc = static_cast<int64_t>(n) * n * (static_cast<int64_t>(n) * n + 1) / 2;
  const int64_t L8 = packed_len(n);
  const int P = N * (N + 1) / 2;
  EVC_REQUIRE(P <= 65535, "evc_stack_pack8: too many state pairs (%d)", P);
  dim3 grid(static_cast<unsigned>((L8 + 255) / 256), P);
  pack8_stack_kernel<<<grid, 256, 0, ctx->stream>>>(sv, L8, one_rdm, RH, RG);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_energy_with_grad_packed_workspace_bytes(int N, int n, int natm, int nbatch, size_t* bytes) {
  EVC_REQUIRE(bytes != nullptr && N >= 1 && n >= 1 && nbatch >= 0, "evc_energy_with_grad_packed_workspace_bytes: bad arguments");
  const size_t G = static_cast<size_t>(nbatch);
  const size_t n2 = static_cast<size_t>(n) * n, n4 = n2 * n2;
  const int64_t L8 = packed_len(n);
  const int P = N * (N + 1) / 2;
  const size_t np = npair_of(n);
  size_t tot = 0;
  tot += 2 * evc_align_up(G * n2 * 8, 256);        // X, evecs
  tot += evc_align_up(G * n * 8, 256);             // evals
  tot += 2 * evc_align_up(G * L8 * 8, 256);        // hvec, out7
  tot += 2 * evc_align_up(G * P * 8, 256);         // Hp, w
  tot += evc_align_up(G * 8, 256);                 // E0
  tot += evc_align_up(G * N * 8, 256);             // C
  tot += evc_align_up(evc_rows_dot_ws_bytes(L8, P, nbatch), 256);
  tot += evc_align_up(evc_rows_axpy_ws_bytes(L8, P, nbatch), 256);
  if (n <= kPackedMaxNorb) {
    tot += evc_align_up(G * erip_len(n) * 8, 256);  // T ([np][pA], like erip)
    tot += evc_align_up(G * np * w_pitch_of(n) * 8, 256);  // W
    tot += 2 * evc_align_up(G * n2 * 8, 256);       // OmS, Pao
    // room for the packed copies of int2e / int2e_ip1 when the caller passes the full tensors
    tot += evc_align_up(G * erip_len(n) * 8, 256) + evc_align_up(G * ip1p_len(n) * 8, 256);
  } else {
    size_t gb = 0;
    int rc = evc_grad_workspace_bytes(n, natm, nbatch, &gb);
    if (rc) return rc;
    tot += 2 * evc_align_up(G * n2 * 8, 256);      // h1, gamma
    tot += 4 * evc_align_up(G * n4 * 8, 256);      // h2, t3, rot scratch, Gamma8
    tot += evc_align_up(gb, 256);
  }
  *bytes = tot;
  return 0;
}

int evc_energy_with_grad_packed(evc_ctx* ctx, int N, int n, int natm, const double* RH, const double* RG,
                                const double* Linv, int nbatch, const evc_ao_bundle* ao, double* E,
                                double* grad, double* Cvec, void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && RH && RG && Linv && ao && E && workspace, "evc_energy_with_grad_packed: NULL argument");
  // grad == NULL: energies (and Cvec) only -- the approximate_ground_state_OAO part of the step (K3, K4, K5, K6),
  // evcont/ab_initio_eigenvector_continuation.py:178-211; the derivative arrays of the bundle are not read
  const bool want_grad = grad != nullptr;
  EVC_REQUIRE(n >= 1 && n <= 32 && N >= 1 && N <= 112, "evc_energy_with_grad_packed: n=%d N=%d unsupported", n, N);
  const bool small = n <= kPackedMaxNorb;
  // two-electron arrays: packed (erip / eri_ip1p) or full tensors (eri / eri_ip1); n > 13 needs the tensors
  EVC_REQUIRE(ao->ovlp && ao->hcore && (ao->eri || (small && ao->erip)),
              "evc_energy_with_grad_packed: incomplete AO bundle");
  EVC_REQUIRE(!want_grad || (ao->ipovlp && ao->hcore_deriv && ao->aoslices && (ao->eri_ip1 || (small && ao->eri_ip1p))),
              "evc_energy_with_grad_packed: incomplete AO bundle (derivative arrays)");
  if (nbatch <= 0) return 0;
  const size_t G = static_cast<size_t>(nbatch);
  const size_t n2 = static_cast<size_t>(n) * n, n4 = n2 * n2;
  const int64_t L8 = packed_len(n);
  const int P = N * (N + 1) / 2;
  const size_t np = npair_of(n);
  const size_t dot_b = evc_rows_dot_ws_bytes(L8, P, nbatch), axpy_b = evc_rows_axpy_ws_bytes(L8, P, nbatch);
  evc_arena ar(workspace, workspace_bytes);
  double* X = ar.take<double>(G * n2);
  double* evecs = ar.take<double>(G * n2);
  double* evals = ar.take<double>(G * n);
  double* hvec = ar.take<double>(G * L8);
  double* out7 = ar.take<double>(G * L8);
  double* Hp = ar.take<double>(G * P);
  double* w = ar.take<double>(G * P);
  double* E0 = ar.take<double>(G);
  double* C = ar.take<double>(G * N);
  char* dot_ws = ar.take<char>(dot_b);
  char* axpy_ws = ar.take<char>(axpy_b);
  EVC_REQUIRE(X && evecs && evals && hvec && out7 && Hp && w && E0 && C && dot_ws && axpy_ws,
              "evc_energy_with_grad_packed: workspace too small (%zu bytes)", workspace_bytes);
  double *T = nullptr, *Wg = nullptr, *OmS = nullptr, *Pao = nullptr, *h1 = nullptr, *gamma = nullptr, *h2 = nullptr, *t3 = nullptr, *scratch = nullptr,
         *Gamma8 = nullptr;
  char* grad_ws = nullptr;
  size_t grad_b = 0;
  int rc;
  const double* erip = ao->erip;
  const double* ip1p = ao->eri_ip1p;
  if (small) {
    T = ar.take<double>(G * erip_len(n));
    Wg = ar.take<double>(G * np * w_pitch_of(n));
    OmS = ar.take<double>(G * n2);
    Pao = ar.take<double>(G * n2);
    EVC_REQUIRE(T && Wg && OmS && Pao, "evc_energy_with_grad_packed: workspace too small (%zu bytes)", workspace_bytes);
    // callers holding the full libcint-style tensors: pack them first (one extra streaming pass)
    const bool need_e = erip == nullptr, need_i = want_grad && ip1p == nullptr;
    if (need_e || need_i) {
      double* ep = need_e ? ar.take<double>(G * erip_len(n)) : nullptr;
      double* ipp = need_i ? ar.take<double>(G * ip1p_len(n)) : nullptr;
      EVC_REQUIRE((!need_e || ep) && (!need_i || ipp), "evc_energy_with_grad_packed: workspace too small (%zu bytes)",
                  workspace_bytes);
      if ((rc = evc_ao_pack8(ctx, nbatch, n, need_e ? ao->eri : nullptr, need_i ? ao->eri_ip1 : nullptr, ep, ipp)))
        return rc;
      if (need_e) erip = ep;
      if (need_i) ip1p = ipp;
    }
  } else {
    if ((rc = evc_grad_workspace_bytes(n, natm, nbatch, &grad_b))) return rc;
    h1 = ar.take<double>(G * n2);
    gamma = ar.take<double>(G * n2);
    h2 = ar.take<double>(G * n4);
    t3 = ar.take<double>(G * n4);
    scratch = ar.take<double>(G * n4);
    Gamma8 = ar.take<double>(G * n4);
    grad_ws = ar.take<char>(grad_b);
    EVC_REQUIRE(h1 && gamma && h2 && t3 && scratch && Gamma8 && grad_ws,
                "evc_energy_with_grad_packed: workspace too small (%zu bytes)", workspace_bytes);
  }
  if (Cvec) C = Cvec;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_LOEWDIN))) return rc;
  if ((rc = evc_loewdin(ctx, nbatch, n, ao->ovlp, X, evals, evecs))) return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_AO2OAO))) return rc;
  if (small) {
    if ((rc = evc_packed_ao2oao(ctx, nbatch, n, X, ao->hcore, erip, hvec, T))) return rc;
  } else {
    if ((rc = evc_ao2oao(ctx, nbatch, n, ao->hcore, ao->eri, X, 0, h1, h2, t3, scratch, evc_align_up(G * n4 * 8, 256)))) return rc;
    if ((rc = evc_packed_hvec_from_full(ctx, nbatch, n, h1, h2, hvec))) return rc;
  }
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_SUBSPACE_H))) return rc;
  if ((rc = evc_rows_dot(ctx, RH, L8, P, hvec, nbatch, Hp, dot_ws, dot_b))) return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_GENEIG))) return rc;
  if ((rc = evc_launch_geneig(ctx, nbatch, N, 1, Hp, Linv, 1, E0, C))) return rc;
  if (want_grad) {
    if ((rc = evc_stage_mark(ctx, EVC_STAGE_PREDICT))) return rc;
    if ((rc = evc_packed_pair_weights(ctx, nbatch, N, C, N, w))) return rc;
    if ((rc = evc_rows_axpy(ctx, RG, L8, P, w, nbatch, out7, axpy_ws, axpy_b))) return rc;
    if ((rc = evc_stage_mark(ctx, EVC_STAGE_GRAD))) return rc;
    if (small) {
      if ((rc = evc_packed_grad(ctx, nbatch, n, natm, ao->aoslices, X, evals, evecs, ao->hcore, T, out7, ao->ipovlp,
                                ao->hcore_deriv, ip1p, ao->grad_nuc, Wg, OmS, Pao, grad)))
        return rc;
    } else {
      if ((rc = evc_packed_unpack_rdms(ctx, nbatch, n, out7, gamma, Gamma8))) return rc;
      if ((rc = evc_grad_elec_full(ctx, nbatch, n, natm, ao->aoslices, evals, evecs, X, ao->hcore, t3, gamma, Gamma8,
                                   ao->ipovlp, ao->hcore_deriv, ao->eri_ip1, ao->grad_nuc, grad, grad_ws, grad_b, 1)))
        return rc;
    }
    if (!small && (rc = evc_stage_mark(ctx, EVC_STAGE_GRAD_STREAM))) return rc;
  }
  add_enuc_kernel_p<<<(nbatch + 127) / 128, 128, 0, ctx->stream>>>(nbatch, E0, ao->e_nuc, E);
  EVC_CHECK_LAUNCH();
  return evc_stage_mark(ctx, EVC_NSTAGE);
}

}  // extern "C"
