// K3 (Loewdin + its derivative), K4 (AO->OAO transform), K6 (generalized
// eigenproblem): small dense linear algebra, batched over geometries.
//
// Reference anchors: evcont/electron_integral_utils.py:6-18 (get_loewdin_trafo),
// :122-138 (get_integrals); evcont/ab_initio_gradients_loewdin.py:41-134
// (loewdin_trafo_grad, get_derivative_ao_mo_trafo); evcont/
// ab_initio_eigenvector_continuation.py:75-88,157-173 (scipy eigh(H,S) + root pick).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------
// Two-sided Jacobi eigensolver for a symmetric m x m matrix held in shared
// memory, parallel (round-robin tournament) ordering.  Every round applies
// mp/2 disjoint rotations J = prod R(p_k, q_k).  A' = J^T A J decomposes into
// independent 2x2 blocks A'[{p1,q1},{p2,q2}] = R1^T A[{p1,q1},{p2,q2}] R2, so a
// round is two phases only: (1) the mp/2 rotation angles, (2) one thread per 2x2
// block applies the row and the column rotation at once (and V <- V J), i.e. two
// block barriers per round and no integer division / modulo in the inner loops
// (the tournament schedule is tabulated once).
//
// Storage: A is [mp][lda] with mp = m + (m & 1) (for odd m a zero row/column m is
// the bye of the tournament; its rotation is the identity), V is [m][ldv] with
// mp columns used.  On exit diag(A) holds the eigenvalues, the columns of V the
// eigenvectors.  Scratch: cs 4*(mp/2) doubles, red 18 doubles,
// sched (mp-1)*(mp/2) uint16 (p | q << 8).
// ---------------------------------------------------------------------------
__host__ __device__ inline size_t jacobi_scratch_bytes(int m) {
  const int mp = m + (m & 1), half = mp / 2;
  size_t b = (4 * static_cast<size_t>(half) + 18) * sizeof(double);
  b += (static_cast<size_t>(mp > 1 ? mp - 1 : 1) * half * sizeof(unsigned short) + 7) / 8 * 8;
  return b;
}

__device__ void jacobi_eigh_smem(double* A, int lda, double* V, int ldv, int m, double* cs,
                                 double* red, unsigned short* sched) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int mp = m + (m & 1);
  const int half = mp / 2;
  for (int k = tid; k < m * mp; k += nt) {
    const int i = k / mp, j = k - i * mp;
    V[i * ldv + j] = (i == j) ? 1.0 : 0.0;
  }
  if (m & 1) {  // bye row / column
    for (int k = tid; k < mp; k += nt) { A[m * lda + k] = 0.0; A[k * lda + m] = 0.0; }
  }
  for (int k = tid; k < (mp - 1) * half; k += nt) {
    const int r = k / half, i = k - r * half;
    int p = (i == 0) ? 0 : 1 + (i - 1 + r) % (mp - 1);
    const int i2 = mp - 1 - i;
    int q = 1 + (i2 - 1 + r) % (mp - 1);
    if (p > q) { const int t = p; p = q; q = t; }
    sched[k] = static_cast<unsigned short>(p | (q << 8));
  }
  __syncthreads();
  if (m == 1) return;
  const double tol = (static_cast<double>(m) * 2.3e-16) * (static_cast<double>(m) * 2.3e-16);
  for (int sweep = 0; sweep < 30; ++sweep) {
    // convergence: squared off-diagonal norm vs squared Frobenius norm (fixed
    // summation order: per-warp partials, then warp 0 .. nwarps-1)
    {
      double off = 0.0, tot = 0.0;
      for (int k = tid; k < m * m; k += nt) {
        const int i = k / m, j = k - i * m;
        const double v = A[i * lda + j];
        tot += v * v;
        if (i != j) off += v * v;
      }
      for (int o = 16; o > 0; o >>= 1) {
        off += __shfl_xor_sync(0xffffffffu, off, o);
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
      }
      if ((tid & 31) == 0) { red[2 + 2 * (tid >> 5)] = off; red[3 + 2 * (tid >> 5)] = tot; }
    }
    __syncthreads();
    double off = 0.0, tot = 0.0;
    for (int w = 0; w < (nt + 31) / 32; ++w) { off += red[2 + 2 * w]; tot += red[3 + 2 * w]; }
    __syncthreads();
    if (off <= tol * tot) break;
    for (int r = 0; r < mp - 1; ++r) {
      const unsigned short* sr = sched + r * half;
      // 1. rotation angles
      for (int k = tid; k < half; k += nt) {
        const int p = sr[k] & 0xff, q = sr[k] >> 8;
        const double app = A[p * lda + p], aqq = A[q * lda + q], apq = A[p * lda + q];
        double c = 1.0, s = 0.0, npp = app, nqq = aqq;
        if (fabs(apq) > 1.0e-300) {
          const double theta = (aqq - app) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          c = rsqrt(t * t + 1.0);
          s = t * c;
          npp = app - t * apq;
          nqq = aqq + t * apq;
        }
        cs[4 * k] = c; cs[4 * k + 1] = s; cs[4 * k + 2] = npp; cs[4 * k + 3] = nqq;
      }
      __syncthreads();
      // 2. 2x2 blocks of A (rows then columns) and V <- V J
      const int nblk = half * half;
      for (int idx = tid; idx < nblk + half * m; idx += nt) {
        if (idx < nblk) {
          const int k1 = idx / half, k2 = idx - k1 * half;
          const int p1 = sr[k1] & 0xff, q1 = sr[k1] >> 8;
          const int p2 = sr[k2] & 0xff, q2 = sr[k2] >> 8;
          double* a_pp = A + p1 * lda + p2;
          double* a_pq = A + p1 * lda + q2;
          double* a_qp = A + q1 * lda + p2;
          double* a_qq = A + q1 * lda + q2;
          if (k1 == k2) {
            *a_pp = cs[4 * k1 + 2]; *a_pq = 0.0; *a_qp = 0.0; *a_qq = cs[4 * k1 + 3];
          } else {
            const double c1 = cs[4 * k1], s1 = cs[4 * k1 + 1], c2 = cs[4 * k2], s2 = cs[4 * k2 + 1];
            const double a = *a_pp, b = *a_pq, c = *a_qp, d = *a_qq;
            const double ra = c1 * a - s1 * c, rc = s1 * a + c1 * c;  // rows
            const double rb = c1 * b - s1 * d, rd = s1 * b + c1 * d;
            *a_pp = c2 * ra - s2 * rb; *a_pq = s2 * ra + c2 * rb;     // columns
            *a_qp = c2 * rc - s2 * rd; *a_qq = s2 * rc + c2 * rd;
          }
        } else {
          const int e = idx - nblk;
          const int k = e / m, i = e - k * m;
          const int p = sr[k] & 0xff, q = sr[k] >> 8;
          const double c = cs[4 * k], s = cs[4 * k + 1];
          const double vx = V[i * ldv + p], vy = V[i * ldv + q];
          V[i * ldv + p] = c * vx - s * vy;
          V[i * ldv + q] = s * vx + c * vy;
        }
      }
      __syncthreads();
    }
  }
}

// rank[i] = position of eigenvalue i in ascending order (ties by index)
__device__ __forceinline__ int ascending_rank(const double* A, int lda, int m, int i) {
  const double wi = A[i * lda + i];
  int rk = 0;
  for (int j = 0; j < m; ++j) {
    const double wj = A[j * lda + j];
    rk += (wj < wi) || (wj == wi && j < i);
  }
  return rk;
}

// ---------------------------------------------------------------------------
// K3: Loewdin
// ---------------------------------------------------------------------------
__global__ void loewdin_kernel(int n, const double* __restrict__ s_ao, double* __restrict__ x,
                               double* __restrict__ evals, double* __restrict__ evecs) {
  extern __shared__ __align__(16) double sm[];
  const int np = n + (n & 1);
  const int lda = np + 1;
  double* A = sm;                // [np][lda]
  double* V = A + np * lda;      // [n][lda]
  double* w = V + n * lda;       // sorted eigenvalues
  double* f = w + n;             // s^-1/2 (0 if s <= 1e-15)
  double* cs = f + n;
  double* red = cs + 4 * (np / 2);
  unsigned short* sched = reinterpret_cast<unsigned short*>(red + 18);
  int* perm = reinterpret_cast<int*>(reinterpret_cast<char*>(cs) + jacobi_scratch_bytes(n));  // perm[rank] = column
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double* S = s_ao + static_cast<int64_t>(b) * n * n;
  // symmetrise on load: numpy.linalg.eigh reads the lower triangle only
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    A[i * lda + j] = (i >= j) ? S[i * n + j] : S[j * n + i];
  }
  __syncthreads();
  jacobi_eigh_smem(A, lda, V, lda, n, cs, red, sched);
  for (int i = tid; i < n; i += nt) {
    const int rk = ascending_rank(A, lda, n, i);
    const double wi = A[i * lda + i];
    perm[rk] = i;
    w[rk] = wi;
    f[rk] = wi > 1.0e-15 ? 1.0 / sqrt(wi) : 0.0;
  }
  __syncthreads();
  const int64_t o2 = static_cast<int64_t>(b) * n * n;
  for (int k = tid; k < n; k += nt) evals[static_cast<int64_t>(b) * n + k] = w[k];
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    evecs[o2 + k] = V[i * lda + perm[j]];
    double acc = 0.0;
    for (int r = 0; r < n; ++r) {
      const int c = perm[r];
      acc += V[i * lda + c] * f[r] * V[j * lda + c];
    }
    x[o2 + k] = acc;
  }
}

// dX_xi = V (G o (V^T dS_xi V)) V^T
__global__ void loewdin_grad_kernel(int n, int nder, const double* __restrict__ evals,
                                    const double* __restrict__ evecs, const double* __restrict__ dS,
                                    double* __restrict__ dX) {
  extern __shared__ double sm[];
  const int ld = n + 1;
  double* V = sm;
  double* B = V + n * ld;
  double* C = B + n * ld;
  double* w = C + n * ld;
  const int b = blockIdx.y, xi = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int64_t ob = static_cast<int64_t>(b) * n * n;
  const int64_t ox = (static_cast<int64_t>(b) * nder + xi) * n * n;
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    V[i * ld + j] = evecs[ob + k];
    B[i * ld + j] = dS[ox + k];
  }
  for (int k = tid; k < n; k += nt) w[k] = evals[static_cast<int64_t>(b) * n + k];
  __syncthreads();
  // C = V^T B
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += V[r * ld + i] * B[r * ld + j];
    C[i * ld + j] = acc;
  }
  __syncthreads();
  // B = (C V) o G
  for (int k = tid; k < n * n; k += nt) {
    const int p = k / n, q = k - p * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += C[p * ld + r] * V[r * ld + q];
    const double sp = w[p], sq = w[q];
    const bool kp = sp > 1.0e-15, kq = sq > 1.0e-15;
    double gpq;
    if (kp && kq) {
      const double rp = sqrt(sp), rq = sqrt(sq);
      gpq = -1.0 / (rp * rq * (rp + rq));
    } else if (kp != kq && sp != sq) {
      gpq = ((kp ? 1.0 / sqrt(sp) : 0.0) - (kq ? 1.0 / sqrt(sq) : 0.0)) / (sp - sq);
    } else {
      gpq = 0.0;
    }
    B[p * ld + q] = acc * gpq;
  }
  __syncthreads();
  // C = V B
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += V[i * ld + r] * B[r * ld + j];
    C[i * ld + j] = acc;
  }
  __syncthreads();
  // dX = C V^T
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += C[i * ld + r] * V[j * ld + r];
    dX[ox + k] = acc;
  }
}

// ---------------------------------------------------------------------------
// K4: AO -> OAO.  h1 = C^T h C (one CTA per geometry) and one "rotating" pass
// of the four-index transform:
//    out[m, a] = sum_i in[i, m] M[i, a],   m = 0 .. n^3-1
// i.e. the leading index is contracted and re-appears as the trailing one, so
// four passes return the tensor to its original index order.  Each CTA owns
// 128 consecutive m: the [n x 128] slab of `in` is staged in shared memory with
// coalesced loads, multiplied on the FP64 tensor cores (DMMA m8n8k4) and the
// [128 x n] result (contiguous in `out`) is written back coalesced.
// ---------------------------------------------------------------------------
__global__ void h1_transform_kernel(int n, const double* __restrict__ hcore,
                                    const double* __restrict__ c, int transpose_c,
                                    double* __restrict__ h1) {
  extern __shared__ double sm[];
  const int ld = n + 1;
  double* Cs = sm;
  double* Hs = Cs + n * ld;
  double* Ts = Hs + n * ld;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int64_t o = static_cast<int64_t>(b) * n * n;
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    const double v = c[o + k];
    if (transpose_c) Cs[j * ld + i] = v; else Cs[i * ld + j] = v;
    Hs[i * ld + j] = hcore[o + k];
  }
  __syncthreads();
  for (int k = tid; k < n * n; k += nt) {  // T = H C
    const int i = k / n, j = k - i * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += Hs[i * ld + r] * Cs[r * ld + j];
    Ts[i * ld + j] = acc;
  }
  __syncthreads();
  for (int k = tid; k < n * n; k += nt) {  // h1 = C^T T
    const int i = k / n, j = k - i * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += Cs[r * ld + i] * Ts[r * ld + j];
    h1[o + k] = acc;
  }
}

constexpr int kRotMT = 128;       // m rows per CTA
constexpr int kRotPitchA = 132;   // == 4 (mod 16): conflict-free DMMA fragment loads

template <int NT8>
__global__ void __launch_bounds__(256)
rot_pass_kernel(int n, int64_t n3, const double* __restrict__ in, const double* __restrict__ M,
                int transpose_m, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const int K4 = (n + 3) & ~3;
  constexpr int pitchM = NT8 * 8 + 4;
  double* Ms = sm;                    // [K4][pitchM]
  double* As = Ms + K4 * pitchM;      // [K4][kRotPitchA]; later reused as Cs [128][n]
  const int b = blockIdx.y, tid = threadIdx.x;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kRotMT;
  const int64_t mleft = n3 - m0;
  const int mcount = mleft < kRotMT ? static_cast<int>(mleft) : kRotMT;
  const double* Mg = M + static_cast<int64_t>(b) * n * n;
  const double* ing = in + static_cast<int64_t>(b) * n3 * n;
  double* outg = out + static_cast<int64_t>(b) * n3 * n;
  for (int k = tid; k < K4 * pitchM; k += 256) {
    const int i = k / pitchM, a = k - i * pitchM;
    double v = 0.0;
    if (i < n && a < n) v = transpose_m ? Mg[a * n + i] : Mg[i * n + a];
    Ms[k] = v;
  }
  for (int k = tid; k < K4 * kRotMT; k += 256) {
    const int i = k / kRotMT, m = k - i * kRotMT;
    double v = 0.0;
    if (i < n && m < mcount) v = ing[static_cast<int64_t>(i) * n3 + m0 + m];
    As[i * kRotPitchA + m] = v;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  double acc[2][NT8][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int at = 0; at < NT8; ++at) acc[mt][at][0] = acc[mt][at][1] = 0.0;
  for (int k0 = 0; k0 < K4; k0 += 4) {
    double af[2], bf[NT8];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) af[mt] = As[(k0 + tg) * kRotPitchA + warp * 16 + mt * 8 + g];
#pragma unroll
    for (int at = 0; at < NT8; ++at) bf[at] = Ms[(k0 + tg) * pitchM + at * 8 + g];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int at = 0; at < NT8; ++at) dmma8x8x4(acc[mt][at][0], acc[mt][at][1], af[mt], bf[at]);
  }
  __syncthreads();
  double* Cs = As;  // [128][n]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int at = 0; at < NT8; ++at)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int m = warp * 16 + mt * 8 + g, a = at * 8 + tg * 2 + e;
        if (a < n) Cs[m * n + a] = acc[mt][at][e];
      }
  __syncthreads();
  double* dst = outg + m0 * n;
  const int tot = mcount * n;
  for (int k = tid; k < tot; k += 256) dst[k] = Cs[k];
}

int launch_rot_pass(cudaStream_t st, int nbatch, int n, const double* in, const double* M,
                    int transpose_m, double* out) {
  const int64_t n3 = static_cast<int64_t>(n) * n * n;
  const int K4 = (n + 3) & ~3;
  const int nt8 = (n + 7) / 8;
  const size_t smem = (static_cast<size_t>(K4) * (nt8 * 8 + 4) +
                       static_cast<size_t>(max(K4 * kRotPitchA, kRotMT * n))) * sizeof(double);
  dim3 grid(static_cast<unsigned>((n3 + kRotMT - 1) / kRotMT), nbatch);
#define EVC_ROT_CASE(NT8)                                                                      \
  case NT8: {                                                                                  \
    auto kern = rot_pass_kernel<NT8>;                                                          \
    EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                        static_cast<int>(smem)));                              \
    kern<<<grid, 256, smem, st>>>(n, n3, in, M, transpose_m, out);                             \
    break;                                                                                     \
  }
  switch (nt8) {
    EVC_ROT_CASE(1)
    EVC_ROT_CASE(2)
    EVC_ROT_CASE(3)
    EVC_ROT_CASE(4)
    default:
      evc_set_error("rot_pass: n=%d unsupported (n <= 32)", n);
      return -1;
  }
#undef EVC_ROT_CASE
  EVC_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------
// K6: generalized symmetric-definite eigenproblem
// ---------------------------------------------------------------------------
// S = L L^T, Linv = L^-1.  Single CTA (once per stack).
__global__ void cholesky_inverse_kernel(int N, const double* __restrict__ S, double* __restrict__ Linv,
                                        int* __restrict__ info) {
  extern __shared__ double sm[];
  const int ld = N + 1;
  double* L = sm;
  __shared__ int fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) fail = 0;
  // LAPACK dsygvd with the default lower=True reads the lower triangle of S
  for (int k = tid; k < N * N; k += nt) {
    const int i = k / N, j = k - i * N;
    L[i * ld + j] = (i >= j) ? S[i * N + j] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < N; ++j) {
    if (tid == 0) {
      const double d = L[j * ld + j];
      if (!(d > 0.0)) fail = j + 1; else L[j * ld + j] = sqrt(d);
    }
    __syncthreads();
    if (fail) break;
    const double piv = L[j * ld + j];
    for (int i = j + 1 + tid; i < N; i += nt) L[i * ld + j] /= piv;
    __syncthreads();
    // trailing update of the lower triangle
    const int rem = N - j - 1;
    for (int k = tid; k < rem * rem; k += nt) {
      const int i = j + 1 + k / rem, c = j + 1 + k % rem;
      if (c <= i) L[i * ld + c] -= L[i * ld + j] * L[c * ld + j];
    }
    __syncthreads();
  }
  if (tid == 0) *info = fail;
  if (fail) return;
  // column c of L^-1 by forward substitution (one thread per column)
  for (int c = tid; c < N; c += nt) {
    for (int i = 0; i < N; ++i) {
      double v = (i == c) ? 1.0 : 0.0;
      if (i < c) { Linv[i * N + c] = 0.0; continue; }
      for (int k = c; k < i; ++k) v -= L[i * ld + k] * Linv[k * N + c];
      Linv[i * N + c] = v / L[i * ld + i];
    }
  }
}

// packed_lower != 0: H holds the lower triangle only, [nbatch][N(N+1)/2] in
// np.tril_indices order (the form the packed K5 GEMM writes).
__global__ void geneig_kernel(int N, int packed_lower, const double* __restrict__ H,
                              const double* __restrict__ Linv, int nroots, double* __restrict__ E,
                              double* __restrict__ C) {
  extern __shared__ __align__(16) double sm[];
  const int Np = N + (N & 1);
  const int ld = Np + 1;
  double* A = sm;              // [Np][ld]
  double* V = A + Np * ld;     // [N][ld]
  double* cs = V + N * ld;
  double* red = cs + 4 * (Np / 2);
  unsigned short* sched = reinterpret_cast<unsigned short*>(red + 18);
  int* perm = reinterpret_cast<int*>(reinterpret_cast<char*>(cs) + jacobi_scratch_bytes(N));
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  // A <- lower triangle of H, mirrored (scipy eigh(H, S) reads the lower triangle)
  if (packed_lower) {
    const double* Hb = H + static_cast<int64_t>(b) * (N * (N + 1) / 2);
    for (int k = tid; k < N * N; k += nt) {
      const int i = k / N, j = k - i * N;
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      A[i * ld + j] = Hb[hi * (hi + 1) / 2 + lo];
    }
  } else {
    const double* Hb = H + static_cast<int64_t>(b) * N * N;
    for (int k = tid; k < N * N; k += nt) {
      const int i = k / N, j = k - i * N;
      A[i * ld + j] = (i >= j) ? Hb[i * N + j] : Hb[j * N + i];
    }
  }
  __syncthreads();
  // V <- Linv * A
  for (int k = tid; k < N * N; k += nt) {
    const int i = k / N, j = k - i * N;
    double acc = 0.0;
    for (int r = 0; r <= i; ++r) acc += __ldg(Linv + i * N + r) * A[r * ld + j];
    V[i * ld + j] = acc;
  }
  __syncthreads();
  // A <- V * Linv^T  (symmetric; computed for i >= j and mirrored)
  for (int k = tid; k < N * N; k += nt) {
    const int i = k / N, j = k - i * N;
    if (j > i) continue;
    double acc = 0.0;
    for (int r = 0; r <= j; ++r) acc += V[i * ld + r] * __ldg(Linv + j * N + r);
    A[i * ld + j] = acc;
  }
  __syncthreads();
  for (int k = tid; k < N * N; k += nt) {
    const int i = k / N, j = k - i * N;
    if (j > i) A[i * ld + j] = A[j * ld + i];
  }
  __syncthreads();
  jacobi_eigh_smem(A, ld, V, ld, N, cs, red, sched);
  for (int i = tid; i < N; i += nt) perm[ascending_rank(A, ld, N, i)] = i;
  __syncthreads();
  for (int r = tid; r < nroots; r += nt) E[static_cast<int64_t>(b) * nroots + r] = A[perm[r] * ld + perm[r]];
  // c = Linv^T y
  for (int k = tid; k < nroots * N; k += nt) {
    const int r = k / N, i = k - r * N;
    const int col = perm[r];
    double acc = 0.0;
    for (int q = i; q < N; ++q) acc += __ldg(Linv + q * N + i) * V[q * ld + col];
    C[(static_cast<int64_t>(b) * nroots + r) * N + i] = acc;
  }
}

size_t loewdin_smem_bytes(int n) {
  const int np = n + (n & 1);
  return (static_cast<size_t>(np + n) * (np + 1) + 2 * n) * sizeof(double) + jacobi_scratch_bytes(n) +
         (n + 2) * sizeof(int);
}

size_t geneig_smem_bytes(int N) {
  const int Np = N + (N & 1);
  return static_cast<size_t>(Np + N) * (Np + 1) * sizeof(double) + jacobi_scratch_bytes(N) +
         (N + 2) * sizeof(int);
}

int launch_geneig(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H, const double* Linv,
                  int nroots, double* E, double* C) {
  // ground state only: tridiagonalisation + bisection + inverse iteration, one warp per geometry
  if (nroots == 1) return evc_launch_geneig_lowest(ctx, nbatch, N, packed_lower, H, Linv, E, C);
  const size_t smem = geneig_smem_bytes(N);
  EVC_CHECK_CUDA(cudaFuncSetAttribute(geneig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  geneig_kernel<<<nbatch, N <= 32 ? 128 : 256, smem, ctx->stream>>>(N, packed_lower, H, Linv, nroots, E, C);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// exported to the other translation units
int evc_launch_rot_pass(cudaStream_t st, int nbatch, int n, const double* in, const double* M,
                        int transpose_m, double* out) {
  return launch_rot_pass(st, nbatch, n, in, M, transpose_m, out);
}

int evc_launch_geneig(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H,
                      const double* Linv, int nroots, double* E, double* C) {
  return launch_geneig(ctx, nbatch, N, packed_lower, H, Linv, nroots, E, C);
}

extern "C" {

int evc_loewdin(evc_ctx* ctx, int nbatch, int n, const double* s_ao, double* x, double* evals,
                double* evecs) {
  EVC_REQUIRE(ctx && s_ao && x && evals && evecs, "evc_loewdin: NULL argument");
  EVC_REQUIRE(n >= 1 && n <= 32, "evc_loewdin: n=%d unsupported (1..32)", n);
  if (nbatch <= 0) return 0;
  // batches: the register-resident form (loewdin_reg.cu, several matrices per warp)
  if (evc_loewdin_reg_supported(n) && nbatch >= evc_loewdin_reg_min_batch())
    return evc_loewdin_reg(ctx, nbatch, n, s_ao, x, evals, evecs);
  const size_t smem = loewdin_smem_bytes(n);
  // many small problems (>= 16 per SM): one warp per geometry (0.136 ms per 4096 H10-size geometries against
  // 0.177 ms with 128 threads, whose barriers and idle lanes dominate); fewer problems: wider teams for latency
  // (1024 geometries: 0.063 ms with 128 threads, 0.083 ms with 32)
  int threads = n <= 16 ? 128 : 256;
  if (n <= 12 && nbatch >= 16 * ctx->sm_count) threads = 32;
  loewdin_kernel<<<nbatch, threads, smem, ctx->stream>>>(n, s_ao, x, evals, evecs);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_loewdin_grad(evc_ctx* ctx, int nbatch, int n, int nder, const double* evals,
                     const double* evecs, const double* dS, double* dX) {
  EVC_REQUIRE(ctx && evals && evecs && dS && dX, "evc_loewdin_grad: NULL argument");
  EVC_REQUIRE(n >= 1 && n <= 32, "evc_loewdin_grad: n=%d unsupported (1..32)", n);
  if (nbatch <= 0 || nder <= 0) return 0;
  const size_t smem = (3 * n * (n + 1) + n) * sizeof(double);
  dim3 grid(nder, nbatch);
  loewdin_grad_kernel<<<grid, n <= 16 ? 128 : 256, smem, ctx->stream>>>(n, nder, evals, evecs, dS, dX);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_ao2oao(evc_ctx* ctx, int nbatch, int n, const double* hcore, const double* eri,
               const double* c, int transpose_c, double* h1, double* h2, double* t3,
               void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && c, "evc_ao2oao: NULL argument");
  EVC_REQUIRE(n >= 1 && n <= 32, "evc_ao2oao: n=%d unsupported (1..32)", n);
  if (nbatch <= 0) return 0;
  if (hcore != nullptr) {
    EVC_REQUIRE(h1 != nullptr, "evc_ao2oao: h1 is NULL");
    const size_t smem = 3 * n * (n + 1) * sizeof(double);
    h1_transform_kernel<<<nbatch, n <= 16 ? 128 : 256, smem, ctx->stream>>>(n, hcore, c, transpose_c, h1);
    EVC_CHECK_LAUNCH();
  }
  if (eri != nullptr) {
    EVC_REQUIRE(h2 != nullptr, "evc_ao2oao: h2 is NULL");
    const size_t n4 = static_cast<size_t>(n) * n * n * n;
    evc_arena ar(workspace, workspace_bytes);
    double* bufA = ar.take<double>(nbatch * n4);
    double* bufB = t3 ? t3 : ar.take<double>(nbatch * n4);
    EVC_REQUIRE(bufA && bufB, "evc_ao2oao: workspace too small (%zu bytes given, need %zu)",
                workspace_bytes, 2 * evc_align_up(nbatch * n4 * 8, 256));
    // eri -> h2 (as scratch) -> bufA -> bufB (= t3) -> h2
    int rc;
    if ((rc = launch_rot_pass(ctx->stream, nbatch, n, eri, c, transpose_c, h2))) return rc;
    if ((rc = launch_rot_pass(ctx->stream, nbatch, n, h2, c, transpose_c, bufA))) return rc;
    if ((rc = launch_rot_pass(ctx->stream, nbatch, n, bufA, c, transpose_c, bufB))) return rc;
    if ((rc = launch_rot_pass(ctx->stream, nbatch, n, bufB, c, transpose_c, h2))) return rc;
  }
  return 0;
}

int evc_geneig_prepare(evc_ctx* ctx, int N, const double* S, double* Linv, int* info) {
  EVC_REQUIRE(ctx && S && Linv && info, "evc_geneig_prepare: NULL argument");
  EVC_REQUIRE(N >= 1 && N <= 112, "evc_geneig_prepare: N=%d unsupported (1..112)", N);
  const size_t smem = static_cast<size_t>(N) * (N + 1) * sizeof(double);
  EVC_CHECK_CUDA(cudaFuncSetAttribute(cholesky_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  cholesky_inverse_kernel<<<1, 256, smem, ctx->stream>>>(N, S, Linv, info);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_geneig(evc_ctx* ctx, int nbatch, int N, const double* H, const double* Linv, int nroots,
               double* E, double* C) {
  EVC_REQUIRE(ctx && H && Linv && E && C, "evc_geneig: NULL argument");
  EVC_REQUIRE(N >= 1 && N <= 112, "evc_geneig: N=%d unsupported (1..112)", N);
  EVC_REQUIRE(nroots >= 1 && nroots <= N, "evc_geneig: nroots=%d out of range (1..%d)", nroots, N);
  if (nbatch <= 0) return 0;
  return launch_geneig(ctx, nbatch, N, 0, H, Linv, nroots, E, C);
}

}  // extern "C"
