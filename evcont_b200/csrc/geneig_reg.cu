// K6, lowest root, with the whole eigensolver of one problem in the registers of one warp (N <= 24).
//
// Reference anchor: evcont/ab_initio_eigenvector_continuation.py:75-88 (scipy.linalg.eigh(H, S), argmin).
//
// Same algorithm as geneig_lowest_kernel (geneig.cu) -- A = L^-1 H L^-T, Householder tridiagonalisation,
// 32-way multisection on Sturm counts, inverse iteration with a pivoted tridiagonal LU, back-transform,
// c = L^-T y -- but N is a template parameter and lane j owns column j (= row j) of the working matrix in
// registers: every register index is a compile-time constant (all loops unrolled), the only communication
// is warp shuffles, and shared memory is touched three times (two transposes of the product phase, the
// tridiagonal for the Sturm counts).  The shared-memory kernel needs 28.6 k warp instructions and ~260 k
// cycles per N = 20 problem, most of them dependent shared-memory round trips with run-time addressing
// (profiles/r01d_geneig_phase_clocks.txt); this one 15.7 k instructions and 0.046 ms for one problem
// (profiles/r02_packed_step_ncu_full.txt).
//
// Symmetry is kept EXACT through the reduction (lower triangle mirrored after the products, rank-2 update as
// the commutative sum of two rounded products): lane j reads the component of the Householder vector that
// belongs to it from its own register a[k] = A[k][j] instead of a lane-dependent register of lane k.
#include <cfloat>

#include "common.cuh"

namespace {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double bc(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// problems per CTA; 7 CTAs = 28 warps per SM hold all 4096 problems of the bench at once (72 registers per thread:
// the eigensolver stage of the step 0.133 -> 0.104 ms against two waves at 124 registers)
constexpr int kGrWarps = 4;

template <int N>
struct GrGeom {
  static constexpr int ld = N + 1;                       // odd or even: columns are read with stride ld
  static constexpr int per_warp = N * ld + 2 * N;        // transpose buffer | d | e^2
  static constexpr size_t smem = (static_cast<size_t>(N) * N + kGrWarps * per_warp) * sizeof(double);
};

template <int N>
__global__ void __launch_bounds__(kGrWarps * 32, 7)
geneig_reg_kernel(int packed_lower, int nbatch, const double* __restrict__ H, const double* __restrict__ Linv,
                  double* __restrict__ E, double* __restrict__ C) {
  using G = GrGeom<N>;
  constexpr int ld = G::ld;
  extern __shared__ __align__(16) double sm[];
  double* Ls = sm;  // [N][N], shared by the CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* As = sm + N * N + warp * G::per_warp;
  double* ds = As + N * ld;
  double* e2s = ds + N;
  for (int k = threadIdx.x; k < N * N; k += blockDim.x) Ls[k] = __ldg(Linv + k);
  __syncthreads();
  const int b = blockIdx.x * kGrWarps + warp;
  if (b >= nbatch) return;  // whole warp
  const bool in = lane < N;
  const int j = in ? lane : N - 1;  // idle lanes shadow the last column (no stores)

  // ---- column j of H (lower triangle mirrored) ----
  double a[N];
  {
    const double* Hb = H + static_cast<int64_t>(b) * (packed_lower ? N * (N + 1) / 2 : N * N);
#pragma unroll
    for (int r = 0; r < N; ++r) {
      const int hi = r > j ? r : j, lo = r > j ? j : r;
      a[r] = __ldg(Hb + (packed_lower ? hi * (hi + 1) / 2 + lo : hi * N + lo));
    }
  }
  // ---- C = Linv H: column j of C is Linv times column j of H (Linv uniform over the lanes) ----
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r <= i; ++r) acc = fma(Ls[i * N + r], a[r], acc);
    a[i] = acc;  // rows descend: a[r], r <= i, are still the old entries
  }
  // ---- A = C Linv^T: A[i][j] = A[j][i] = sum_{r <= i} Linv[i][r] C[j][r]: lane j needs ROW j of C ----
  if (in) {
#pragma unroll
    for (int i = 0; i < N; ++i) As[i * ld + j] = a[i];  // As[i][j] = C[i][j]
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < N; ++r) a[r] = As[j * ld + r];    // row j of C
  __syncwarp();
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r <= i; ++r) acc = fma(Ls[i * N + r], a[r], acc);
    a[i] = acc;  // = A[j][i], taken as A[i][j]
  }
  // exact symmetry: the lower triangle is authoritative (A[i][j], i < j, is replaced by lane i's A[j][i])
  if (in) {
#pragma unroll
    for (int i = 0; i < N; ++i) As[i * ld + j] = a[i];
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double t = As[j * ld + i];  // lane i's entry for row j
    a[i] = i < j ? t : a[i];
  }
  __syncwarp();

  // ---- Householder tridiagonalisation; lane k keeps e_k and tau_k, lane j > k keeps its component of the
  //      k-th reflector in a[k] (row k of the matrix is finished by then) ----
  double ek = 0.0, tk = 0.0;
#pragma unroll
  for (int k = 0; k + 2 < N; ++k) {
    // column k below the diagonal lives in lane k (which is idle from here on: its registers do not change during
    // the step), so its entries are broadcast where they are used instead of being kept (40 registers at N = 20)
    double sigma = 0.0, p = 0.0;
#pragma unroll
    for (int i = k + 2; i < N; ++i) {
      const double ui = bc(a[i], k);
      sigma = fma(ui, ui, sigma);
      p = fma(a[i], ui, p);
    }
    const double x0 = bc(a[k + 1], k);
    if (sigma == 0.0) {  // already tridiagonal in this column (warp-uniform)
      if (lane == k) { ek = x0; tk = 0.0; }
      continue;
    }
    const double mu = sqrt(fma(x0, x0, sigma));
    const double alpha = (x0 <= 0.0) ? mu : -mu;
    const double u0 = x0 - alpha;
    const double taup = 2.0 / fma(u0, u0, sigma);
    if (lane == k) { ek = alpha; tk = taup; }
    const bool act = in && lane > k;
    // p = taup A22 u (column sums are local), own component of u from the own register a[k]
    p = taup * fma(a[k + 1], u0, p);
    // (zero in the lanes that take no part: their updates below subtract an exact zero, no select per element)
    const double uo = act ? ((lane == k + 1) ? u0 : a[k]) : 0.0;
    const double pu = wsum(p * uo);
    const double q = act ? fma(-0.5 * taup * pu, uo, p) : 0.0;
    // A22 <- A22 - u q^T - q u^T, as the commutative sum of two rounded products (exactly symmetric)
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double ui = (i == k + 1) ? u0 : bc(a[i], k);
      const double qi = bc(q, i);
      a[i] -= __dadd_rn(__dmul_rn(ui, q), __dmul_rn(qi, uo));
    }
    a[k] = act ? uo : a[k];
  }
  if (N >= 2 && lane == N - 2) ek = a[N - 1];
  double dj = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) dj = (i == j) ? a[i] : dj;
  if (lane >= N - 1) ek = 0.0;  // e_{N-1} does not exist; idle lanes carry zeros
  if (!in) dj = 0.0;

  // ---- lowest eigenvalue of T = tridiag(e, d, e): 32-way multisection on Sturm counts ----
  double lam, tnorm;
  {
    const double el = __shfl_up_sync(0xffffffffu, fabs(ek), 1);
    const double eleft = lane > 0 ? el : 0.0, eright = fabs(ek);
    double glo = in ? dj - eleft - eright : DBL_MAX, ghi = in ? dj + eleft + eright : -DBL_MAX;
    double emax = eright * eright;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
      ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
      emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    tnorm = fmax(fabs(glo), fabs(ghi));
    // The counts run on T scaled by an exact power of two to a norm in [1/2, 1): the determinant recurrence
    //   p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}      (sign changes = eigenvalues below x)
    // then cannot overflow within N <= 24 steps (|p_i| <= 3^i), and only underflow has to be watched.  Signs and
    // magnitudes are read off the high word with integer instructions: FP64 compares share the FP64 pipe with the
    // arithmetic, and this loop is 40 % of the kernel's instructions.
    int ex2 = 0;
    (void)frexp(tnorm, &ex2);
    const double sc = tnorm > 0.0 ? ldexp(1.0, -ex2) : 1.0, sci = tnorm > 0.0 ? ldexp(1.0, ex2) : 1.0;
    if (in) { ds[lane] = dj * sc; e2s[lane] = -(ek * sc) * (ek * sc); }
    __syncwarp();
    const double tn = tnorm * sc;
    const double pivmin = DBL_MIN * fmax(1.0, emax * sc * sc);
    double lo = glo * sc - 2.0 * DBL_EPSILON * tn * N - 2.0 * pivmin;
    double hi = ghi * sc + 2.0 * DBL_EPSILON * tn * N + 2.0 * pivmin;
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
      const double width = hi - lo;
      if (width <= 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) break;
      const double h = width / 33.0;
      const double xs = lo + (lane + 1) * h;
      double pm = 1.0, pc = ds[0] - xs;
      if (pc == 0.0) pc = -1.0e-100;  // an exact zero takes the sign opposite to its predecessor (p_{-1} = 1)
      int hc = __double2hiint(pc);
      int cnt = static_cast<unsigned>(hc) >> 31;
#pragma unroll
      for (int i = 1; i < N; ++i) {
        double pn = fma(ds[i] - xs, pc, e2s[i - 1] * pm);
        int hn = __double2hiint(pn);
        if (((hn >> 20) & 0x7ff) < 523) {  // |p| < 2^-500 (or zero): rare
          if (pn == 0.0) {
            pn = -pc * 1.0e-100;
          } else {
            pn *= 3.2733906078961419e150;  // 2^500
            pc *= 3.2733906078961419e150;
          }
          hn = __double2hiint(pn);
        }
        cnt += static_cast<unsigned>(hn ^ hc) >> 31;
        pm = pc;
        pc = pn;
        hc = hn;
      }
      const unsigned ball = __ballot_sync(0xffffffffu, cnt >= 1);
      const int f = ball ? __ffs(ball) - 1 : 32;  // first sample point with an eigenvalue below it
      const double nlo = lo + f * h;
      hi = (f < 32) ? lo + (f + 1) * h : hi;
      lo = nlo;
    }
    lam = 0.5 * (lo + hi) * sci;
  }

  // ---- eigenvector of T by inverse iteration: pivoted LU of T - lam I (dgttrf / dgttrs form), element i of
  //      every band in lane i, the elimination walks down the lanes ----
  double z = 1.0;
  if (N > 1) {
    const double tiny = fmax(DBL_EPSILON * tnorm, DBL_MIN * 1e16);
    double dd = dj - lam, dl = ek, du = ek, du2 = 0.0;
    int piv = 0;
#pragma unroll
    for (int i = 0; i + 1 < N; ++i) {
      double ddi = bc(dd, i);
      const double dli = bc(dl, i), dui = bc(du, i), ddn = bc(dd, i + 1), dun = bc(du, i + 1);
      const bool nopiv = fabs(ddi) >= fabs(dli);
      if (nopiv && ddi == 0.0) ddi = tiny;
      const double fact = nopiv ? dli / ddi : ddi / dli;
      const double n_dd_i = nopiv ? ddi : dli;
      const double n_du_i = nopiv ? dui : ddn;
      const double n_dd_n = nopiv ? fma(-fact, dui, ddn) : fma(-fact, ddn, dui);
      const double n_du2_i = (!nopiv && i + 2 < N) ? dun : 0.0;
      const double n_du_n = (!nopiv && i + 2 < N) ? -fact * dun : dun;
      if (lane == i) { dd = n_dd_i; dl = fact; du = n_du_i; du2 = n_du2_i; piv = nopiv ? 0 : 1; }
      if (lane == i + 1) { dd = n_dd_n; du = n_du_n; }
    }
    if (lane == N - 1 && fabs(dd) < tiny) dd = (dd < 0.0) ? -tiny : tiny;
    const double rdd = 1.0 / (in ? dd : 1.0);
    for (int iter = 0; iter < 3; ++iter) {
      if (iter > 0) {  // forward substitution with the row interchanges
#pragma unroll
        for (int i = 0; i + 1 < N; ++i) {
          const double zi = bc(z, i), zn = bc(z, i + 1), dli = bc(dl, i);
          const int pv = __shfl_sync(0xffffffffu, piv, i);
          if (lane == i) z = pv ? zn : zi;
          if (lane == i + 1) z = pv ? fma(-dli, zn, zi) : fma(-dli, zi, zn);
        }
      }
#pragma unroll
      for (int i = N - 1; i >= 0; --i) {
        const double z1 = i + 1 < N ? bc(z, i + 1) : 0.0, z2 = i + 2 < N ? bc(z, i + 2) : 0.0;
        if (lane == i) z = (z - du * z1 - du2 * z2) * rdd;
      }
      double big = in ? fabs(z) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) big = fmax(big, __shfl_xor_sync(0xffffffffu, big, o));
      z *= 1.0 / big;
    }
  }
  if (!in) z = 0.0;
  z *= 1.0 / sqrt(wsum(z * z));
  // Rayleigh quotient on T
  {
    const double zn = __shfl_down_sync(0xffffffffu, z, 1);
    const double rq = wsum(in ? fma(dj * z, z, (lane + 1 < N) ? 2.0 * ek * z * zn : 0.0) : 0.0);
    if (lane == 0) E[b] = rq;
  }

  // ---- y = H_0 H_1 ... H_{N-3} z ----
#pragma unroll
  for (int k = N - 3; k >= 0; --k) {
    const double taup = bc(tk, k);
    if (taup == 0.0) continue;  // warp-uniform
    const bool act = in && lane > k;
    const double s = taup * wsum(act ? a[k] * z : 0.0);
    z = act ? fma(-s, a[k], z) : z;
  }
  // ---- c = Linv^T y ----
  {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      const double yr = bc(z, r);
      acc = (r >= j) ? fma(Ls[r * N + j], yr, acc) : acc;
    }
    if (in) C[static_cast<int64_t>(b) * N + lane] = acc;
  }
}

template <int N>
int launch_geneig_reg(evc_ctx* ctx, int nbatch, int packed_lower, const double* H, const double* Linv, double* E,
                      double* C) {
  using G = GrGeom<N>;
  geneig_reg_kernel<N><<<(nbatch + kGrWarps - 1) / kGrWarps, kGrWarps * 32, G::smem, ctx->stream>>>(
      packed_lower, nbatch, H, Linv, E, C);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool evc_geneig_reg_supported(int N) { return N >= 2 && N <= 24; }

int evc_geneig_reg(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H, const double* Linv, double* E,
                   double* C) {
  switch (N) {
#define EVC_GR_CASE(N_) case N_: return launch_geneig_reg<N_>(ctx, nbatch, packed_lower, H, Linv, E, C);
    EVC_GR_CASE(2) EVC_GR_CASE(3) EVC_GR_CASE(4) EVC_GR_CASE(5) EVC_GR_CASE(6) EVC_GR_CASE(7) EVC_GR_CASE(8)
    EVC_GR_CASE(9) EVC_GR_CASE(10) EVC_GR_CASE(11) EVC_GR_CASE(12) EVC_GR_CASE(13) EVC_GR_CASE(14) EVC_GR_CASE(15)
    EVC_GR_CASE(16) EVC_GR_CASE(17) EVC_GR_CASE(18) EVC_GR_CASE(19) EVC_GR_CASE(20) EVC_GR_CASE(21) EVC_GR_CASE(22)
    EVC_GR_CASE(23) EVC_GR_CASE(24)
#undef EVC_GR_CASE
    default: break;
  }
  EVC_REQUIRE(false, "evc_geneig_reg: N=%d unsupported (2..24)", N);
  return -1;
}
