// K6, lowest root only: H c = E S c for the ground state of the subspace problem
// (evcont/ab_initio_eigenvector_continuation.py:75-88: eigh(H, S), argmin).
//
// One TEAM of threads per geometry: a warp (no block-level barriers; large batches) or a whole
// 256-thread CTA (few geometries: the O(N^3) products and the Householder updates are spread over
// the CTA instead of 32 lanes -- 1.49 ms -> ~0.1 ms for one N = 100 problem):
//   A = L^-1 H L^-T                        (S = L L^T factored once per stack)
//   A = Q T Q^T                            Householder tridiagonalisation in shared memory
//   lambda_min(T)                          32-way multisection on Sturm counts (LAPACK dstebz recurrence)
//   (T - lambda) z = b                     inverse iteration, pivoted tridiagonal LU (dgttrf/dgttrs form)
//   E = z^T T z,  y = Q z,  c = L^-T y     (c^T S c = y^T y = 1)
// About 15x fewer instructions than the cyclic Jacobi sweep, which is kept for
// nroots > 1 (approximate_multistate) in dense.cu.
#include <cfloat>
#include <cstdlib>

#include "common.cuh"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// doubles of shared memory per warp
// shared-memory doubles of one problem: the CTA team keeps two N x (N+1) matrices (A, M), the warp team only M
// (its first triangular product reads H from global memory, the second runs in place), so that 4096 problems of
// N = 20 are resident at once (28 warps per SM x 5 KB) instead of running as two waves
__host__ __device__ inline size_t lowest_warp_doubles(int N, bool warp_team = false) {
  return (warp_team ? 1 : 2) * static_cast<size_t>(N) * (N + 1) + 10 * static_cast<size_t>(N) + 2;
}

#ifdef EVC_PHASE_TIMING
__device__ long long g_geneig_phase[16];
#define GEN_MARK(idx) do { if (TEAM == 32 && blockIdx.x == 0 && threadIdx.x == 0) g_geneig_phase[idx] = clock64(); } while (0)
#else
#define GEN_MARK(idx) do { } while (0)
#endif

template <int TEAM>
__global__ void __launch_bounds__(TEAM == 32 ? 128 : TEAM)
geneig_lowest_kernel(int N, int packed_lower, int nbatch, const double* __restrict__ H,
                     const double* __restrict__ Linv, double* __restrict__ E, double* __restrict__ C) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double tred[TEAM == 32 ? 1 : TEAM / 32];
  __shared__ double pw[TEAM == 32 ? 1 : TEAM / 32][TEAM == 32 ? 1 : 112];  // N <= 112
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
  const int b = TEAM == 32 ? blockIdx.x * wpc + warp : blockIdx.x;
  if (b >= nbatch) return;  // whole team
  const int lane = TEAM == 32 ? wl : static_cast<int>(threadIdx.x);  // index inside the team
  auto team_sync = [&]() {
    if (TEAM == 32) __syncwarp(); else __syncthreads();
  };
  // fixed-order sum over the team, result in every thread
  auto team_sum = [&](double v) {
    v = warp_sum(v);
    if (TEAM == 32) return v;
    __syncthreads();
    if (wl == 0) tred[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < (TEAM == 32 ? 1 : TEAM / 32); ++w) t += tred[w];
    return t;
  };
  const int ld = N + 1;
  double* A = sm + (TEAM == 32 ? static_cast<size_t>(warp) * lowest_warp_doubles(N, true) : 0);
  double* M = TEAM == 32 ? A : A + N * ld;
  double* d = M + N * ld;
  double* e = d + N;        // e[i] couples i and i+1
  double* tau = e + N;
  double* z = tau + N;
  double* w1 = z + N;
  double* dl = w1 + N;
  double* dd = dl + N;
  double* du = dd + N;
  double* du2 = du + N;
  int* ipiv = reinterpret_cast<int*>(du2 + N);  // N ints fit in N doubles

  GEN_MARK(0);
  if (TEAM != 32) {
    // ---- M <- lower triangle of H, mirrored ----
    if (packed_lower) {
      const double* Hb = H + static_cast<int64_t>(b) * (N * (N + 1) / 2);
      for (int k = lane; k < N * N; k += TEAM) {
        const int i = k / N, j = k - i * N;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        M[i * ld + j] = Hb[hi * (hi + 1) / 2 + lo];
      }
    } else {
      const double* Hb = H + static_cast<int64_t>(b) * N * N;
      for (int k = lane; k < N * N; k += TEAM) {
        const int i = k / N, j = k - i * N;
        M[i * ld + j] = (i >= j) ? Hb[i * N + j] : Hb[j * N + i];
      }
    }
    team_sync();
  }
  if (TEAM == 32) {
    // ---- M <- lower triangle of H, mirrored ----
    {
      const double* Hb = H + static_cast<int64_t>(b) * (packed_lower ? N * (N + 1) / 2 : N * N);
      for (int k = lane; k < N * N; k += TEAM) {
        const int i = k / N, j = k - i * N;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        M[i * ld + j] = __ldg(Hb + (packed_lower ? hi * (hi + 1) / 2 + lo : hi * N + lo));
      }
    }
    team_sync();
    // Both triangular products run IN PLACE on the one matrix the warp keeps in shared memory, with the
    // Linv element of every step uniform over the lanes (one broadcast load) and four independent partial
    // sums per lane (the dependent FP64 latency, not the flop count, is what a lone warp waits for).
    // ---- M <- Linv M: lanes own columns j, rows i descend (row i needs the old rows r <= i) ----
    for (int j0 = 0; j0 < N; j0 += TEAM) {
      const int j = j0 + lane;
      for (int i = N - 1; i >= 0; --i) {
        const double* li = Linv + i * N;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int r = 0;
        if (j < N) {
          for (; r + 3 <= i; r += 4) {
            a0 = fma(__ldg(li + r), M[r * ld + j], a0);
            a1 = fma(__ldg(li + r + 1), M[(r + 1) * ld + j], a1);
            a2 = fma(__ldg(li + r + 2), M[(r + 2) * ld + j], a2);
            a3 = fma(__ldg(li + r + 3), M[(r + 3) * ld + j], a3);
          }
          for (; r <= i; ++r) a0 = fma(__ldg(li + r), M[r * ld + j], a0);
          M[i * ld + j] = (a0 + a1) + (a2 + a3);
        }
      }
    }
    team_sync();
    // ---- M <- M Linv^T: lanes own rows i, columns j descend (entry j needs the old entries r <= j of the
    //      row); only j <= i is kept, the upper triangle is mirrored so that M stays exactly symmetric ----
    for (int i0 = 0; i0 < N; i0 += TEAM) {
      const int i = i0 + lane;
      double* row = M + (i < N ? i : 0) * ld;
      for (int j = N - 1; j >= 0; --j) {
        const double* lj = Linv + j * N;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int r = 0;
        if (i < N && j <= i) {
          for (; r + 3 <= j; r += 4) {
            a0 = fma(row[r], __ldg(lj + r), a0);
            a1 = fma(row[r + 1], __ldg(lj + r + 1), a1);
            a2 = fma(row[r + 2], __ldg(lj + r + 2), a2);
            a3 = fma(row[r + 3], __ldg(lj + r + 3), a3);
          }
          for (; r <= j; ++r) a0 = fma(row[r], __ldg(lj + r), a0);
          row[j] = (a0 + a1) + (a2 + a3);
        }
      }
    }
    team_sync();
    for (int k = lane; k < N * N; k += TEAM) {
      const int i = k / N, j = k - i * N;
      if (j > i) M[i * ld + j] = M[j * ld + i];
    }
    team_sync();
  } else {
    // CTA team: dst[i][:] = sum_{r <= i} Linv[i][r] src[r][:], one warp per row i, the row of Linv in
    // registers (broadcast by shuffle), lanes over the columns.  Pass 1 writes A^T = (Linv M)^T, pass 2
    // gives Linv A^T = (A Linv^T)^T = M (symmetric), N <= 128.
    constexpr int NWT = TEAM / 32;
    for (int pass = 0; pass < 2; ++pass) {
      const double* src = pass == 0 ? M : A;
      double* dst = pass == 0 ? A : M;
      for (int i = warp; i < N; i += NWT) {
        double lr[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) lr[q] = (wl + 32 * q <= i) ? __ldg(Linv + i * N + wl + 32 * q) : 0.0;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = 0; r <= i; ++r) {
          const double lv = __shfl_sync(0xffffffffu, lr[r >> 5], r & 31);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (wl + 32 * q < N) acc[q] = fma(lv, src[r * ld + wl + 32 * q], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (wl + 32 * q < N) {
            if (pass == 0) dst[(wl + 32 * q) * ld + i] = acc[q];   // transposed
            else dst[i * ld + wl + 32 * q] = acc[q];
          }
      }
      __syncthreads();
    }
  }

  GEN_MARK(1);
  // ---- Householder tridiagonalisation of M (both triangles kept up to date) ----
  for (int k = 0; k + 2 < N; ++k) {
    const int m = N - k - 1;
    double* col = M + (k + 1) * ld + k;  // u_i lives at col[i * ld]
    double part = 0.0;
    for (int i = 1 + lane; i < m; i += TEAM) {
      const double xv = col[i * ld];
      part += xv * xv;
    }
    const double sigma = team_sum(part);
    const double x0 = col[0];
    if (sigma == 0.0) {  // already tridiagonal in this column (warp-uniform)
      if (lane == 0) { e[k] = x0; tau[k] = 0.0; }
      team_sync();
      continue;
    }
    const double mu = sqrt(x0 * x0 + sigma);
    const double alpha = (x0 <= 0.0) ? mu : -mu;
    const double u0 = x0 - alpha;
    const double taup = 2.0 / (u0 * u0 + sigma);
    team_sync();
    if (lane == 0) { col[0] = u0; e[k] = alpha; tau[k] = taup; }
    team_sync();
    // p = taup * M22 u ;  pu = p . u
    double pu_part = 0.0;
    if (TEAM == 32) {
      for (int i = lane; i < m; i += TEAM) {
        const double* row = M + (k + 1 + i) * ld + k + 1;
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
        int j = 0;
        for (; j + 3 < m; j += 4) {
          c0 = fma(row[j], col[j * ld], c0);
          c1 = fma(row[j + 1], col[(j + 1) * ld], c1);
          c2 = fma(row[j + 2], col[(j + 2) * ld], c2);
          c3 = fma(row[j + 3], col[(j + 3) * ld], c3);
        }
        for (; j < m; ++j) c0 = fma(row[j], col[j * ld], c0);
        const double acc = (c0 + c1) + (c2 + c3);
        const double p = taup * acc;
        w1[i] = p;
        pu_part += p * col[i * ld];
      }
    } else {
      // M22 is symmetric: p_j = sum_i M22[i][j] u_i as column sums -- lanes own columns, the warps
      // split the rows, no shuffle reduction; the per-warp partials are combined in a fixed order
      for (int c0 = 0; c0 < m; c0 += 32) {
        const int j = c0 + wl;
        if (j < m) {
          double acc = 0.0;
          for (int i = warp; i < m; i += TEAM / 32) acc = fma(M[(k + 1 + i) * ld + k + 1 + j], col[i * ld], acc);
          pw[warp][j] = acc;
        }
      }
      __syncthreads();
      for (int j = lane; j < m; j += TEAM) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < TEAM / 32; ++w) t += pw[w][j];
        w1[j] = taup * t;
      }
      __syncthreads();
      for (int i = lane; i < m; i += TEAM) pu_part += w1[i] * col[i * ld];
    }
    const double pu = team_sum(pu_part);
    team_sync();
    // q = p - (taup/2) (p.u) u
    for (int i = lane; i < m; i += TEAM) w1[i] -= 0.5 * taup * pu * col[i * ld];
    team_sync();
    // M22 <- M22 - u q^T - q u^T
    if (TEAM == 32) {
      for (int i = lane; i < m; i += TEAM) {
        double* row = M + (k + 1 + i) * ld + k + 1;
        const double ui = col[i * ld], qi = w1[i];
#pragma unroll 4
        for (int j = 0; j < m; ++j) row[j] -= ui * w1[j] + qi * col[j * ld];
      }
    } else {
      for (int i = warp; i < m; i += TEAM / 32) {
        double* row = M + (k + 1 + i) * ld + k + 1;
        const double ui = col[i * ld], qi = w1[i];
        for (int j = wl; j < m; j += 32) row[j] -= ui * w1[j] + qi * col[j * ld];
      }
    }
    team_sync();
  }
  for (int i = lane; i < N; i += TEAM) d[i] = M[i * ld + i];
  if (lane == 0 && N >= 2) e[N - 2] = M[(N - 1) * ld + N - 2];
  team_sync();

  GEN_MARK(2);
  // ---- lowest eigenvalue of T = tridiag(e, d, e): multisection on Sturm counts (first warp) ----
  double lam = 0.0, tnorm = 0.0;
  if (TEAM == 32) {
    const int lane = wl;
  double glo = DBL_MAX, ghi = -DBL_MAX, emax = 0.0;
  for (int i = lane; i < N; i += 32) {
    const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i + 1 < N ? fabs(e[i]) : 0.0;
    glo = fmin(glo, d[i] - el - er);
    ghi = fmax(ghi, d[i] + el + er);
    emax = fmax(emax, er * er);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
    ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
    emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  }
  tnorm = fmax(fabs(glo), fabs(ghi));
  const double pivmin = DBL_MIN * fmax(1.0, emax);
  double lo = glo - 2.0 * DBL_EPSILON * tnorm * N - 2.0 * pivmin;
  double hi = ghi + 2.0 * DBL_EPSILON * tnorm * N + 2.0 * pivmin;
  for (int it = 0; it < 16; ++it) {
    const double width = hi - lo;
    if (width <= 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) break;
    const double h = width / 33.0;
    const double xs = lo + (lane + 1) * h;
    // Sturm count from the determinant recurrence p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (sign changes
    // = eigenvalues below x): two dependent FMAs per row instead of the FP64 division of the quotient
    // form, which is the longest dependent chain of this kernel; rescaled by 2^-+500 against over/underflow,
    // an exact zero takes the sign opposite to its predecessor (as the -pivmin rule of the quotient form)
    double pm = 1.0, pc = d[0] - xs;
    if (pc == 0.0) pc = -1.0e-100;
    int cnt = pc < 0.0;
    for (int i = 1; i < N; ++i) {
      const double ei = e[i - 1];
      double pn = fma(d[i] - xs, pc, -(ei * ei) * pm);
      if (pn == 0.0) pn = -pc * 1.0e-100;
      cnt += (pn < 0.0) != (pc < 0.0);
      pm = pc;
      pc = pn;
      const double mag = fabs(pc);
      if (mag > 3.2733906078961419e150) { pc *= 3.0549363634996047e-151; pm *= 3.0549363634996047e-151; }        // 2^500, 2^-500
      else if (mag < 3.0549363634996047e-151) { pc *= 3.2733906078961419e150; pm *= 3.2733906078961419e150; }
    }
    const unsigned ball = __ballot_sync(0xffffffffu, cnt >= 1);
    const int f = ball ? __ffs(ball) - 1 : 32;  // first sample point with an eigenvalue below it
    const double nlo = lo + f * h;
    hi = (f < 32) ? lo + (f + 1) * h : hi;
    lo = nlo;
  }
  lam = 0.5 * (lo + hi);

  }
  if (TEAM != 32) {
    // CTA team: TEAM-way multisection (every thread one sample point per round)
    double glo = DBL_MAX, ghi = -DBL_MAX, emax = 0.0;
    for (int i = wl; i < N; i += 32) {   // every warp redundantly: identical values everywhere
      const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i + 1 < N ? fabs(e[i]) : 0.0;
      glo = fmin(glo, d[i] - el - er);
      ghi = fmax(ghi, d[i] + el + er);
      emax = fmax(emax, er * er);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
      ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
      emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    tnorm = fmax(fabs(glo), fabs(ghi));
    const double pivmin = DBL_MIN * fmax(1.0, emax);
    double lo = glo - 2.0 * DBL_EPSILON * tnorm * N - 2.0 * pivmin;
    double hi = ghi + 2.0 * DBL_EPSILON * tnorm * N + 2.0 * pivmin;
    for (int it = 0; it < 8; ++it) {
      const double width = hi - lo;
      if (width <= 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) break;  // team-uniform
      const double h = width / (TEAM + 1);
      const double xs = lo + (lane + 1) * h;
      double q = d[0] - xs;
      if (fabs(q) < pivmin) q = -pivmin;
      int cnt = q < 0.0;
      for (int i = 1; i < N; ++i) {
        const double ei = e[i - 1];
        q = d[i] - xs - ei * ei / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0;
      }
      const unsigned ball = __ballot_sync(0xffffffffu, cnt >= 1);
      __syncthreads();
      if (wl == 0) tred[warp] = static_cast<double>(ball ? warp * 32 + __ffs(ball) - 1 : TEAM);
      __syncthreads();
      double fmin_ = static_cast<double>(TEAM);
#pragma unroll
      for (int w = 0; w < TEAM / 32; ++w) fmin_ = fmin(fmin_, tred[w]);
      const int f = static_cast<int>(fmin_);   // first sample point with an eigenvalue below it
      const double nlo = lo + f * h;
      hi = (f < TEAM) ? lo + (f + 1) * h : hi;
      lo = nlo;
    }
    lam = 0.5 * (lo + hi);
  }

  GEN_MARK(3);
  // ---- eigenvector of T by inverse iteration (lane 0; pivoted LU of T - lam I) ----
  if (lane == 0) {
    const double tiny = fmax(DBL_EPSILON * tnorm, DBL_MIN * 1e16);
    if (N == 1) {
      z[0] = 1.0;
    } else {
      for (int i = 0; i < N; ++i) { dd[i] = d[i] - lam; z[i] = 1.0; }
      for (int i = 0; i + 1 < N; ++i) { dl[i] = e[i]; du[i] = e[i]; du2[i] = 0.0; }
      for (int i = 0; i + 1 < N; ++i) {
        if (fabs(dd[i]) >= fabs(dl[i])) {
          if (dd[i] == 0.0) dd[i] = tiny;
          const double fact = dl[i] / dd[i];
          dl[i] = fact;
          dd[i + 1] -= fact * du[i];
          ipiv[i] = 0;
        } else {
          const double fact = dd[i] / dl[i];
          dd[i] = dl[i];
          dl[i] = fact;
          const double t = du[i];
          du[i] = dd[i + 1];
          dd[i + 1] = t - fact * dd[i + 1];
          if (i + 2 < N) {
            du2[i] = du[i + 1];
            du[i + 1] = -fact * du[i + 1];
          }
          ipiv[i] = 1;
        }
      }
      if (fabs(dd[N - 1]) < tiny) dd[N - 1] = (dd[N - 1] < 0.0) ? -tiny : tiny;
      for (int iter = 0; iter < 3; ++iter) {
        if (iter > 0) {  // forward substitution with the row interchanges
          for (int i = 0; i + 1 < N; ++i) {
            if (ipiv[i] == 0) {
              z[i + 1] -= dl[i] * z[i];
            } else {
              const double t = z[i];
              z[i] = z[i + 1];
              z[i + 1] = t - dl[i] * z[i];
            }
          }
        }
        z[N - 1] /= dd[N - 1];
        if (N > 1) z[N - 2] = (z[N - 2] - du[N - 2] * z[N - 1]) / dd[N - 2];
        for (int i = N - 3; i >= 0; --i) z[i] = (z[i] - du[i] * z[i + 1] - du2[i] * z[i + 2]) / dd[i];
        double big = 0.0;
        for (int i = 0; i < N; ++i) big = fmax(big, fabs(z[i]));
        const double sc = 1.0 / big;
        for (int i = 0; i < N; ++i) z[i] *= sc;
      }
    }
    double nn = 0.0;
    for (int i = 0; i < N; ++i) nn += z[i] * z[i];
    const double sc = 1.0 / sqrt(nn);
    for (int i = 0; i < N; ++i) z[i] *= sc;
    // Rayleigh quotient on T
    double rq = 0.0;
    for (int i = 0; i < N; ++i) {
      rq += d[i] * z[i] * z[i];
      if (i + 1 < N) rq += 2.0 * e[i] * z[i] * z[i + 1];
    }
    E[b] = rq;
  }
  team_sync();

  GEN_MARK(4);
  // ---- y = H_0 H_1 ... H_{N-3} z ----
  for (int k = N - 3; k >= 0; --k) {
    const double taup = tau[k];
    if (taup == 0.0) continue;
    const int m = N - k - 1;
    const double* col = M + (k + 1) * ld + k;
    double part = 0.0;
    for (int i = lane; i < m; i += TEAM) part += col[i * ld] * z[k + 1 + i];
    const double s = taup * team_sum(part);
    for (int i = lane; i < m; i += TEAM) z[k + 1 + i] -= s * col[i * ld];
    team_sync();
  }
  GEN_MARK(5);
  // ---- c = Linv^T y ----
  double* Cb = C + static_cast<int64_t>(b) * N;
  for (int i = lane; i < N; i += TEAM) {
    double acc = 0.0;
    for (int r = i; r < N; ++r) acc += __ldg(Linv + r * N + i) * z[r];
    Cb[i] = acc;
  }
  GEN_MARK(6);
}

}  // namespace

#ifdef EVC_PHASE_TIMING
extern "C" int evc_debug_geneig_clocks(long long* out_host) {   // development aid, not part of the ABI
  EVC_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_geneig_phase, sizeof(long long) * 16));
  return 0;
}
#endif

int evc_launch_geneig_lowest(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H,
                             const double* Linv, double* E, double* C) {
  // N <= 24: the register-resident form (geneig_reg.cu); EVC_GENEIG_REG=0 in the environment keeps the
  // shared-memory kernels (A/B timing of the two forms in one build; development aid)
  static const bool use_reg = [] {
    const char* e = getenv("EVC_GENEIG_REG");
    return !(e && e[0] == '0');
  }();
  if (use_reg && evc_geneig_reg_supported(N)) return evc_geneig_reg(ctx, nbatch, N, packed_lower, H, Linv, E, C);
  const size_t per_warp = lowest_warp_doubles(N) * sizeof(double);
  const size_t per_warp32 = lowest_warp_doubles(N, true) * sizeof(double);
  EVC_REQUIRE(per_warp <= ctx->smem_optin, "geneig: N=%d needs %zu bytes of shared memory", N, per_warp);
  if (nbatch <= 2 * ctx->sm_count) {
    // few geometries: one 256-thread CTA each
    EVC_CHECK_CUDA(cudaFuncSetAttribute(geneig_lowest_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(per_warp)));
    geneig_lowest_kernel<256><<<nbatch, 256, per_warp, ctx->stream>>>(N, packed_lower, nbatch, H, Linv, E, C);
    EVC_CHECK_LAUNCH();
    return 0;
  }
  int wpc = 4;
  while (wpc > 1 && (wpc * per_warp32 > ctx->smem_optin || wpc * per_warp32 > 48 * 1024)) wpc >>= 1;
  const size_t smem = wpc * per_warp32;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(geneig_lowest_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  geneig_lowest_kernel<32><<<(nbatch + wpc - 1) / wpc, wpc * 32, smem, ctx->stream>>>(N, packed_lower, nbatch, H, Linv,
                                                                                      E, C);
  EVC_CHECK_LAUNCH();
  return 0;
}
