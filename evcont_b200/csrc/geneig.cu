// K6, lowest root only: H c = E S c for the ground state of the subspace problem
// (evcont/ab_initio_eigenvector_continuation.py:75-88: eigh(H, S), argmin).
//
// One WARP per geometry, no block-level barriers:
//   A = L^-1 H L^-T                        (S = L L^T factored once per stack)
//   A = Q T Q^T                            Householder tridiagonalisation in shared memory
//   lambda_min(T)                          32-way multisection on Sturm counts (LAPACK dstebz recurrence)
//   (T - lambda) z = b                     inverse iteration, pivoted tridiagonal LU (dgttrf/dgttrs form)
//   E = z^T T z,  y = Q z,  c = L^-T y     (c^T S c = y^T y = 1)
// About 15x fewer instructions than the cyclic Jacobi sweep, which is kept for
// nroots > 1 (approximate_multistate) in dense.cu.
#include <cfloat>

#include "common.cuh"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// doubles of shared memory per warp
__host__ __device__ inline size_t lowest_warp_doubles(int N) {
  return 2 * static_cast<size_t>(N) * (N + 1) + 10 * static_cast<size_t>(N) + 2;
}

__global__ void __launch_bounds__(128)
geneig_lowest_kernel(int N, int packed_lower, int nbatch, const double* __restrict__ H,
                     const double* __restrict__ Linv, double* __restrict__ E, double* __restrict__ C) {
  extern __shared__ __align__(16) double sm[];
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * wpc + warp;
  if (b >= nbatch) return;  // whole warp; no block barriers below
  const int ld = N + 1;
  double* A = sm + static_cast<size_t>(warp) * lowest_warp_doubles(N);
  double* M = A + N * ld;
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

  // ---- M <- lower triangle of H, mirrored ----
  if (packed_lower) {
    const double* Hb = H + static_cast<int64_t>(b) * (N * (N + 1) / 2);
    for (int k = lane; k < N * N; k += 32) {
      const int i = k / N, j = k - i * N;
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      M[i * ld + j] = Hb[hi * (hi + 1) / 2 + lo];
    }
  } else {
    const double* Hb = H + static_cast<int64_t>(b) * N * N;
    for (int k = lane; k < N * N; k += 32) {
      const int i = k / N, j = k - i * N;
      M[i * ld + j] = (i >= j) ? Hb[i * N + j] : Hb[j * N + i];
    }
  }
  __syncwarp();
  // ---- A <- Linv M   (Linv lower triangular) ----
  for (int k = lane; k < N * N; k += 32) {
    const int i = k / N, j = k - i * N;
    double acc = 0.0;
    for (int r = 0; r <= i; ++r) acc += __ldg(Linv + i * N + r) * M[r * ld + j];
    A[i * ld + j] = acc;
  }
  __syncwarp();
  // ---- M <- A Linv^T, lower triangle computed and mirrored ----
  for (int k = lane; k < N * N; k += 32) {
    const int i = k / N, j = k - i * N;
    if (j > i) continue;
    double acc = 0.0;
    for (int r = 0; r <= j; ++r) acc += A[i * ld + r] * __ldg(Linv + j * N + r);
    M[i * ld + j] = acc;
    M[j * ld + i] = acc;
  }
  __syncwarp();

  // ---- Householder tridiagonalisation of M (both triangles kept up to date) ----
  for (int k = 0; k + 2 < N; ++k) {
    const int m = N - k - 1;
    double* col = M + (k + 1) * ld + k;  // u_i lives at col[i * ld]
    double part = 0.0;
    for (int i = 1 + lane; i < m; i += 32) {
      const double xv = col[i * ld];
      part += xv * xv;
    }
    const double sigma = warp_sum(part);
    const double x0 = col[0];
    if (sigma == 0.0) {  // already tridiagonal in this column (warp-uniform)
      if (lane == 0) { e[k] = x0; tau[k] = 0.0; }
      __syncwarp();
      continue;
    }
    const double mu = sqrt(x0 * x0 + sigma);
    const double alpha = (x0 <= 0.0) ? mu : -mu;
    const double u0 = x0 - alpha;
    const double taup = 2.0 / (u0 * u0 + sigma);
    __syncwarp();
    if (lane == 0) { col[0] = u0; e[k] = alpha; tau[k] = taup; }
    __syncwarp();
    // p = taup * M22 u ;  pu = p . u
    double pu_part = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double* row = M + (k + 1 + i) * ld + k + 1;
      double acc = 0.0;
      for (int j = 0; j < m; ++j) acc += row[j] * col[j * ld];
      const double p = taup * acc;
      w1[i] = p;
      pu_part += p * col[i * ld];
    }
    const double pu = warp_sum(pu_part);
    __syncwarp();
    // q = p - (taup/2) (p.u) u
    for (int i = lane; i < m; i += 32) w1[i] -= 0.5 * taup * pu * col[i * ld];
    __syncwarp();
    // M22 <- M22 - u q^T - q u^T
    for (int i = lane; i < m; i += 32) {
      double* row = M + (k + 1 + i) * ld + k + 1;
      const double ui = col[i * ld], qi = w1[i];
      for (int j = 0; j < m; ++j) row[j] -= ui * w1[j] + qi * col[j * ld];
    }
    __syncwarp();
  }
  for (int i = lane; i < N; i += 32) d[i] = M[i * ld + i];
  if (lane == 0 && N >= 2) e[N - 2] = M[(N - 1) * ld + N - 2];
  __syncwarp();

  // ---- lowest eigenvalue of T = tridiag(e, d, e): multisection on Sturm counts ----
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
  const double tnorm = fmax(fabs(glo), fabs(ghi));
  const double pivmin = DBL_MIN * fmax(1.0, emax);
  double lo = glo - 2.0 * DBL_EPSILON * tnorm * N - 2.0 * pivmin;
  double hi = ghi + 2.0 * DBL_EPSILON * tnorm * N + 2.0 * pivmin;
  for (int it = 0; it < 16; ++it) {
    const double width = hi - lo;
    if (width <= 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) break;
    const double h = width / 33.0;
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
    const int f = ball ? __ffs(ball) - 1 : 32;  // first sample point with an eigenvalue below it
    const double nlo = lo + f * h;
    hi = (f < 32) ? lo + (f + 1) * h : hi;
    lo = nlo;
  }
  const double lam = 0.5 * (lo + hi);

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
  __syncwarp();

  // ---- y = H_0 H_1 ... H_{N-3} z ----
  for (int k = N - 3; k >= 0; --k) {
    const double taup = tau[k];
    if (taup == 0.0) continue;
    const int m = N - k - 1;
    const double* col = M + (k + 1) * ld + k;
    double part = 0.0;
    for (int i = lane; i < m; i += 32) part += col[i * ld] * z[k + 1 + i];
    const double s = taup * warp_sum(part);
    for (int i = lane; i < m; i += 32) z[k + 1 + i] -= s * col[i * ld];
    __syncwarp();
  }
  // ---- c = Linv^T y ----
  double* Cb = C + static_cast<int64_t>(b) * N;
  for (int i = lane; i < N; i += 32) {
    double acc = 0.0;
    for (int r = i; r < N; ++r) acc += __ldg(Linv + r * N + i) * z[r];
    Cb[i] = acc;
  }
}

}  // namespace

int evc_launch_geneig_lowest(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H,
                             const double* Linv, double* E, double* C) {
  const size_t per_warp = lowest_warp_doubles(N) * sizeof(double);
  EVC_REQUIRE(per_warp <= ctx->smem_optin, "geneig: N=%d needs %zu bytes of shared memory", N, per_warp);
  int wpc = 4;
  while (wpc > 1 && (wpc * per_warp > ctx->smem_optin || wpc * per_warp > 48 * 1024)) wpc >>= 1;
  const size_t smem = wpc * per_warp;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(geneig_lowest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  geneig_lowest_kernel<<<(nbatch + wpc - 1) / wpc, wpc * 32, smem, ctx->stream>>>(N, packed_lower, nbatch, H, Linv, E, C);
  EVC_CHECK_LAUNCH();
  return 0;
}
