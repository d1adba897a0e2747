// K9g: AO integrals over contracted Cartesian s and p Gaussians (6-31G H / O: the H2O and Zundel
// configurations) for a batch of geometries, McMurchie-Davidson scheme.  Same outputs and reference
// call sites as integrals.cu (evcont/ab_initio_gradients_loewdin.py:25, 130, 147, 177, 283-284,
// 338-339, 370, 378); s-only molecules keep the specialised kernel there.
//
//   G_i(x; a, A) G_j(x; b, B) = K sum_t E^{ij}_t Lambda_t(x; p, P),   K = exp(-mu X_AB^2)
//     raising i:  E'_t = E_{t-1} / (2p) + X_PA E_t + (t + 1) E_{t+1}        (j: X_PB)
//   (ab|cd) = 2 pi^2.5 / (p q sqrt(p+q)) K_ab K_cd sum_{tuv} E^{ab}_{tuv} sum_{t'u'v'} (-1)^{t'+u'+v'}
//             E^{cd}_{t'u'v'} R_{t+t',u+u',v+v'}(rho, P - Q),
//     R^n_000 = (-2 rho)^n F_n(rho |PQ|^2),  R^n_{t+1,u,v} = t R^{n+1}_{t-1,u,v} + X_PQ R^{n+1}_{tuv}
//   d/dA_x of a Cartesian Gaussian: 2a G_{l+1} - l G_{l-1};  (nabla a b|cd) = -d/dA (ab|cd).
//
// Two-electron part: gclass.cu (class kernels over contracted SHELL quartets, everything in registers;
// one lane group of 1..32 lanes, sized to the number of primitive quartets, per shell quartet; value and
// the derivatives with respect to the centres of a, b and c from shifted angular momenta over one R table,
// the fourth centre from translational invariance; fixed butterfly reduction; the eight index permutations
// written by the lanes of the group).  This file: tables, work lists, the one-electron kernel, the C ABI.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

#include "integrals_sp.cuh"

namespace {
using namespace evc_gint;

void boys_host_g(int mmax, long double t, long double* out) {
  const long double et = expl(-t);
  long double term = 1.0L / (2 * mmax + 1), acc = term;
  for (int k = 1; k < 500; ++k) {
    term *= 2.0L * t / (2 * mmax + 2 * k + 1);
    acc += term;
    if (term < acc * 1e-22L) break;
  }
  out[mmax] = et * acc;
  for (int m = mmax; m > 0; --m) out[m - 1] = (2.0L * t * out[m] + et) / (2 * m - 1);
}

__device__ __forceinline__ double gwarp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// F_0 .. F_L at T.  T < Tmax: F_L..F_{L+5} at the grid point by downward recursion from the tabulated
// F_11, 5th-order Taylor series for F_L(T), exp(-T) = exp(-T0) exp(T0 - T), downward recursion in T.
// T >= Tmax: F_0 = sqrt(pi / T) / 2 and the upward recursion with exp(-T) (stable for T > m).
__device__ __noinline__ void boys_upto(int L, double T, const double* __restrict__ tab, double* F) {
  if (T < static_cast<double>(kGTmax)) {
    const double r = fma(T, static_cast<double>(kGPerUnit), 6755399441055744.0);
    const int i = __double2loint(r);
    const double t0 = (r - 6755399441055744.0) * (1.0 / kGPerUnit);
    const double d = t0 - T, tt = t0 + t0;
    const double2 fe = *reinterpret_cast<const double2*>(tab + 2 * i);
    double g[kGTop + 1];
    g[kGTop] = fe.x;
    for (int m = kGTop; m > L; --m) g[m - 1] = fma(tt, g[m], fe.y) / static_cast<double>(2 * m - 1);
    // F_L(T) = sum_k F_{L+k}(T0) d^k / k!
    double fl = g[L + 5];
    fl = fma(fl, d * 0.2, g[L + 4]);
    fl = fma(fl, d * 0.25, g[L + 3]);
    fl = fma(fl, d * (1.0 / 3.0), g[L + 2]);
    fl = fma(fl, d * 0.5, g[L + 1]);
    fl = fma(fl, d, g[L]);
    // exp(-T) = exp(-T0) exp(d), |d| <= 1/128
    const double ed = 1.0 + d * (1.0 + d * (0.5 + d * (1.0 / 6.0 + d * (1.0 / 24.0 + d * (1.0 / 120.0 + d * (1.0 / 720.0))))));
    const double et = fe.y * ed, t2 = T + T;
    F[L] = fl;
    for (int m = L; m > 0; --m) F[m - 1] = fma(t2, F[m], et) / static_cast<double>(2 * m - 1);
  } else {
    const double ri = 1.0 / T, et = exp(-T);
    F[0] = 0.88622692545275801365 * sqrt(ri);
    for (int m = 0; m < L; ++m) F[m + 1] = (static_cast<double>(2 * m + 1) * F[m] - et) * (0.5 * ri);
  }
}

// Hermite coefficients E^{ij}_t, t = 0..i+j (without the exp(-mu X_AB^2) factor); E has room for 8
__device__ __noinline__ void herm_E(int i, int j, double xpa, double xpb, double h, double* E) {
  E[0] = 1.0;
  int deg = 0;
  for (int s = 0; s < i + j; ++s) {
    const double x = s < i ? xpa : xpb;
    double prev = 0.0;  // E_{t-1} of the old polynomial
    for (int t = 0; t <= deg; ++t) {
      const double cur = E[t];
      const double nxt = t + 1 <= deg ? E[t + 1] : 0.0;
      E[t] = h * prev + x * cur + static_cast<double>(t + 1) * nxt;
      prev = cur;
    }
    E[deg + 1] = h * prev;
    ++deg;
  }
}

// compact index of (t, u, v), t + u + v <= 6 (84 entries, ordered by t, then u, then v).  All lanes of a
// warp walk the same (t, u, v) at the same time, so the table sits in constant memory (one broadcast).
__constant__ unsigned char c_ridx[7][7][7];
__device__ __forceinline__ int ridx(int t, int u, int v) { return c_ridx[t][u][v]; }

// R^0_{tuv}(alpha, X) for t + u + v <= L (L <= 6) into R (84 entries), with scratch S (84)
__device__ __noinline__ void build_R(int L, double alpha, double X, double Y, double Z, const double* F, double* R, double* S) {
  double* A = S;   // level n + 1
  double* B = R;   // level n
  // make sure the final level lands in R: levels L, L-1, ..., 0 alternate; start so that n = 0 writes R
  if (L & 1) { A = R; B = S; }
  double pw = 1.0;
  double m2a[kGMaxL + 1];
  for (int n = 0; n <= L; ++n) { m2a[n] = pw; pw *= -2.0 * alpha; }
  for (int n = L; n >= 0; --n) {
    const int ord = L - n;
    B[ridx(0, 0, 0)] = m2a[n] * F[n];
    for (int t = 0; t <= ord; ++t)
      for (int u = 0; u + t <= ord; ++u)
        for (int v = (t + u == 0 ? 1 : 0); v + t + u <= ord; ++v) {
          double val;
          if (t > 0) val = (t > 1 ? static_cast<double>(t - 1) * A[ridx(t - 2, u, v)] : 0.0) + X * A[ridx(t - 1, u, v)];
          else if (u > 0) val = (u > 1 ? static_cast<double>(u - 1) * A[ridx(t, u - 2, v)] : 0.0) + Y * A[ridx(t, u - 1, v)];
          else val = (v > 1 ? static_cast<double>(v - 1) * A[ridx(t, u, v - 2)] : 0.0) + Z * A[ridx(t, u, v - 1)];
          B[ridx(t, u, v)] = val;
        }
    double* tmp = A; A = B; B = tmp;
  }
}

// ---- one-electron part: one warp per ORDERED pair (a, b), lanes over primitive pairs -----------------
// overlap-type primitives <G_la|G_lb> / (pi/p)^1.5 K as products of E^{ij}_0
__device__ __forceinline__ double ovl3(const int* la, const int* lb, const double* xpa, const double* xpb, double hp) {
  double r = 1.0, E[8];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    herm_E(la[d], lb[d], xpa[d], xpb[d], hp, E);
    r *= E[0];
  }
  return r;
}

// <G_la| -1/2 nabla^2 |G_lb> / ((pi/p)^1.5 K)
__device__ double kin3(const int* la, const int* lb, double eb, const double* xpa, const double* xpb, double hp) {
  int l2[3] = {lb[0], lb[1], lb[2]};
  double r = eb * static_cast<double>(2 * (lb[0] + lb[1] + lb[2]) + 3) * ovl3(la, lb, xpa, xpb, hp);
  for (int d = 0; d < 3; ++d) {
    l2[d] = lb[d] + 2;
    r -= 2.0 * eb * eb * ovl3(la, l2, xpa, xpb, hp);
    if (lb[d] >= 2) {
      l2[d] = lb[d] - 2;
      r -= 0.5 * static_cast<double>(lb[d] * (lb[d] - 1)) * ovl3(la, l2, xpa, xpb, hp);
    }
    l2[d] = lb[d];
  }
  return r;
}

// sum_{tuv} E^{ab}_{tuv} R_{tuv}
__device__ double rinv3(const int* la, const int* lb, const double* xpa, const double* xpb, double hp, const double* R) {
  double Eb[3][8];
#pragma unroll
  for (int d = 0; d < 3; ++d) herm_E(la[d], lb[d], xpa[d], xpb[d], hp, Eb[d]);
  double acc = 0.0;
  for (int t = 0; t <= la[0] + lb[0]; ++t)
    for (int u = 0; u <= la[1] + lb[1]; ++u)
      for (int v = 0; v <= la[2] + lb[2]; ++v) acc = fma(Eb[0][t] * Eb[1][u] * Eb[2][v], R[ridx(t, u, v)], acc);
  return acc;
}

__global__ void __launch_bounds__(kGThreads)
gint1e_kernel(GView bs, const double* __restrict__ coords, GOut out) {
  extern __shared__ __align__(16) double sm[];
  double* boys = sm;
  double* Rc = boys + 2 * kGBoysN + 2;
  const int n = bs.nao, natm = bs.natm;
  const int g = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kGThreads / 32;
  for (int k = tid; k < 2 * kGBoysN; k += kGThreads) boys[k] = __ldg(bs.boys + k);
  for (int k = tid; k < 3 * natm; k += kGThreads) Rc[k] = coords[static_cast<int64_t>(g) * natm * 3 + k];
  __syncthreads();
  const int64_t n2 = static_cast<int64_t>(n) * n;
  for (int ab = blockIdx.x * NW + warp; ab < n * n; ab += gridDim.x * NW) {
    const int a = ab / n, b = ab - a * n;
    const int atA = bs.ao_atom[a], atB = bs.ao_atom[b];
    int la[3], lb[3];
    double A[3], B[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      la[d] = bs.ao_pow[3 * a + d]; lb[d] = bs.ao_pow[3 * b + d];
      A[d] = Rc[3 * atA + d]; B[d] = Rc[3 * atB + d];
    }
    const int pa0 = bs.ao_poff[a], na = bs.ao_poff[a + 1] - pa0, pb0 = bs.ao_poff[b], nb = bs.ao_poff[b + 1] - pb0;
    const double ab2 = (A[0] - B[0]) * (A[0] - B[0]) + (A[1] - B[1]) * (A[1] - B[1]) + (A[2] - B[2]) * (A[2] - B[2]);
    const int lab = la[0] + la[1] + la[2] + lb[0] + lb[1] + lb[2];
    double s = 0.0, tk = 0.0, ds[3] = {0, 0, 0}, dt[3] = {0, 0, 0};
    double rv[kGMaxAtoms], dr[kGMaxAtoms][3];
    for (int C = 0; C < natm; ++C) { rv[C] = 0.0; dr[C][0] = dr[C][1] = dr[C][2] = 0.0; }
    for (int t = lane; t < na * nb; t += 32) {
      const int i = t / nb, j = t - i * nb;
      const double ea = bs.prim_exp[pa0 + i], eb = bs.prim_exp[pb0 + j];
      const double p = ea + eb, hp = 0.5 / p;
      const double w = bs.prim_wt[pa0 + i] * bs.prim_wt[pb0 + j] * exp(-(ea * eb / p) * ab2);
      double P[3], xpa[3], xpb[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        P[d] = (ea * A[d] + eb * B[d]) / p;
        xpa[d] = P[d] - A[d]; xpb[d] = P[d] - B[d];
      }
      const double so = w * pow(3.14159265358979323846 / p, 1.5);
      s = fma(so, ovl3(la, lb, xpa, xpb, hp), s);
      tk = fma(so, kin3(la, lb, eb, xpa, xpb, hp), tk);
      int ls[3] = {la[0], la[1], la[2]};
      for (int d = 0; d < 3; ++d) {   // d/dA_d of the bra function
        ls[d] = la[d] + 1;
        double o = 2.0 * ea * ovl3(ls, lb, xpa, xpb, hp), k = 2.0 * ea * kin3(ls, lb, eb, xpa, xpb, hp);
        if (la[d] > 0) {
          ls[d] = la[d] - 1;
          o -= static_cast<double>(la[d]) * ovl3(ls, lb, xpa, xpb, hp);
          k -= static_cast<double>(la[d]) * kin3(ls, lb, eb, xpa, xpb, hp);
        }
        ls[d] = la[d];
        ds[d] = fma(so, o, ds[d]);
        dt[d] = fma(so, k, dt[d]);
      }
      const double vo = w * 6.28318530717958647692 / p;
      for (int C = 0; C < natm; ++C) {
        const double X = P[0] - Rc[3 * C], Y = P[1] - Rc[3 * C + 1], Z = P[2] - Rc[3 * C + 2];
        double F[kGMaxL + 2], R[84], S[84];
        const int L = lab + 1;
        boys_upto(L, p * (X * X + Y * Y + Z * Z), boys, F);
        build_R(L, p, X, Y, Z, F, R, S);
        rv[C] = fma(vo, rinv3(la, lb, xpa, xpb, hp, R), rv[C]);
        for (int d = 0; d < 3; ++d) {
          ls[d] = la[d] + 1;
          double r = 2.0 * ea * rinv3(ls, lb, xpa, xpb, hp, R);
          if (la[d] > 0) {
            ls[d] = la[d] - 1;
            r -= static_cast<double>(la[d]) * rinv3(ls, lb, xpa, xpb, hp, R);
          }
          ls[d] = la[d];
          dr[C][d] = fma(vo, r, dr[C][d]);
        }
      }
    }
    s = gwarp_sum(s); tk = gwarp_sum(tk);
    for (int d = 0; d < 3; ++d) { ds[d] = gwarp_sum(ds[d]); dt[d] = gwarp_sum(dt[d]); }
    double vsum = 0.0, nsum[3] = {0, 0, 0};
    for (int C = 0; C < natm; ++C) {
      rv[C] = gwarp_sum(rv[C]);
      const double z = bs.charges[C];
      vsum += z * rv[C];
      for (int d = 0; d < 3; ++d) {
        dr[C][d] = gwarp_sum(dr[C][d]);
        nsum[d] += z * dr[C][d];
      }
    }
    if (lane == 0) {
      // <nabla a|O|b> = -d/dA <a|O|b>
      out.ovlp[static_cast<int64_t>(g) * n2 + ab] = s;
      out.hcore[static_cast<int64_t>(g) * n2 + ab] = tk - vsum;
      for (int d = 0; d < 3; ++d) out.ipovlp[(static_cast<int64_t>(g) * 3 + d) * n2 + ab] = -ds[d];
      // v[C][x][a][b] = -Z_C iprinv^C[x][a][b] - [atom(a) == C] (ipkin + ipnuc)[x][a][b]
      //   iprinv^C = -dr[C],  ipkin = -dt,  ipnuc = -sum_C Z_C iprinv^C = +nsum
      for (int C = 0; C < natm; ++C)
        for (int d = 0; d < 3; ++d) {
          double v = bs.charges[C] * dr[C][d];
          if (C == atA) v -= -dt[d] + nsum[d];
          out.vtmp[((static_cast<int64_t>(g) * natm + C) * 3 + d) * n2 + ab] = v;
        }
    }
  }
  if (blockIdx.x == 0 && warp == 0) {  // nuclear repulsion
    double e = 0.0;
    for (int c0 = 0; c0 < natm; c0 += 32) {
      const int A = c0 + lane;
      double gx = 0, gy = 0, gz = 0;
      if (A < natm) {
        const double za = bs.charges[A];
        for (int B = 0; B < natm; ++B) {
          if (B == A) continue;
          const double dx = Rc[3 * A] - Rc[3 * B], dy = Rc[3 * A + 1] - Rc[3 * B + 1], dz = Rc[3 * A + 2] - Rc[3 * B + 2];
          const double r2 = dx * dx + dy * dy + dz * dz, ri = rsqrt(r2), zz = za * bs.charges[B];
          if (B < A) e += zz * ri;
          const double f = zz * ri * ri * ri;
          gx -= f * dx; gy -= f * dy; gz -= f * dz;
        }
        double* gn = out.grad_nuc + (static_cast<int64_t>(g) * natm + A) * 3;
        gn[0] = gx; gn[1] = gy; gn[2] = gz;
      }
    }
    e = gwarp_sum(e);
    if (lane == 0) out.e_nuc[g] = e;
  }
}

// hcore_deriv[g][C][x][a][b] = v[a][b] + v[b][a]
__global__ void ghd_sym_kernel(int64_t total, int n, const double* __restrict__ vtmp, double* __restrict__ hd) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= total) return;
  const int64_t n2 = static_cast<int64_t>(n) * n;
  const int64_t blk = k / n2, ab = k - blk * n2;
  const int a = static_cast<int>(ab / n), b = static_cast<int>(ab - static_cast<int64_t>(a) * n);
  hd[k] = vtmp[k] + vtmp[blk * n2 + static_cast<int64_t>(b) * n + a];
}

template <typename T>
int gupload(T** dst, const std::vector<T>& src) {
  EVC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), std::max<size_t>(1, src.size()) * sizeof(T)));
  if (!src.empty()) EVC_CHECK_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

extern "C" {

int evc_gbasis_create(evc_ctx* ctx, int natm, const double* charges_host, int nao, const int32_t* ao_atom_host,
                      const int32_t* ao_pow_host, const int32_t* ao_nprim_host, const double* prim_exp_host,
                      const double* prim_wt_host, evc_gbasis** out) {
  EVC_REQUIRE(ctx && charges_host && ao_atom_host && ao_pow_host && ao_nprim_host && prim_exp_host && prim_wt_host && out,
              "evc_gbasis_create: NULL argument");
  EVC_REQUIRE(natm >= 1 && natm <= kGMaxAtoms && nao >= 1 && nao <= 64, "evc_gbasis_create: natm=%d nao=%d unsupported",
              natm, nao);
  EVC_CHECK_CUDA(cudaSetDevice(ctx->device));
  std::vector<int32_t> ao_atom(ao_atom_host, ao_atom_host + nao), pw(ao_pow_host, ao_pow_host + 3 * nao), poff(nao + 1, 0),
      slices(2 * natm, 0);
  for (int a = 0; a < nao; ++a) {
    EVC_REQUIRE(ao_nprim_host[a] >= 1 && ao_nprim_host[a] <= 16, "evc_gbasis_create: AO %d has %d primitives", a,
                ao_nprim_host[a]);
    EVC_REQUIRE(ao_atom[a] >= 0 && ao_atom[a] < natm && (a == 0 || ao_atom[a] >= ao_atom[a - 1]),
                "evc_gbasis_create: AOs must be grouped by atom, in atom order");
    const int l = pw[3 * a] + pw[3 * a + 1] + pw[3 * a + 2];
    EVC_REQUIRE(pw[3 * a] >= 0 && pw[3 * a + 1] >= 0 && pw[3 * a + 2] >= 0 && l <= 1,
                "evc_gbasis_create: AO %d has angular momentum %d (s and p only)", a, l);
    poff[a + 1] = poff[a] + ao_nprim_host[a];
  }
  {
    int a = 0;
    for (int A = 0; A < natm; ++A) {
      slices[2 * A] = a;
      while (a < nao && ao_atom[a] == A) ++a;
      slices[2 * A + 1] = a;
    }
  }
  const int nprim = poff[nao];
  std::vector<double> ex(prim_exp_host, prim_exp_host + nprim), wt(prim_wt_host, prim_wt_host + nprim),
      ch(charges_host, charges_host + natm), boys(2 * static_cast<size_t>(kGBoysN));
  for (int i = 0; i < kGBoysN; ++i) {
    long double f[kGTop + 1];
    const long double t0 = static_cast<long double>(i) / kGPerUnit;
    boys_host_g(kGTop, t0, f);
    boys[2 * static_cast<size_t>(i)] = static_cast<double>(f[kGTop]);
    boys[2 * static_cast<size_t>(i) + 1] = static_cast<double>(expl(-t0));
  }
  {
    unsigned char tab[7][7][7];
    for (int t = 0; t < 7; ++t)
      for (int u = 0; u < 7; ++u)
        for (int v = 0; v < 7; ++v) {
          const int a = 7 - t;
          const int idx = 84 - a * (a + 1) * (a + 2) / 6 + u * (7 - t) - u * (u - 1) / 2 + v;
          tab[t][u][v] = static_cast<unsigned char>((t + u + v <= 6) ? idx : 0);
        }
    EVC_CHECK_CUDA(cudaMemcpyToSymbol(c_ridx, tab, sizeof(tab)));
  }
  // shells: an s AO, or the three consecutive components x, y, z of a p shell
  std::vector<int32_t> sh_atom, sh_ao0, sh_p0, sh_np, sh_l;
  for (int a = 0; a < nao;) {
    const int l = pw[3 * a] + pw[3 * a + 1] + pw[3 * a + 2];
    if (l == 1) {
      EVC_REQUIRE(a + 2 < nao && pw[3 * a] == 1 && pw[3 * (a + 1) + 1] == 1 && pw[3 * (a + 2) + 2] == 1 &&
                      ao_atom[a + 1] == ao_atom[a] && ao_atom[a + 2] == ao_atom[a] &&
                      ao_nprim_host[a + 1] == ao_nprim_host[a] && ao_nprim_host[a + 2] == ao_nprim_host[a],
                  "evc_gbasis_create: p functions must come as consecutive x, y, z components of one shell (AO %d)", a);
      for (int k = 0; k < ao_nprim_host[a]; ++k)
        EVC_REQUIRE(ex[poff[a] + k] == ex[poff[a + 1] + k] && ex[poff[a] + k] == ex[poff[a + 2] + k] &&
                        wt[poff[a] + k] == wt[poff[a + 1] + k] && wt[poff[a] + k] == wt[poff[a + 2] + k],
                    "evc_gbasis_create: the components of p shell at AO %d differ in their primitives", a);
    }
    sh_atom.push_back(ao_atom[a]); sh_ao0.push_back(a); sh_p0.push_back(poff[a]); sh_np.push_back(ao_nprim_host[a]);
    sh_l.push_back(l);
    a += l == 1 ? 3 : 1;
  }
  const int nshell = static_cast<int>(sh_atom.size());
  EVC_REQUIRE(nshell <= 255, "evc_gbasis_create: too many shells (%d)", nshell);
  // shell-quartet work lists per class: canonical order (p first inside bra and ket, heavier pair as bra),
  // sorted by lane-group size, packed 32 / gs quartets to a warp
  std::vector<int32_t> cq, cunits;
  int cq_off[kGClasses + 1] = {0}, cunit_off[kGClasses + 1] = {0};
  {
    std::vector<std::vector<std::pair<int, int32_t>>> per(kGClasses);
    for (int I = 0; I < nshell; ++I)
      for (int J = 0; J <= I; ++J)
        for (int K = 0; K <= I; ++K)
          for (int Lq = 0; Lq <= (K == I ? J : K); ++Lq) {
            int A = I, B = J, Cc = K, D = Lq;
            if (sh_l[A] < sh_l[B]) std::swap(A, B);
            if (sh_l[Cc] < sh_l[D]) std::swap(Cc, D);
            if (sh_l[Cc] + sh_l[D] > sh_l[A] + sh_l[B]) { std::swap(A, Cc); std::swap(B, D); }
            const int npf = sh_l[A] + sh_l[B] + sh_l[Cc] + sh_l[D];
            const int cls = npf <= 1 ? npf : npf == 2 ? (sh_l[B] ? 2 : 3) : npf + 1;
            const long long tot = static_cast<long long>(sh_np[A]) * sh_np[B] * sh_np[Cc] * sh_np[D];
            int lg = 0;
            while ((1 << lg) < 32 && (1LL << lg) < tot) ++lg;
            per[cls].emplace_back(lg, static_cast<int32_t>(A | (B << 8) | (Cc << 16) | (D << 24)));
          }
    for (int c = 0; c < kGClasses; ++c) {
      std::stable_sort(per[c].begin(), per[c].end(), [](const auto& x, const auto& y) { return x.first < y.first; });
      cq_off[c] = static_cast<int>(cq.size());
      cunit_off[c] = static_cast<int>(cunits.size() / 2);
      size_t k = 0;
      while (k < per[c].size()) {
        const int lg = per[c][k].first, perw = 32 >> lg;
        size_t e = k;
        while (e < per[c].size() && per[c][e].first == lg && e - k < static_cast<size_t>(perw)) ++e;
        cunits.push_back(static_cast<int32_t>(cq.size() + 0));
        cunits.push_back(static_cast<int32_t>(((e - k) << 8) | lg));
        for (size_t x = k; x < e; ++x) cq.push_back(per[c][x].second);
        k = e;
      }
    }
    cq_off[kGClasses] = static_cast<int>(cq.size());
    cunit_off[kGClasses] = static_cast<int>(cunits.size() / 2);
  }
  evc_gbasis* b = new evc_gbasis();
  b->natm = natm; b->nao = nao; b->nprim = nprim; b->nshell = nshell;
  for (int c = 0; c <= kGClasses; ++c) { b->cq_off[c] = cq_off[c]; b->cunit_off[c] = cunit_off[c]; }
  int rc = 0;
  if ((rc = gupload(&b->ao_atom, ao_atom)) || (rc = gupload(&b->ao_pow, pw)) || (rc = gupload(&b->ao_poff, poff)) ||
      (rc = gupload(&b->aoslices, slices)) || (rc = gupload(&b->prim_exp, ex)) || (rc = gupload(&b->prim_wt, wt)) ||
      (rc = gupload(&b->charges, ch)) || (rc = gupload(&b->boys, boys)) ||
      (rc = gupload(&b->sh_atom, sh_atom)) || (rc = gupload(&b->sh_ao0, sh_ao0)) ||
      (rc = gupload(&b->sh_p0, sh_p0)) || (rc = gupload(&b->sh_np, sh_np)) || (rc = gupload(&b->cq, cq)) ||
      (rc = gupload(&b->cunits, cunits))) {
    delete b;
    return rc;
  }
  *out = b;
  return 0;
}

int evc_gbasis_destroy(evc_gbasis* b) {
  if (b) {
    cudaFree(b->ao_atom); cudaFree(b->ao_pow); cudaFree(b->ao_poff); cudaFree(b->aoslices); cudaFree(b->prim_exp);
    cudaFree(b->prim_wt); cudaFree(b->charges); cudaFree(b->boys);
    cudaFree(b->sh_atom); cudaFree(b->sh_ao0); cudaFree(b->sh_p0); cudaFree(b->sh_np); cudaFree(b->cq); cudaFree(b->cunits);
  }
  delete b;
  return 0;
}

int evc_ao_integrals_sp_workspace_bytes(const evc_gbasis* b, int nbatch, size_t* bytes) {
  EVC_REQUIRE(b && bytes && nbatch >= 0, "evc_ao_integrals_sp_workspace_bytes: bad arguments");
  *bytes = evc_align_up(static_cast<size_t>(nbatch) * b->natm * 3 * b->nao * b->nao * sizeof(double), 256) + 256;
  return 0;
}

int evc_ao_integrals_sp(evc_ctx* ctx, const evc_gbasis* b, int nbatch, const double* coords, double* ovlp,
                        double* hcore, double* eri, double* ipovlp, double* hcore_deriv, double* eri_ip1,
                        double* e_nuc, double* grad_nuc, void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && b && coords && ovlp && hcore && eri && ipovlp && hcore_deriv && eri_ip1 && e_nuc && grad_nuc &&
                  workspace,
              "evc_ao_integrals_sp: NULL argument");
  if (nbatch <= 0) return 0;
  size_t need = 0;
  evc_ao_integrals_sp_workspace_bytes(b, nbatch, &need);
  EVC_REQUIRE(workspace_bytes >= need, "evc_ao_integrals_sp: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  GView v{b->natm, b->nao, b->ao_atom, b->ao_pow, b->ao_poff,
          b->sh_atom, b->sh_ao0, b->sh_p0, b->sh_np, b->prim_exp, b->prim_wt, b->charges, b->boys};
  GOut o{ovlp, hcore, eri, ipovlp, static_cast<double*>(workspace), eri_ip1, e_nuc, grad_nuc};
  const size_t smem = (2 * static_cast<size_t>(kGBoysN) + 2 + 3 * kGMaxAtoms) * sizeof(double);
  EVC_CHECK_CUDA(cudaFuncSetAttribute(gint1e_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int nw = kGThreads / 32;
  {
    int rc;
    if ((rc = launch_gclass_part0(ctx->stream, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_gclass_part1(ctx->stream, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_gclass_part2(ctx->stream, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_gclass_part3(ctx->stream, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_gclass_part4(ctx->stream, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
  }
  int split1 = 1;
  while (static_cast<long long>(nbatch) * split1 < 2LL * ctx->sm_count && split1 * nw < b->nao * b->nao) split1 *= 2;
  gint1e_kernel<<<dim3(split1, nbatch), kGThreads, smem, ctx->stream>>>(v, coords, o);
  EVC_CHECK_LAUNCH();
  const int64_t total = static_cast<int64_t>(nbatch) * b->natm * 3 * b->nao * b->nao;
  ghd_sym_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctx->stream>>>(total, b->nao, o.vtmp, hcore_deriv);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
