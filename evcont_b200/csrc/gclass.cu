// K9g two-electron class kernels (see integrals_sp.cuh / integrals_sp.cu).  Compiled once per
// EVC_GCLASS_PART (Makefile): 0 = ssss, psss, ppss, psps; 1 = ppps; 2, 3, 4 = pppp with JB = 0, 1, 2;
// 5 = the one-electron class kernels.
#include "integrals_sp.cuh"

#include <algorithm>

#ifndef EVC_GCLASS_PART
#error "compile with -DEVC_GCLASS_PART=0..5"
#endif

namespace evc_gint {
namespace {

// =====================================================================================================
// Class kernels: contracted SHELL quartets, everything in registers.
//
// The first version of K9g evaluated one contracted AO quartet at a time and kept its Hermite tables in
// per-thread local arrays (9 GB of local-memory DRAM traffic per Zundel launch, FP64 pipe 4 % active:
// profiles/r01c_gint_sp_ncu_full.txt).
// Here the work unit is a shell quartet in canonical class order (p shells first inside bra and ket, the
// pair with more p shells as the bra): ssss, psss, ppss, psps, ppps, pppp.  Lanes of a group run over the
// primitive quartets as before, but per primitive quartet the Boys values, the R table and all Hermite
// coefficients are compile-time indexed (fully unrolled) and live in registers, and every Cartesian
// component that shares them is evaluated from the same table:
//   * the component of the first p function is fixed by ROTATING the axes cyclically (pass r: x' = axis r),
//     so one code path serves the three components;
//   * the components of the last p function are enumerated inside the pass (3 x 10 accumulators: value and
//     d/dA, d/dB, d/dC);
//   * the components of the p functions in between are template parameters of the kernel: 3 (ppps) or 9
//     (pppp) launches over the same unit list (and separate compilation parts: the fully unrolled pppp
//     code takes minutes to compile).
// =====================================================================================================

template <int L>
__device__ __forceinline__ void boys_fixed(double T, const double* __restrict__ tab, double (&F)[L + 1]) {
  static_assert(L + 5 <= kGTop, "Boys table order too low");
  if (T < static_cast<double>(kGTmax)) {
    const double r = fma(T, static_cast<double>(kGPerUnit), 6755399441055744.0);
    const int i = __double2loint(r);
    const double t0 = (r - 6755399441055744.0) * (1.0 / kGPerUnit);
    const double d = t0 - T, tt = t0 + t0;
    const double2 fe = __ldg(reinterpret_cast<const double2*>(tab) + i);
    double g[kGTop + 1];
    g[kGTop] = fe.x;
#pragma unroll
    for (int m = kGTop; m > L; --m) g[m - 1] = fma(tt, g[m], fe.y) * (1.0 / static_cast<double>(2 * m - 1));
    double fl = g[L + 5];
    fl = fma(fl, d * 0.2, g[L + 4]);
    fl = fma(fl, d * 0.25, g[L + 3]);
    fl = fma(fl, d * (1.0 / 3.0), g[L + 2]);
    fl = fma(fl, d * 0.5, g[L + 1]);
    fl = fma(fl, d, g[L]);
    const double ed = 1.0 + d * (1.0 + d * (0.5 + d * (1.0 / 6.0 + d * (1.0 / 24.0 + d * (1.0 / 120.0 + d * (1.0 / 720.0))))));
    const double et = fe.y * ed, t2 = T + T;
    F[L] = fl;
#pragma unroll
    for (int m = L; m > 0; --m) F[m - 1] = fma(t2, F[m], et) * (1.0 / static_cast<double>(2 * m - 1));
  } else {
    const double ri = 1.0 / T, et = exp(-T);
    F[0] = 0.88622692545275801365 * sqrt(ri);
#pragma unroll
    for (int m = 0; m < L; ++m) F[m + 1] = (static_cast<double>(2 * m + 1) * F[m] - et) * (0.5 * ri);
  }
}

// R^0_{tuv}, t + u + v <= L, in R[t][u][v]; S is the ping-pong partner (both become registers)
template <int L>
__device__ __forceinline__ void build_R_fixed(double alpha, double X, double Y, double Z, const double (&F)[L + 1],
                                              double (&R)[L + 1][L + 1][L + 1]) {
  double S[L + 1][L + 1][L + 1];
  double pw[L + 1];
  pw[0] = 1.0;
#pragma unroll
  for (int n = 1; n <= L; ++n) pw[n] = pw[n - 1] * (-2.0 * alpha);
#pragma unroll
  for (int n = L; n >= 0; --n) {
    const int ord = L - n;
    double (&dst)[L + 1][L + 1][L + 1] = (n & 1) ? S : R;
    double (&src)[L + 1][L + 1][L + 1] = (n & 1) ? R : S;
    dst[0][0][0] = pw[n] * F[n];
#pragma unroll
    for (int t = 0; t <= L; ++t)
#pragma unroll
      for (int u = 0; u <= L; ++u)
#pragma unroll
        for (int v = 0; v <= L; ++v) {
          if (t + u + v > ord || t + u + v == 0) continue;
          if (t > 0) dst[t][u][v] = t > 1 ? fma(static_cast<double>(t - 1), src[t - 2][u][v], X * src[t - 1][u][v]) : X * src[t - 1][u][v];
          else if (u > 0) dst[t][u][v] = u > 1 ? fma(static_cast<double>(u - 1), src[t][u - 2][v], Y * src[t][u - 1][v]) : Y * src[t][u - 1][v];
          else dst[t][u][v] = v > 1 ? fma(static_cast<double>(v - 1), src[t][u][v - 2], Z * src[t][u][v - 1]) : Z * src[t][u][v - 1];
        }
  }
}

// E^{ij}_t, t = 0..i+j, for compile-time (after inlining) i, j <= 2, i + j <= 3; no arithmetic on known zeros
__device__ __forceinline__ void herm_fixed(int i, int j, double xa, double xb, double h, double (&E)[4]) {
  E[0] = 1.0;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s < i + j) {
      const double x = s < i ? xa : xb;
      double nw[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t <= s + 1) {
          double v = (t <= s) ? x * E[t] : h * E[t - 1];
          if (t >= 1 && t <= s) v = fma(h, E[t - 1], v);
          if (t + 1 <= s) v = fma(static_cast<double>(t + 1), E[t + 1], v);
          nw[t] = v;
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t <= s + 1) E[t] = nw[t];
    }
  }
}

struct PrimQ {
  double xpa[3], xpb[3], xqc[3], xqd[3];
  double hp, hq, ea2, eb2, ec2, pref;
};

// value and d/dA, d/dB, d/dC (rotated axes) of one primitive quartet with compile-time Cartesian powers:
// acc[0] += pref (ab|cd), acc[1 + d] += pref d/dA_d, acc[4 + d] += pref d/dB_d, acc[7 + d] += pref d/dC_d
template <int AX, int AY, int AZ, int BX, int BY, int BZ, int CX, int CY, int CZ, int DX, int DY, int DZ, int L>
__device__ __forceinline__ void quartet_block(const PrimQ& q, const double (&R)[L + 1][L + 1][L + 1], double (&acc)[10]) {
  constexpr int a[3] = {AX, AY, AZ}, b[3] = {BX, BY, BZ}, c[3] = {CX, CY, CZ}, dd[3] = {DX, DY, DZ};
  constexpr int nb[3] = {AX + BX, AY + BY, AZ + BZ}, nk[3] = {CX + DX, CY + DY, CZ + DZ};
  double E0[3][4], EAu[3][4], EAd[3][4], EBu[3][4], EBd[3][4], K0[3][4], KCu[3][4], KCd[3][4];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    herm_fixed(a[d], b[d], q.xpa[d], q.xpb[d], q.hp, E0[d]);
    herm_fixed(a[d] + 1, b[d], q.xpa[d], q.xpb[d], q.hp, EAu[d]);
    if (a[d] > 0) herm_fixed(a[d] - 1, b[d], q.xpa[d], q.xpb[d], q.hp, EAd[d]);
    herm_fixed(a[d], b[d] + 1, q.xpa[d], q.xpb[d], q.hp, EBu[d]);
    if (b[d] > 0) herm_fixed(a[d], b[d] - 1, q.xpa[d], q.xpb[d], q.hp, EBd[d]);
    herm_fixed(c[d], dd[d], q.xqc[d], q.xqd[d], q.hq, K0[d]);
    herm_fixed(c[d] + 1, dd[d], q.xqc[d], q.xqd[d], q.hq, KCu[d]);
    if (c[d] > 0) herm_fixed(c[d] - 1, dd[d], q.xqc[d], q.xqd[d], q.hq, KCd[d]);
  }
  // ket contraction: G(kx, ky, kz tables; ranges) at bra index (t, u, v)
  auto ketG = [&](const double (&kx)[4], const double (&ky)[4], const double (&kz)[4], int rx, int ry, int rz, int t,
                  int u, int v) {
    double g = 0.0;
    bool first = true;
#pragma unroll
    for (int t2 = 0; t2 < 4; ++t2)
#pragma unroll
      for (int u2 = 0; u2 < 4; ++u2)
#pragma unroll
        for (int v2 = 0; v2 < 4; ++v2) {
          if (t2 > rx || u2 > ry || v2 > rz) continue;
          const double e = kx[t2] * ky[u2] * kz[v2];
          const double term = ((t2 + u2 + v2) & 1) ? -e : e;
          if (first) { g = term * R[t + t2][u + u2][v + v2]; first = false; }
          else g = fma(term, R[t + t2][u + u2][v + v2], g);
        }
    return g;
  };
  // base ket against the bra range widened by one in every single direction
  double G0[4][4][4];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (t > nb[0] + 1 || u > nb[1] + 1 || v > nb[2] + 1) continue;
        if ((t > nb[0]) + (u > nb[1]) + (v > nb[2]) > 1) continue;
        G0[t][u][v] = ketG(K0[0], K0[1], K0[2], nk[0], nk[1], nk[2], t, u, v);
      }
  auto braDot = [&](const double (&ex)[4], const double (&ey)[4], const double (&ez)[4], int rx, int ry, int rz,
                    const double (&G)[4][4][4]) {
    double s = 0.0;
    bool first = true;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (t > rx || u > ry || v > rz) continue;
          const double e = ex[t] * ey[u] * ez[v];
          if (first) { s = e * G[t][u][v]; first = false; }
          else s = fma(e, G[t][u][v], s);
        }
    return s;
  };
  acc[0] = fma(q.pref, braDot(E0[0], E0[1], E0[2], nb[0], nb[1], nb[2], G0), acc[0]);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int ux = d == 0, uy = d == 1, uz = d == 2;
    double sa = q.ea2 * braDot(d == 0 ? EAu[0] : E0[0], d == 1 ? EAu[1] : E0[1], d == 2 ? EAu[2] : E0[2], nb[0] + ux,
                               nb[1] + uy, nb[2] + uz, G0);
    if (a[d] > 0)
      sa = fma(-static_cast<double>(a[d]),
               braDot(d == 0 ? EAd[0] : E0[0], d == 1 ? EAd[1] : E0[1], d == 2 ? EAd[2] : E0[2], nb[0] - ux, nb[1] - uy,
                      nb[2] - uz, G0), sa);
    double sb = q.eb2 * braDot(d == 0 ? EBu[0] : E0[0], d == 1 ? EBu[1] : E0[1], d == 2 ? EBu[2] : E0[2], nb[0] + ux,
                               nb[1] + uy, nb[2] + uz, G0);
    if (b[d] > 0)
      sb = fma(-static_cast<double>(b[d]),
               braDot(d == 0 ? EBd[0] : E0[0], d == 1 ? EBd[1] : E0[1], d == 2 ? EBd[2] : E0[2], nb[0] - ux, nb[1] - uy,
                      nb[2] - uz, G0), sb);
    acc[1 + d] = fma(q.pref, sa, acc[1 + d]);
    acc[4 + d] = fma(q.pref, sb, acc[4 + d]);
    // d/dC_d: shifted ket tables against the base bra
    double G1[4][4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (t > nb[0] || u > nb[1] || v > nb[2]) continue;
          G1[t][u][v] = ketG(d == 0 ? KCu[0] : K0[0], d == 1 ? KCu[1] : K0[1], d == 2 ? KCu[2] : K0[2], nk[0] + ux,
                             nk[1] + uy, nk[2] + uz, t, u, v);
        }
    double sc = q.ec2 * braDot(E0[0], E0[1], E0[2], nb[0], nb[1], nb[2], G1);
    if (c[d] > 0) {
      double G2[4][4][4];
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (t > nb[0] || u > nb[1] || v > nb[2]) continue;
            G2[t][u][v] = ketG(d == 0 ? KCd[0] : K0[0], d == 1 ? KCd[1] : K0[1], d == 2 ? KCd[2] : K0[2], nk[0] - ux,
                               nk[1] - uy, nk[2] - uz, t, u, v);
          }
      sc = fma(-static_cast<double>(c[d]), braDot(E0[0], E0[1], E0[2], nb[0], nb[1], nb[2], G2), sc);
    }
    acc[7 + d] = fma(q.pref, sc, acc[7 + d]);
  }
}

struct ShellQ {
  int atom[4], ao0[4], p0[4], np[4];
};

// component of function f (1 = b, 2 = c, 3 = d) in block JIN of pass (JB, JC)
template <int LA, int LB, int LC, int LD>
struct ClassInfo {
  static constexpr int L = LA + LB + LC + LD + 1;
  static constexpr bool DIN = LD != 0, CIN = (LC != 0) && !DIN, BIN = (LB != 0) && !CIN && !DIN;
  static constexpr int NIN = (DIN || CIN || BIN) ? 3 : 1;
};

template <int LA, int LB, int LC, int LD, int JB, int JC, int JIN>
__device__ __forceinline__ void class_block(const PrimQ& q, const double (&R)[LA + LB + LC + LD + 2][LA + LB + LC + LD + 2][LA + LB + LC + LD + 2],
                                            double (&acc)[10]) {
  using CI = ClassInfo<LA, LB, LC, LD>;
  constexpr int cb = CI::BIN ? JIN : JB, cc = CI::CIN ? JIN : JC, cd = JIN;
  quartet_block<LA, 0, 0,
                (LB && cb == 0), (LB && cb == 1), (LB && cb == 2),
                (LC && cc == 0), (LC && cc == 1), (LC && cc == 2),
                (LD && cd == 0), (LD && cd == 1), (LD && cd == 2), CI::L>(q, R, acc);
}

// all primitive quartets of one shell quartet dealt to the lanes of the group, one (rotation, JB, JC) pass
template <int LA, int LB, int LC, int LD, int JB, int JC>
__device__ __forceinline__ void class_prim_loop(const GView& bs, const ShellQ& sq, const double (&ctr)[4][3], int lig, int gs,
                                                int tot, double (&acc)[ClassInfo<LA, LB, LC, LD>::NIN][10]) {
  using CI = ClassInfo<LA, LB, LC, LD>;
  constexpr int L = CI::L;
  const double ab2 = (ctr[0][0] - ctr[1][0]) * (ctr[0][0] - ctr[1][0]) + (ctr[0][1] - ctr[1][1]) * (ctr[0][1] - ctr[1][1]) +
                     (ctr[0][2] - ctr[1][2]) * (ctr[0][2] - ctr[1][2]);
  const double cd2 = (ctr[2][0] - ctr[3][0]) * (ctr[2][0] - ctr[3][0]) + (ctr[2][1] - ctr[3][1]) * (ctr[2][1] - ctr[3][1]) +
                     (ctr[2][2] - ctr[3][2]) * (ctr[2][2] - ctr[3][2]);
  // mixed-radix decode of the primitive index without integer division: (r + 1/2) / np in single precision is
  // exact for r < 2^16 (at most 16 primitives per shell)
  const float inv3 = 1.0f / static_cast<float>(sq.np[3]), inv2 = 1.0f / static_cast<float>(sq.np[2]),
              inv1 = 1.0f / static_cast<float>(sq.np[1]);
#pragma unroll 1
  for (int t = lig; t < tot; t += gs) {
    int r = t;
    int qd = __float2int_rz((static_cast<float>(r) + 0.5f) * inv3);
    const int il = r - qd * sq.np[3]; r = qd;
    qd = __float2int_rz((static_cast<float>(r) + 0.5f) * inv2);
    const int ik = r - qd * sq.np[2]; r = qd;
    qd = __float2int_rz((static_cast<float>(r) + 0.5f) * inv1);
    const int ij = r - qd * sq.np[1];
    const int ii = qd;
    const double ea = __ldg(bs.prim_exp + sq.p0[0] + ii), eb = __ldg(bs.prim_exp + sq.p0[1] + ij),
                 ec = __ldg(bs.prim_exp + sq.p0[2] + ik), ed = __ldg(bs.prim_exp + sq.p0[3] + il);
    const double w4 = __ldg(bs.prim_wt + sq.p0[0] + ii) * __ldg(bs.prim_wt + sq.p0[1] + ij) *
                      __ldg(bs.prim_wt + sq.p0[2] + ik) * __ldg(bs.prim_wt + sq.p0[3] + il);
    const double p = ea + eb, qq = ec + ed, ip = 1.0 / p, iq = 1.0 / qq;
    const double kk = exp(-(ea * eb * ip) * ab2 - (ec * ed * iq) * cd2);
    PrimQ q;
    q.pref = w4 * kk * 34.986836655249725 * ip * iq * rsqrt(p + qq);   // 2 pi^2.5 / (p q sqrt(p + q))
    if (fabs(q.pref) < 1.0e-18) continue;
    double pq[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double P = (ea * ctr[0][d] + eb * ctr[1][d]) * ip, Q = (ec * ctr[2][d] + ed * ctr[3][d]) * iq;
      q.xpa[d] = P - ctr[0][d]; q.xpb[d] = P - ctr[1][d];
      q.xqc[d] = Q - ctr[2][d]; q.xqd[d] = Q - ctr[3][d];
      pq[d] = P - Q;
    }
    q.hp = 0.5 * ip; q.hq = 0.5 * iq;
    q.ea2 = 2.0 * ea; q.eb2 = 2.0 * eb; q.ec2 = 2.0 * ec;
    const double rho = p * qq / (p + qq);
    double F[L + 1], R[L + 1][L + 1][L + 1];
    boys_fixed<L>(rho * (pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2]), bs.boys, F);
    build_R_fixed<L>(rho, pq[0], pq[1], pq[2], F, R);
    class_block<LA, LB, LC, LD, JB, JC, 0>(q, R, acc[0]);
    if constexpr (CI::NIN == 3) {
      class_block<LA, LB, LC, LD, JB, JC, 1>(q, R, acc[1]);
      class_block<LA, LB, LC, LD, JB, JC, 2>(q, R, acc[2]);
    }
  }
}

// 128-thread CTAs, EVC_GCLASS_MINB CTAs per SM.  At 2 the p classes take ~255 registers without spilling
// (8 warps per SM; ncu: "wait" is the top stall, the FP64 pipe waits on its own dependent latency); at 3
// (168 registers, 12 warps) they spill 200-900 bytes per thread and run at the same speed (measured).
constexpr int kCThreads = 128;
#ifndef EVC_GCLASS_MINB
#define EVC_GCLASS_MINB 2
#endif

template <int LA, int LB, int LC, int LD, int JB, int JC>
__global__ void __launch_bounds__(kCThreads, EVC_GCLASS_MINB)
gint2e_class_kernel(GView bs, const int32_t* __restrict__ qlist, const int32_t* __restrict__ units, int nunits,
                    const double* __restrict__ coords, GOut out) {
  using CI = ClassInfo<LA, LB, LC, LD>;
  const int n = bs.nao, natm = bs.natm;
  const int g = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kCThreads / 32;
  const double* Rc = coords + static_cast<int64_t>(g) * natm * 3;
  const int64_t n2 = static_cast<int64_t>(n) * n, n3 = n2 * n, n4 = n2 * n2;
  double* eri = out.eri + static_cast<int64_t>(g) * n4;
  double* ip1 = out.eri_ip1 + static_cast<int64_t>(g) * 3 * n4;
  for (int un = blockIdx.x * NW + warp; un < nunits; un += gridDim.x * NW) {
    const int ustart = units[2 * un], uinfo = units[2 * un + 1];
    const int gs = 1 << (uinfo & 0xff), ucount = uinfo >> 8;
    const int grp = lane / gs, lig = lane - grp * gs;
    const bool live = grp < ucount;
    const int code = qlist[ustart + (live ? grp : 0)];
    ShellQ sq;
    int sh[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      sh[f] = (code >> (8 * f)) & 0xff;
      sq.atom[f] = bs.sh_atom[sh[f]];
      sq.ao0[f] = bs.sh_ao0[sh[f]];
      sq.p0[f] = bs.sh_p0[sh[f]];
      sq.np[f] = bs.sh_np[sh[f]];
    }
    const int tot = live ? sq.np[0] * sq.np[1] * sq.np[2] * sq.np[3] : 0;
    const bool same_bra = sh[0] == sh[1], same_ket = sh[2] == sh[3];
#pragma unroll 1
    for (int r = 0; r < (LA ? 3 : 1); ++r) {
      double ctr[4][3];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          int od = d + r;
          if (od >= 3) od -= 3;
          ctr[f][d] = Rc[3 * sq.atom[f] + od];
        }
      {
        constexpr int jb = JB, jc = JC;
        double acc[CI::NIN][10];
#pragma unroll
        for (int k = 0; k < CI::NIN; ++k)
#pragma unroll
          for (int c = 0; c < 10; ++c) acc[k][c] = 0.0;
        class_prim_loop<LA, LB, LC, LD, JB, JC>(bs, sq, ctr, lig, gs, tot, acc);
        // butterfly inside the lane group (fixed order)
        for (int o = gs >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int k = 0; k < CI::NIN; ++k)
#pragma unroll
            for (int c = 0; c < 10; ++c) acc[k][c] += __shfl_xor_sync(0xffffffffu, acc[k][c], o);
        }
        if (!live) continue;
        // rotated component x' is axis r, y' axis r + 1, z' axis r + 2 (mod 3)
        auto orig = [&](int comp) { const int o = comp + r; return o >= 3 ? o - 3 : o; };
#pragma unroll
        for (int k = 0; k < CI::NIN; ++k) {
          const int cb = CI::BIN ? k : jb, cc = CI::CIN ? k : jc, cdd = k;
          const int a = sq.ao0[0] + (LA ? orig(0) : 0), b = sq.ao0[1] + (LB ? orig(cb) : 0),
                    c = sq.ao0[2] + (LC ? orig(cc) : 0), d = sq.ao0[3] + (LD ? orig(cdd) : 0);
          // each AO quartet once: same-shell pairs and the diagonal shell quartet hold duplicates
          if ((same_bra && a < b) || (same_ket && c < d)) continue;
          const int I = a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a;
          const int K = c >= d ? c * (c + 1) / 2 + d : d * (d + 1) / 2 + c;
          if (sh[0] == sh[2] && sh[1] == sh[3] && I < K) continue;
          for (int pr = lig; pr < 8; pr += gs) {
            const int who = pr >> 1;
            int i0, i1, i2, i3;
            if (who == 0) { i0 = a; i1 = b; i2 = (pr & 1) ? d : c; i3 = (pr & 1) ? c : d; }
            else if (who == 1) { i0 = b; i1 = a; i2 = (pr & 1) ? d : c; i3 = (pr & 1) ? c : d; }
            else if (who == 2) { i0 = c; i1 = d; i2 = (pr & 1) ? b : a; i3 = (pr & 1) ? a : b; }
            else { i0 = d; i1 = c; i2 = (pr & 1) ? b : a; i3 = (pr & 1) ? a : b; }
            int eff = who;  // coinciding index permutations carry bit-identical values
            if (eff == 1 && a == b) eff = 0;
            if (eff == 3 && c == d) eff = 2;
            if (I == K) eff -= (eff >= 2) ? 2 : 0;
            double g3[3];
#pragma unroll
            for (int dxyz = 0; dxyz < 3; ++dxyz) {
              const double da = acc[k][1 + dxyz], db = acc[k][4 + dxyz], dc = acc[k][7 + dxyz];
              g3[dxyz] = eff == 0 ? da : eff == 1 ? db : eff == 2 ? dc : -(da + db + dc);
            }
            const int64_t idx = i0 * n3 + i1 * n2 + i2 * n + i3;
            eri[idx] = acc[k][0];
            ip1[orig(0) * n4 + idx] = -g3[0];
            ip1[orig(1) * n4 + idx] = -g3[1];
            ip1[orig(2) * n4 + idx] = -g3[2];
          }
        }
      }
    }
  }
}


template <int LA, int LB, int LC, int LD, int JB, int JC>
int launch_one(cudaStream_t st, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits, int nun,
               const double* coords, const GOut& o) {
  if (nun <= 0) return 0;
  constexpr int nw = kCThreads / 32;
  // CTAs sized so that the grid covers the device a few times over
  const long long want = (nun + nw - 1) / nw, cap = std::max(1LL, 16LL * sm_count / nbatch);
  const int split = static_cast<int>(std::max(1LL, std::min(want, cap)));
  gint2e_class_kernel<LA, LB, LC, LD, JB, JC><<<dim3(split, nbatch), kCThreads, 0, st>>>(v, cq, cunits, nun, coords, o);
  EVC_CHECK_LAUNCH();
  return 0;
}

#if EVC_GCLASS_PART == 5
// ---------------------------------------------------------------------------------------------------------
// One-electron class kernels: one warp per ORDERED shell pair (A, B) of class (LA, LB) (the derivative acts on
// the first function only, evcont/ab_initio_gradients_loewdin.py:25, 147: <nabla a|O|b>).  The lanes form
// 32 / gs groups of gs = 2^k >= #primitive pairs lanes; a group takes one "slot" at a time -- nucleus C
// (nuclear attraction and its derivative, needed per nucleus) or the extra slot natm (overlap and kinetic
// energy) -- with its lanes over the primitive pairs; Boys values, R table and Hermite coefficients are
// compile-time indexed registers as in the two-electron kernels.  Group results go through shared memory,
// then the lanes assemble ovlp, hcore, int1e_ipovlp and the per-nucleus rows the caller symmetrises into
// hcore_generator's arrays.
// ---------------------------------------------------------------------------------------------------------
constexpr int k1eThreads = 128;

// E^{ij}_0 for i <= 2, j <= 3 (1-D overlap factor without sqrt(pi / p))
__device__ __forceinline__ double ovl1d(int i, int j, double xa, double xb, double h) {
  double E[6];
  E[0] = 1.0;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    if (s < i + j) {
      const double x = s < i ? xa : xb;
      double nw[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        if (t <= s + 1) {
          double v = (t <= s) ? x * E[t] : h * E[t - 1];
          if (t >= 1 && t <= s) v = fma(h, E[t - 1], v);
          if (t + 1 <= s) v = fma(static_cast<double>(t + 1), E[t + 1], v);
          nw[t] = v;
        }
      }
#pragma unroll
      for (int t = 0; t < 6; ++t)
        if (t <= s + 1) E[t] = nw[t];
    }
  }
  return E[0];
}

// per primitive pair: 1-D overlap factors O[d][i][j] = E^{ij}_0 (i <= LA + 1, j <= LB + 2) and 1-D Hermite
// tables EH[d][i][j][t] (i <= LA + 1, j <= LB), built once and indexed at compile time by every component
struct Prim1 {
  double O[3][3][4];
  double EH[3][3][2][4];
  double ea2, eb;
};

__device__ __forceinline__ double ovl3c(int ax, int ay, int az, int bx, int by, int bz, const Prim1& q) {
  return q.O[0][ax][bx] * q.O[1][ay][by] * q.O[2][az][bz];
}

// <a| -1/2 nabla^2 |b> / ((pi/p)^1.5 K) for b in {s, p}
__device__ __forceinline__ double kin3c(int ax, int ay, int az, int bx, int by, int bz, const Prim1& q) {
  double r = q.eb * static_cast<double>(2 * (bx + by + bz) + 3) * ovl3c(ax, ay, az, bx, by, bz, q);
  const double m = -2.0 * q.eb * q.eb;
  r = fma(m, ovl3c(ax, ay, az, bx + 2, by, bz, q), r);
  r = fma(m, ovl3c(ax, ay, az, bx, by + 2, bz, q), r);
  r = fma(m, ovl3c(ax, ay, az, bx, by, bz + 2, q), r);
  return r;
}

template <int L>
__device__ __forceinline__ double rinv3c(int ax, int ay, int az, int bx, int by, int bz, const Prim1& q,
                                         const double (&R)[L + 1][L + 1][L + 1]) {
  double s = 0.0;
  bool first = true;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (t > ax + bx || u > ay + by || v > az + bz) continue;
        const double e = q.EH[0][ax][bx][t] * q.EH[1][ay][by][u] * q.EH[2][az][bz][v];
        if (first) { s = e * R[t][u][v]; first = false; }
        else s = fma(e, R[t][u][v], s);
      }
  return s;
}

// acc[0] value, acc[1..3] d/dA of it (both for operator O given by the functor f(ax..bz))
template <int AX, int AY, int AZ, int BX, int BY, int BZ, typename F>
__device__ __forceinline__ void with_grad(double wgt, double ea2, double* acc, F f) {
  acc[0] = fma(wgt, f(AX, AY, AZ, BX, BY, BZ), acc[0]);
  double gx = ea2 * f(AX + 1, AY, AZ, BX, BY, BZ), gy = ea2 * f(AX, AY + 1, AZ, BX, BY, BZ),
         gz = ea2 * f(AX, AY, AZ + 1, BX, BY, BZ);
  if (AX > 0) gx = fma(-static_cast<double>(AX), f(AX - 1, AY, AZ, BX, BY, BZ), gx);
  if (AY > 0) gy = fma(-static_cast<double>(AY), f(AX, AY - 1, AZ, BX, BY, BZ), gy);
  if (AZ > 0) gz = fma(-static_cast<double>(AZ), f(AX, AY, AZ - 1, BX, BY, BZ), gz);
  acc[1] = fma(wgt, gx, acc[1]);
  acc[2] = fma(wgt, gy, acc[2]);
  acc[3] = fma(wgt, gz, acc[3]);
}

template <int LA, int LB, int IA, int IB, bool NUC, int L>
__device__ __forceinline__ void block1e(const Prim1& q, double so, double vo, const double (&R)[L + 1][L + 1][L + 1],
                                        double* acc) {
  constexpr int AX = LA && IA == 0, AY = LA && IA == 1, AZ = LA && IA == 2;
  constexpr int BX = LB && IB == 0, BY = LB && IB == 1, BZ = LB && IB == 2;
  if constexpr (NUC) {
    with_grad<AX, AY, AZ, BX, BY, BZ>(vo, q.ea2, acc, [&](int ax, int ay, int az, int bx, int by, int bz) {
      return rinv3c<L>(ax, ay, az, bx, by, bz, q, R);
    });
  } else {
    with_grad<AX, AY, AZ, BX, BY, BZ>(so, q.ea2, acc, [&](int ax, int ay, int az, int bx, int by, int bz) {
      return ovl3c(ax, ay, az, bx, by, bz, q);
    });
    with_grad<AX, AY, AZ, BX, BY, BZ>(so, q.ea2, acc + 4, [&](int ax, int ay, int az, int bx, int by, int bz) {
      return kin3c(ax, ay, az, bx, by, bz, q);
    });
  }
}

template <int LA, int LB, bool NUC>
__device__ __forceinline__ void slot1e(const GView& bs, const double (&A)[3], const double (&B)[3], const double* Cc, int p0a,
                                       int npa, int p0b, int npb, int lig, int gs,
                                       double (&acc)[(LA ? 3 : 1) * (LB ? 3 : 1)][8]) {
  constexpr int L = LA + LB + 1, NA = LA ? 3 : 1, NB = LB ? 3 : 1;
  const double ab2 = (A[0] - B[0]) * (A[0] - B[0]) + (A[1] - B[1]) * (A[1] - B[1]) + (A[2] - B[2]) * (A[2] - B[2]);
#pragma unroll 1
  for (int t = lig; t < npa * npb; t += gs) {
    const int i = t / npb, j = t - i * npb;
    const double ea = __ldg(bs.prim_exp + p0a + i), eb = __ldg(bs.prim_exp + p0b + j);
    const double p = ea + eb, ip = 1.0 / p, hp = 0.5 * ip;
    const double w = __ldg(bs.prim_wt + p0a + i) * __ldg(bs.prim_wt + p0b + j) * exp(-(ea * eb * ip) * ab2);
    Prim1 q;
    double P[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      P[d] = (ea * A[d] + eb * B[d]) * ip;
      const double xpa = P[d] - A[d], xpb = P[d] - B[d];
#pragma unroll
      for (int ii = 0; ii <= LA + 1; ++ii) {
        if constexpr (NUC) {
#pragma unroll
          for (int jj = 0; jj <= LB; ++jj) herm_fixed(ii, jj, xpa, xpb, hp, q.EH[d][ii][jj]);
        } else {
#pragma unroll
          for (int jj = 0; jj <= LB + 2; ++jj) q.O[d][ii][jj] = ovl1d(ii, jj, xpa, xpb, hp);
        }
      }
    }
    q.ea2 = 2.0 * ea; q.eb = eb;
    const double pip = 3.14159265358979323846 * ip;
    const double so = w * pip * sqrt(pip), vo = w * 6.28318530717958647692 * ip;
    double R[L + 1][L + 1][L + 1];
    if constexpr (NUC) {
      const double X = P[0] - Cc[0], Y = P[1] - Cc[1], Z = P[2] - Cc[2];
      double F[L + 1];
      boys_fixed<L>(p * (X * X + Y * Y + Z * Z), bs.boys, F);
      build_R_fixed<L>(p, X, Y, Z, F, R);
    }
#define EVC_B1(IA, IB) block1e<LA, LB, IA, IB, NUC, L>(q, so, vo, R, acc[IA * NB + IB]);
    EVC_B1(0, 0)
    if constexpr (NB == 3) { EVC_B1(0, 1) EVC_B1(0, 2) }
    if constexpr (NA == 3) {
      EVC_B1(1, 0) EVC_B1(2, 0)
      if constexpr (NB == 3) { EVC_B1(1, 1) EVC_B1(1, 2) EVC_B1(2, 1) EVC_B1(2, 2) }
    }
#undef EVC_B1
  }
}

template <int LA, int LB>
__global__ void __launch_bounds__(k1eThreads)
gint1e_class_kernel(GView bs, const int32_t* __restrict__ plist, int npairs, const double* __restrict__ coords, GOut out) {
  extern __shared__ double sm1e[];
  constexpr int NA = LA ? 3 : 1, NB = LB ? 3 : 1, NC = NA * NB, NW = k1eThreads / 32;
  const int n = bs.nao, natm = bs.natm, nslot = natm + 1;
  const int g = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* Rc = coords + static_cast<int64_t>(g) * natm * 3;
  double* buf = sm1e + static_cast<size_t>(warp) * (kGMaxAtoms + 1) * NC * 8;   // [slot][comp][8]
  const int64_t n2 = static_cast<int64_t>(n) * n;
  for (int pi = blockIdx.x * NW + warp; pi < npairs; pi += gridDim.x * NW) {
    const int code = plist[pi], shA = code & 0xff, shB = (code >> 8) & 0xff;
    const int atA = bs.sh_atom[shA], atB = bs.sh_atom[shB];
    const int p0a = bs.sh_p0[shA], npa = bs.sh_np[shA], p0b = bs.sh_p0[shB], npb = bs.sh_np[shB];
    const int a0 = bs.sh_ao0[shA], b0 = bs.sh_ao0[shB];
    double A[3], B[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { A[d] = Rc[3 * atA + d]; B[d] = Rc[3 * atB + d]; }
    int gs = 1;
    while (gs < 32 && gs < npa * npb) gs <<= 1;
    const int ngrp = 32 / gs, grp = lane / gs, lig = lane - grp * gs;
    for (int slot0 = 0; slot0 < nslot; slot0 += ngrp) {
      const int slot = slot0 + grp;
      double acc[NC][8];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[c][k] = 0.0;
      if (slot < natm) {
        const double Cc[3] = {Rc[3 * slot], Rc[3 * slot + 1], Rc[3 * slot + 2]};
        slot1e<LA, LB, true>(bs, A, B, Cc, p0a, npa, p0b, npb, lig, gs, acc);
      } else if (slot == natm) {
        slot1e<LA, LB, false>(bs, A, B, nullptr, p0a, npa, p0b, npb, lig, gs, acc);
      }
      for (int o = gs >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[c][k] += __shfl_xor_sync(0xffffffffu, acc[c][k], o);
      }
      if (lig == 0 && slot < nslot) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int k = 0; k < 8; ++k) buf[(slot * NC + c) * 8 + k] = acc[c][k];
      }
    }
    __syncwarp();
    // assemble: buf[C][c][0] = rinv value, [1..3] = d/dA; buf[natm][c][0] = S, [1..3] = dS/dA, [4] = T, [5..7] = dT/dA
    for (int c = lane; c < NC; c += 32) {
      const int ia = c / NB, ib = c - ia * NB;
      const int64_t ab = static_cast<int64_t>(a0 + ia) * n + (b0 + ib);
      const double* st = buf + (natm * NC + c) * 8;
      double vsum = 0.0;
      for (int C = 0; C < natm; ++C) vsum += bs.charges[C] * buf[(C * NC + c) * 8];
      out.ovlp[static_cast<int64_t>(g) * n2 + ab] = st[0];
      out.hcore[static_cast<int64_t>(g) * n2 + ab] = st[4] - vsum;
      for (int d = 0; d < 3; ++d) out.ipovlp[(static_cast<int64_t>(g) * 3 + d) * n2 + ab] = -st[1 + d];
    }
    // v[C][x][a][b] = -Z_C iprinv^C[x][a][b] - [atom(a) == C] (ipkin + ipnuc)[x][a][b],
    //   iprinv^C = -d/dA rinv^C,  ipkin = -dT/dA,  ipnuc = +sum_C Z_C d/dA rinv^C
    for (int it = lane; it < NC * natm * 3; it += 32) {
      const int c = it / (natm * 3), r = it - c * natm * 3, C = r / 3, d = r - 3 * C;
      const int ia = c / NB, ib = c - ia * NB;
      const int64_t ab = static_cast<int64_t>(a0 + ia) * n + (b0 + ib);
      double v = bs.charges[C] * buf[(C * NC + c) * 8 + 1 + d];
      if (C == atA) {
        double nsum = 0.0;
        for (int C2 = 0; C2 < natm; ++C2) nsum += bs.charges[C2] * buf[(C2 * NC + c) * 8 + 1 + d];
        v -= -buf[(natm * NC + c) * 8 + 5 + d] + nsum;
      }
      out.vtmp[((static_cast<int64_t>(g) * natm + C) * 3 + d) * n2 + ab] = v;
    }
    __syncwarp();
  }
}

template <int LA, int LB>
int launch_1e(cudaStream_t st, int sm_count, int nbatch, const GView& v, const int32_t* plist, int npairs, const double* coords,
              const GOut& o) {
  if (npairs <= 0) return 0;
  constexpr int nw = k1eThreads / 32, NC = (LA ? 3 : 1) * (LB ? 3 : 1);
  const size_t smem = static_cast<size_t>(nw) * (kGMaxAtoms + 1) * NC * 8 * sizeof(double);
  auto kern = gint1e_class_kernel<LA, LB>;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long long want = (npairs + nw - 1) / nw, cap = std::max(1LL, 8LL * sm_count / nbatch);
  const int split = static_cast<int>(std::max(1LL, std::min(want, cap)));
  kern<<<dim3(split, nbatch), k1eThreads, smem, st>>>(v, plist, npairs, coords, o);
  EVC_CHECK_LAUNCH();
  return 0;
}
#endif  // EVC_GCLASS_PART == 5

}  // namespace

#define EVC_ARGS sts[(k++) % nst], sm_count, nbatch, v, cq
#define EVC_UNITS(C) cunits + 2 * cunit_off[C], cunit_off[C + 1] - cunit_off[C], coords, o
#if EVC_GCLASS_PART == 0
int launch_gclass_part0(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o) {
  int rc;
  if ((rc = launch_one<0, 0, 0, 0, 0, 0>(EVC_ARGS, EVC_UNITS(0)))) return rc;
  if ((rc = launch_one<1, 0, 0, 0, 0, 0>(EVC_ARGS, EVC_UNITS(1)))) return rc;
  if ((rc = launch_one<1, 1, 0, 0, 0, 0>(EVC_ARGS, EVC_UNITS(2)))) return rc;
  return launch_one<1, 0, 1, 0, 0, 0>(EVC_ARGS, EVC_UNITS(3));
}

#elif EVC_GCLASS_PART == 1
int launch_gclass_part1(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o) {
  int rc;
  if ((rc = launch_one<1, 1, 1, 0, 0, 0>(EVC_ARGS, EVC_UNITS(4)))) return rc;
  if ((rc = launch_one<1, 1, 1, 0, 1, 0>(EVC_ARGS, EVC_UNITS(4)))) return rc;
  return launch_one<1, 1, 1, 0, 2, 0>(EVC_ARGS, EVC_UNITS(4));
}
#elif EVC_GCLASS_PART == 5
// one-electron classes: plist holds the ordered shell pairs of (s|s), (p|s), (s|p), (p|p) back to back
int launch_g1e(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* plist,
               const int* p_off, const double* coords, const GOut& o) {
  int rc;
  if ((rc = launch_1e<0, 0>(sts[(k++) % nst], sm_count, nbatch, v, plist + p_off[0], p_off[1] - p_off[0], coords, o))) return rc;
  if ((rc = launch_1e<1, 0>(sts[(k++) % nst], sm_count, nbatch, v, plist + p_off[1], p_off[2] - p_off[1], coords, o))) return rc;
  if ((rc = launch_1e<0, 1>(sts[(k++) % nst], sm_count, nbatch, v, plist + p_off[2], p_off[3] - p_off[2], coords, o))) return rc;
  return launch_1e<1, 1>(sts[(k++) % nst], sm_count, nbatch, v, plist + p_off[3], p_off[4] - p_off[3], coords, o);
}
#else
#define EVC_PPPP(N, JBV)                                                                                                 \
  int launch_gclass_part##N(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v,          \
                            const int32_t* cq,                                                                           \
                            const int32_t* cunits, const int* cunit_off, const double* coords, const GOut& o) {          \
    int rc;                                                                                                              \
    if ((rc = launch_one<1, 1, 1, 1, JBV, 0>(EVC_ARGS, EVC_UNITS(5)))) return rc;                                        \
    if ((rc = launch_one<1, 1, 1, 1, JBV, 1>(EVC_ARGS, EVC_UNITS(5)))) return rc;                                        \
    return launch_one<1, 1, 1, 1, JBV, 2>(EVC_ARGS, EVC_UNITS(5));                                                       \
  }
#if EVC_GCLASS_PART == 2
EVC_PPPP(2, 0)
#elif EVC_GCLASS_PART == 3
EVC_PPPP(3, 1)
#else
EVC_PPPP(4, 2)
#endif
#endif

}  // namespace evc_gint
