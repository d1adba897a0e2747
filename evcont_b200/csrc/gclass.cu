// K9g two-electron class kernels (see integrals_sp.cuh / integrals_sp.cu).  Compiled once per
// EVC_GCLASS_PART (Makefile): 0 = ssss, psss, ppss, psps; 1 = ppps; 2, 3, 4 = pppp with JB = 0, 1, 2.
#include "integrals_sp.cuh"

#include <algorithm>

#ifndef EVC_GCLASS_PART
#error "compile with -DEVC_GCLASS_PART=0..4"
#endif

namespace evc_gint {
namespace {

// =====================================================================================================
// Class kernels: contracted SHELL quartets, everything in registers.
//
// The generic kernel above evaluates one contracted AO quartet at a time and keeps its Hermite tables in
// per-thread local arrays (9 GB of local-memory DRAM traffic per Zundel launch, FP64 pipe 4 % active).
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
#pragma unroll 1
  for (int t = lig; t < tot; t += gs) {
    int r = t;
    const int il = r % sq.np[3]; r /= sq.np[3];
    const int ik = r % sq.np[2]; r /= sq.np[2];
    const int ij = r % sq.np[1]; r /= sq.np[1];
    const int ii = r;
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

template <int LA, int LB, int LC, int LD, int JB, int JC>
__global__ void __launch_bounds__(kGThreads)
gint2e_class_kernel(GView bs, const int32_t* __restrict__ qlist, const int32_t* __restrict__ units, int nunits,
                    const double* __restrict__ coords, GOut out) {
  using CI = ClassInfo<LA, LB, LC, LD>;
  const int n = bs.nao, natm = bs.natm;
  const int g = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kGThreads / 32;
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
  constexpr int nw = kGThreads / 32;
  // CTAs sized so that the grid covers the device a few times over
  const long long want = (nun + nw - 1) / nw, cap = std::max(1LL, 8LL * sm_count / nbatch);
  const int split = static_cast<int>(std::max(1LL, std::min(want, cap)));
  gint2e_class_kernel<LA, LB, LC, LD, JB, JC><<<dim3(split, nbatch), kGThreads, 0, st>>>(v, cq, cunits, nun, coords, o);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

#define EVC_ARGS st, sm_count, nbatch, v, cq
#define EVC_UNITS(C) cunits + 2 * cunit_off[C], cunit_off[C + 1] - cunit_off[C], coords, o
#if EVC_GCLASS_PART == 0
int launch_gclass_part0(cudaStream_t st, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o) {
  int rc;
  if ((rc = launch_one<0, 0, 0, 0, 0, 0>(EVC_ARGS, EVC_UNITS(0)))) return rc;
  if ((rc = launch_one<1, 0, 0, 0, 0, 0>(EVC_ARGS, EVC_UNITS(1)))) return rc;
  if ((rc = launch_one<1, 1, 0, 0, 0, 0>(EVC_ARGS, EVC_UNITS(2)))) return rc;
  return launch_one<1, 0, 1, 0, 0, 0>(EVC_ARGS, EVC_UNITS(3));
}
#elif EVC_GCLASS_PART == 1
int launch_gclass_part1(cudaStream_t st, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o) {
  int rc;
  if ((rc = launch_one<1, 1, 1, 0, 0, 0>(EVC_ARGS, EVC_UNITS(4)))) return rc;
  if ((rc = launch_one<1, 1, 1, 0, 1, 0>(EVC_ARGS, EVC_UNITS(4)))) return rc;
  return launch_one<1, 1, 1, 0, 2, 0>(EVC_ARGS, EVC_UNITS(4));
}
#else
#define EVC_PPPP(N, JBV)                                                                                                 \
  int launch_gclass_part##N(cudaStream_t st, int sm_count, int nbatch, const GView& v, const int32_t* cq,                \
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
