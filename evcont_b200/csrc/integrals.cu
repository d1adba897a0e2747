// K9: AO integrals over contracted s-type Gaussians for a batch of geometries, on the
// device.  Replaces, for s shells, what the reference asks
// PySCF / libcint for on every prediction step
// (evcont/ab_initio_gradients_loewdin.py:25, 130, 147, 177, 283-284, 338-339, 370, 378):
//
//   int1e_ovlp, scf.hf.get_hcore (int1e_kin + int1e_nuc), int2e, int1e_ipovlp,
//   grad.RHF.hcore_generator (int1e_ipkin, int1e_ipnuc, int1e_iprinv), int2e_ip1,
//   energy_nuc, grad_nuc
//
// and writes them in the evc_ao_bundle layout the prediction step consumes, so that a
// step needs only the nuclear coordinates from the host.
//
// Closed forms (Gaussian product theorem, Boys functions F0, F1).  For a primitive pair
// (i on A, j on B): p = a_i + a_j, mu = a_i a_j / p, bh = a_j / p, P = A - bh AB,
// Kc = w_i w_j exp(-mu |AB|^2).  Every derivative integral of a CONTRACTED pair / quartet
// reduces to a handful of scalar sums over primitives times the contracted-level vectors
// AB, CD, AC -- no per-primitive vector arithmetic beyond PQ for the Boys argument:
//
//   (ab|cd)        = sum w F0(T),        w = f(p,q) Kc_ab Kc_cd,  T = rho |PQ|^2
//   d/dA_x (ab|cd) = -2 AB_x M_ab - 2 VA_x,   d/dB_x = +2 AB_x M_ab - 2 VB_x
//   d/dC_x (ab|cd) = -2 CD_x M_cd + 2 VC_x,   d/dD_x = +2 CD_x M_cd + 2 VD_x
//   M_ab = sum mu_ab w F0,  u = rho w F1,  V = sum u PQ = AC U - AB U_b + CD U_d,
//   VB = AC U_b - AB U_bb + CD U_bd,  VD = AC U_d - AB U_bd + CD U_dd,  VA = V - VB, VC = V - VD
//   (nabla a b|cd) = - d/dA (ab|cd)                                       [int2e_ip1]
//
// One warp owns one contracted quartet (ab|cd), (ab) >= (cd); its lanes run over the
// primitive quartets, nine scalar accumulators each, reduced with a fixed butterfly
// (run-to-run bit-identical); lanes 0..7 then write the eight index permutations.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

constexpr int kBoysTop = 6;          // the table holds F_6 and exp(-T0) per grid point; F_5..F_0 by recursion
constexpr int kBoysPerUnit = 64;     // grid spacing 1/64
constexpr int kBoysTmax = 32;        // asymptotic form beyond
constexpr int kBoysN = kBoysTmax * kBoysPerUnit + 1;
constexpr double kScreen = 1.0e-17;  // primitive quartets with Schwarz bound below this are skipped
constexpr int kIntThreads = 256;
constexpr int kMaxAtoms = 64;

}  // namespace

struct evc_sbasis {
  int natm, nao, nprim, ndexp, npe, npc, maxpp, pairs_in_smem;
  // device tables
  int32_t* ao_atom;    // [nao]
  int32_t* ao_poff;    // [nao + 1]
  int32_t* aoslices;   // [natm][2]
  int32_t* prim_de;    // [nprim] index of the exponent among the distinct exponents
  double* prim_exp;    // [nprim]
  double* prim_wt;     // [nprim] contraction coefficient x primitive and contracted norms
  double* charges;     // [natm]
  double* pe1;         // [npe][4]: p, (pi/p)^1.5, 2 pi/p, sqrt(f(p,p))
  double* pp;          // [npe][npe][2]: rho, 2 pi^2.5 / (p q sqrt(p+q))
  double* boys;        // [kBoysN][2]: F_6(T0), exp(-T0), T0 = i/64
  int32_t* pc_off;     // [npc + 1] primitive-pair offsets of the contracted pairs a >= b
};

namespace {

// F_m(t), m = 0..mmax, accurate to long-double rounding (table generation, host)
void boys_host(int mmax, long double t, long double* out) {
  const long double et = expl(-t);
  long double term = 1.0L / (2 * mmax + 1), acc = term;
  for (int k = 1; k < 400; ++k) {
    term *= 2.0L * t / (2 * mmax + 2 * k + 1);
    acc += term;
    if (term < acc * 1e-22L) break;
  }
  out[mmax] = et * acc;
  for (int m = mmax; m > 0; --m) out[m - 1] = (2.0L * t * out[m] + et) / (2 * m - 1);
}

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

__device__ __forceinline__ void tri_unrank(int t, int& a, int& b) {
  int x = static_cast<int>((sqrtf(8.0f * static_cast<float>(t) + 1.0f) - 1.0f) * 0.5f);
  while (x * (x + 1) / 2 > t) --x;
  while ((x + 1) * (x + 2) / 2 <= t) ++x;
  a = x;
  b = t - x * (x + 1) / 2;
}

// F0(T), F1(T): the shared-memory table holds (F_6(T0), exp(-T0)) on the grid T0 = i/64 -- one
// 16-byte load per evaluation instead of one load per Taylor coefficient (the kernel is bound
// by shared-memory wavefronts, not FP64 issue).  F_5..F_0 at T0 follow from the downward
// recursion F_{k-1} = (2 T0 F_k + exp(-T0)) / (2k - 1) (all terms positive: stable), then a
// 5th-order Taylor series in d = T0 - T, F_m(T) = sum_k F_{m+k}(T0) d^k / k!  (|d| <= 1/128:
// truncation < 4e-16).  Asymptotic form beyond Tmax (error < exp(-32)/64).
__device__ __forceinline__ void boys01(double T, const double* __restrict__ tab, double& f0, double& f1) {
  if (T < static_cast<double>(kBoysTmax)) {
    // round T to the grid with the 1.5 * 2^52 trick: no F2I / I2F on the critical path
    const double r = fma(T, static_cast<double>(kBoysPerUnit), 6755399441055744.0);
    const int i = __double2loint(r);
    const double t0 = (r - 6755399441055744.0) * (1.0 / kBoysPerUnit);
    const double d = t0 - T, tt = t0 + t0;
    const double2 fe = *reinterpret_cast<const double2*>(tab + 2 * i);
    const double r6 = fe.x, e = fe.y;
    const double r5 = fma(tt, r6, e) * (1.0 / 11.0);
    const double r4 = fma(tt, r5, e) * (1.0 / 9.0);
    const double r3 = fma(tt, r4, e) * (1.0 / 7.0);
    const double r2 = fma(tt, r3, e) * (1.0 / 5.0);
    const double r1 = fma(tt, r2, e) * (1.0 / 3.0);
    const double r0 = fma(tt, r1, e);
    const double d2 = d * 0.5, d3 = d * (1.0 / 3.0), d4 = d * 0.25, d5 = d * 0.2;
    f0 = fma(d, fma(d2, fma(d3, fma(d4, fma(d5, r5, r4), r3), r2), r1), r0);
    f1 = fma(d, fma(d2, fma(d3, fma(d4, fma(d5, r6, r5), r4), r3), r2), r1);
  } else {
    // 1/sqrt(T): single-precision seed + two Newton steps (relative error < 1e-15 for T >= 32)
    double r = static_cast<double>(rsqrtf(static_cast<float>(T)));
    const double h = -0.5 * T;
    r = r * fma(h, r * r, 1.5);
    r = r * fma(h, r * r, 1.5);
    f0 = 0.88622692545275801365 * r;
    f1 = 0.5 * f0 * r * r;
  }
}

struct IntOut {
  double *ovlp, *hcore, *eri, *ipovlp, *hcore_deriv, *eri_ip1, *e_nuc, *grad_nuc;
  // packed two-electron output (evc_ao_integrals_s_packed): erip [np][pitch], eri_ip1p [3][n][n][np];
  // eri / eri_ip1 are NULL then
  double *erip, *ip1p;
  int pitch;
};

struct BasisView {
  int natm, nao, nprim, npe, npc, maxpp;
  const int32_t *ao_atom, *ao_poff, *prim_de, *pc_off;
  const double *prim_exp, *prim_wt, *charges, *pe1, *pp, *boys;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Primitive pairs of the contracted pairs a >= b (all ordered primitive combinations) for the warps
// [warp0, warp0 + nwarps) of the caller: records sorted by Schwarz bound sqrt([ij|ij]), descending.
__device__ __forceinline__ void build_pairs(const BasisView& bs, const double* __restrict__ R, int warp, int nwarps,
                                            int lane, double2* pra, double2* prb, float* sb, float* su) {
  for (int I = warp; I < bs.npc; I += nwarps) {
    int a, b;
    tri_unrank(I, a, b);
    const int pa0 = bs.ao_poff[a], na = bs.ao_poff[a + 1] - pa0;
    const int pb0 = bs.ao_poff[b], nb = bs.ao_poff[b + 1] - pb0;
    const double* A = R + 3 * bs.ao_atom[a];
    const double* B = R + 3 * bs.ao_atom[b];
    const double abx = A[0] - B[0], aby = A[1] - B[1], abz = A[2] - B[2];
    const double r2 = abx * abx + aby * aby + abz * abz;
    const int off = bs.pc_off[I], m = na * nb;
    auto record = [&](int t, float& bound) {
      const int i = t / nb, j = t - i * nb;
      const double ai = bs.prim_exp[pa0 + i], aj = bs.prim_exp[pb0 + j];
      const int di = bs.prim_de[pa0 + i], dj = bs.prim_de[pb0 + j];
      const int pe = di >= dj ? tri(di, dj) : tri(dj, di);
      const double p = bs.pe1[4 * pe];
      const double mu = ai * aj / p;
      double4 rec;
      rec.x = aj / p;
      rec.y = mu;
      rec.z = bs.prim_wt[pa0 + i] * bs.prim_wt[pb0 + j] * exp(-mu * r2);
      rec.w = __longlong_as_double(static_cast<long long>(pe));
      // sqrt([ij|ij]) = |Kc| sqrt(f(p,p)); rounded up so that the float never under-estimates
      bound = static_cast<float>(fabs(rec.z) * bs.pe1[4 * pe + 3]) * 1.000001f;
      return rec;
    };
    for (int t = lane; t < m; t += 32) {
      float bd;
      record(t, bd);
      su[off + t] = bd;
    }
    __syncwarp();
    for (int t = lane; t < m; t += 32) {
      float bd;
      const double4 rec = record(t, bd);
      int rank = 0;
      for (int u = 0; u < m; ++u) {
        const float o = su[off + u];
        rank += (o > bd || (o == bd && u < t)) ? 1 : 0;
      }
      pra[off + rank] = make_double2(rec.x, rec.y);
      prb[off + rank] = make_double2(rec.z, rec.w);
      sb[off + rank] = bd;
    }
  }
}

// global-memory pair tables (systems whose tables exceed shared memory): per geometry
//   [maxpp] double2 (bh, mu) | [maxpp] double2 (Kc, pe) | [maxpp] float bounds | [maxpp] float scratch
__host__ __device__ inline size_t pair_table_bytes(int maxpp) {
  return (static_cast<size_t>(maxpp) * (2 * sizeof(double2) + 2 * sizeof(float)) + 255) / 256 * 256;
}

__global__ void __launch_bounds__(kIntThreads)
spairs_kernel(BasisView bs, const double* __restrict__ coords, char* __restrict__ tables) {
  __shared__ double R[3 * kMaxAtoms];
  const int g = blockIdx.x, tid = threadIdx.x;
  for (int k = tid; k < 3 * bs.natm; k += kIntThreads) R[k] = coords[static_cast<int64_t>(g) * bs.natm * 3 + k];
  __syncthreads();
  char* base = tables + static_cast<size_t>(g) * pair_table_bytes(bs.maxpp);
  double2* pra = reinterpret_cast<double2*>(base);
  double2* prb = pra + bs.maxpp;
  float* sb = reinterpret_cast<float*>(prb + bs.maxpp);
  build_pairs(bs, R, tid >> 5, kIntThreads / 32, tid & 31, pra, prb, sb, sb + bs.maxpp);
}

// shared memory: boys table | pp table (if it fits) | atom coordinates | primitive pairs | bounds
//   primitive pair record, two arrays of double2: (bh = a_j/p, mu) and (Kc, pe as an int in the
//   low word) -- 16-byte stride keeps the lane-consecutive 16-byte loads conflict-free; the records
//   of a contracted pair are sorted by their Schwarz bound sqrt([ij|ij]) (descending, in sb[])
//   so that the significant primitive quartets of (ab|cd) form a leading rectangle.
//   PAIRS_IN_SMEM == false: the tables were built by spairs_kernel in global memory (L2-resident).
template <bool PP_IN_SMEM, bool PAIRS_IN_SMEM>
__global__ void __launch_bounds__(kIntThreads, 2)
sint_kernel(BasisView bs, const double* __restrict__ coords, IntOut out, const char* __restrict__ tables) {
  extern __shared__ __align__(16) double sm[];
  const int n = bs.nao, natm = bs.natm, npe = bs.npe;
  double* boys = sm;
  double* ppt = boys + 2 * kBoysN + 2;                            // [npe][npe][2]
  double* R = ppt + (PP_IN_SMEM ? 2 * npe * npe : 0);             // [natm][3]
  const int g = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kIntThreads / 32;
  const double* Rg = coords + static_cast<int64_t>(g) * natm * 3;
  const double2* pra;
  const double2* prb;
  const float* sb;

  if (PP_IN_SMEM)
    for (int k = tid; k < 2 * npe * npe; k += kIntThreads) ppt[k] = __ldg(bs.pp + k);
  for (int k = tid; k < 3 * natm; k += kIntThreads) R[k] = Rg[k];
  __syncthreads();
  if (PAIRS_IN_SMEM) {
    double2* wa = reinterpret_cast<double2*>(R + ((3 * natm + 1) & ~1));  // [maxpp] (bh, mu)
    double2* wb = wa + bs.maxpp;                                          // [maxpp] (Kc, pe)
    float* ws = reinterpret_cast<float*>(wb + bs.maxpp);                  // [maxpp] sorted bounds
    // unsorted bounds (scratch of the pair build): in the Boys region, which is filled afterwards
    float* su = (static_cast<size_t>(bs.maxpp) * sizeof(float) <= 2 * kBoysN * sizeof(double))
                    ? reinterpret_cast<float*>(boys) : ws + bs.maxpp;
    build_pairs(bs, R, warp, NW, lane, wa, wb, ws, su);
    pra = wa; prb = wb; sb = ws;
    __syncthreads();
  } else {
    const char* base = tables + static_cast<size_t>(g) * pair_table_bytes(bs.maxpp);
    pra = reinterpret_cast<const double2*>(base);
    prb = pra + bs.maxpp;
    sb = reinterpret_cast<const float*>(prb + bs.maxpp);
  }
  for (int k = tid; k < 2 * kBoysN; k += kIntThreads) boys[k] = __ldg(bs.boys + k);
  __syncthreads();

  const int64_t n2 = static_cast<int64_t>(n) * n, n3 = n2 * n, n4 = n2 * n2;
  const bool packed = out.erip != nullptr;
  const int64_t npk = bs.npc;  // contracted pairs = n (n + 1) / 2
  double* eri = packed ? out.erip + static_cast<int64_t>(g) * npk * out.pitch : out.eri + static_cast<int64_t>(g) * n4;
  double* ip1 = packed ? out.ip1p + static_cast<int64_t>(g) * 3 * n2 * npk : out.eri_ip1 + static_cast<int64_t>(g) * 3 * n4;
  const double* ppg = PP_IN_SMEM ? ppt : bs.pp;

  // ---- two-electron part: contracted quartets (I >= K) dealt to warps -----------------
  const int nq = bs.npc * (bs.npc + 1) / 2;
  for (int q = blockIdx.x * NW + warp; q < nq; q += gridDim.x * NW) {
    int I, K, a, b, c, d;
    tri_unrank(q, I, K);
    tri_unrank(I, a, b);
    tri_unrank(K, c, d);
    const double* A = R + 3 * bs.ao_atom[a];
    const double* B = R + 3 * bs.ao_atom[b];
    const double* Cc = R + 3 * bs.ao_atom[c];
    const double* D = R + 3 * bs.ao_atom[d];
    const double abx = A[0] - B[0], aby = A[1] - B[1], abz = A[2] - B[2];
    const double cdx = Cc[0] - D[0], cdy = Cc[1] - D[1], cdz = Cc[2] - D[2];
    const double acx = A[0] - Cc[0], acy = A[1] - Cc[1], acz = A[2] - Cc[2];
    const int offI = bs.pc_off[I], offK = bs.pc_off[K];
    int mI = bs.pc_off[I + 1] - offI, mK = bs.pc_off[K + 1] - offK;
    {
      // leading rectangle of primitive quartets whose Schwarz bound reaches kScreen
      const float thr = static_cast<float>(kScreen);
      const float topI = sb[offI], topK = sb[offK];
      int cI = 0, cK = 0;
      for (int l = lane; l < mI; l += 32) cI += (sb[offI + l] * topK >= thr) ? 1 : 0;
      for (int l = lane; l < mK; l += 32) cK += (sb[offK + l] * topI >= thr) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cI += __shfl_xor_sync(0xffffffffu, cI, o);
        cK += __shfl_xor_sync(0xffffffffu, cK, o);
      }
      mI = cI;
      mK = cK;
    }
    double E = 0, Mab = 0, Mcd = 0, U = 0, Ub = 0, Ud = 0, Ubb = 0, Ubd = 0, Udd = 0;
    const int tot = mI * mK;
    int bi = 0, ki = lane;
    while (ki >= mK && mK > 0) { ki -= mK; ++bi; }
    // lanes run over the flattened rectangle of significant primitive quartets
    for (int t = lane; t < tot; t += 32) {
      const double2 b1 = pra[offI + bi], b2 = prb[offI + bi];   // (bh, mu), (Kc, pe)
      const double2 k1 = pra[offK + ki], k2 = prb[offK + ki];
      const int peb = static_cast<int>(__double_as_longlong(b2.y));
      const int pek = static_cast<int>(__double_as_longlong(k2.y));
      const double2 rf = *reinterpret_cast<const double2*>(ppg + 2 * (peb * npe + pek));
      const double px = fma(k1.x, cdx, fma(-b1.x, abx, acx));
      const double py = fma(k1.x, cdy, fma(-b1.x, aby, acy));
      const double pz = fma(k1.x, cdz, fma(-b1.x, abz, acz));
      const double T = rf.x * fma(px, px, fma(py, py, pz * pz));
      double f0, f1;
      boys01(T, boys, f0, f1);
      const double w = rf.y * b2.x * k2.x;
      const double i0 = w * f0, u = rf.x * w * f1;
      const double ub = b1.x * u, ud = k1.x * u;
      E += i0;
      Mab = fma(b1.y, i0, Mab);
      Mcd = fma(k1.y, i0, Mcd);
      U += u;
      Ub += ub;
      Ud += ud;
      Ubb = fma(b1.x, ub, Ubb);
      Ubd = fma(b1.x, ud, Ubd);
      Udd = fma(k1.x, ud, Udd);
      ki += 32;
      while (ki >= mK) { ki -= mK; ++bi; }
    }
    E = warp_sum(E); Mab = warp_sum(Mab); Mcd = warp_sum(Mcd);
    U = warp_sum(U); Ub = warp_sum(Ub); Ud = warp_sum(Ud);
    Ubb = warp_sum(Ubb); Ubd = warp_sum(Ubd); Udd = warp_sum(Udd);
    if (lane < 8) {
      // lane r writes index permutation r; the differentiated function is the first index
      const int who = lane >> 1;  // 0: a, 1: b, 2: c, 3: d
      int i0, i1, i2, i3;
      if (who == 0) { i0 = a; i1 = b; i2 = (lane & 1) ? d : c; i3 = (lane & 1) ? c : d; }
      else if (who == 1) { i0 = b; i1 = a; i2 = (lane & 1) ? d : c; i3 = (lane & 1) ? c : d; }
      else if (who == 2) { i0 = c; i1 = d; i2 = (lane & 1) ? b : a; i3 = (lane & 1) ? a : b; }
      else { i0 = d; i1 = c; i2 = (lane & 1) ? b : a; i3 = (lane & 1) ? a : b; }
      // duplicates (a == b, c == d, I == K) must carry bit-identical values: fold them
      int eff = who;
      if (eff == 1 && a == b) eff = 0;
      if (eff == 3 && c == d) eff = 2;
      if (I == K) eff -= (eff >= 2) ? 2 : 0;
      double gx, gy, gz;  // d/d(centre of the first index) of (ab|cd)
      {
        const double vx = acx * U - abx * Ub + cdx * Ud, vy = acy * U - aby * Ub + cdy * Ud,
                     vz = acz * U - abz * Ub + cdz * Ud;
        const double vbx = acx * Ub - abx * Ubb + cdx * Ubd, vby = acy * Ub - aby * Ubb + cdy * Ubd,
                     vbz = acz * Ub - abz * Ubb + cdz * Ubd;
        const double vdx = acx * Ud - abx * Ubd + cdx * Udd, vdy = acy * Ud - aby * Ubd + cdy * Udd,
                     vdz = acz * Ud - abz * Ubd + cdz * Udd;
        if (eff == 0) {
          gx = -2.0 * abx * Mab - 2.0 * (vx - vbx); gy = -2.0 * aby * Mab - 2.0 * (vy - vby);
          gz = -2.0 * abz * Mab - 2.0 * (vz - vbz);
        } else if (eff == 1) {
          gx = 2.0 * abx * Mab - 2.0 * vbx; gy = 2.0 * aby * Mab - 2.0 * vby; gz = 2.0 * abz * Mab - 2.0 * vbz;
        } else if (eff == 2) {
          gx = -2.0 * cdx * Mcd + 2.0 * (vx - vdx); gy = -2.0 * cdy * Mcd + 2.0 * (vy - vdy);
          gz = -2.0 * cdz * Mcd + 2.0 * (vz - vdz);
        } else {
          gx = 2.0 * cdx * Mcd + 2.0 * vdx; gy = 2.0 * cdy * Mcd + 2.0 * vdy; gz = 2.0 * cdz * Mcd + 2.0 * vdz;
        }
      }
      if (!packed) {
        const int64_t idx = i0 * n3 + i1 * n2 + i2 * n + i3;
        eri[idx] = E;
        ip1[idx] = -gx;
        ip1[n4 + idx] = -gy;
        ip1[2 * n4 + idx] = -gz;
      } else if ((lane & 1) == 0) {
        // one representative per differentiated function: (d i0 i1 | pair of the other side)
        const int64_t idx = (static_cast<int64_t>(i0) * n + i1) * npk + (who < 2 ? K : I);
        const int64_t xs = n2 * npk;
        ip1[idx] = -gx;
        ip1[xs + idx] = -gy;
        ip1[2 * xs + idx] = -gz;
        if (who == 0) eri[static_cast<int64_t>(I) * out.pitch + K] = E;
        if (who == 2) eri[static_cast<int64_t>(K) * out.pitch + I] = E;
      }
    }
  }

  // ---- one-electron part: contracted pairs a >= b dealt to warps, lanes over nuclei -----
  double* S = out.ovlp + static_cast<int64_t>(g) * n2;
  double* Hc = out.hcore + static_cast<int64_t>(g) * n2;
  double* ipo = out.ipovlp + static_cast<int64_t>(g) * 3 * n2;
  double* hd = out.hcore_deriv + static_cast<int64_t>(g) * natm * 3 * n2;
  for (int I = blockIdx.x * NW + warp; I < bs.npc; I += gridDim.x * NW) {
    int a, b;
    tri_unrank(I, a, b);
    const int atA = bs.ao_atom[a], atB = bs.ao_atom[b];
    const double* A = R + 3 * atA;
    const double* B = R + 3 * atB;
    const double abx = A[0] - B[0], aby = A[1] - B[1], abz = A[2] - B[2];
    const double r2 = abx * abx + aby * aby + abz * abz;
    const int off = bs.pc_off[I], m = bs.pc_off[I + 1] - off;
    // overlap-type sums: lanes over primitive pairs
    double s0 = 0, s1 = 0, t0 = 0, t1 = 0;
    for (int t = lane; t < m; t += 32) {
      const double2 q1 = pra[off + t], q2 = prb[off + t];
      const double4 pr = make_double4(q1.x, q1.y, q2.x, q2.y);
      const int pe = static_cast<int>(__double_as_longlong(pr.w));
      const double sp = bs.pe1[4 * pe + 1] * pr.z, mu = pr.y;
      s0 += sp;
      s1 = fma(mu, sp, s1);
      t0 = fma(mu * (3.0 - 2.0 * mu * r2), sp, t0);
      t1 = fma(mu * mu * (5.0 - 2.0 * mu * r2), sp, t1);
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); t0 = warp_sum(t0); t1 = warp_sum(t1);
    // nuclear attraction: lane = nucleus C (strided), serial over primitive pairs
    double vsum = 0.0;                              // sum_C Z_C <a|1/r_C|b>
    double nax = 0, nay = 0, naz = 0;               // sum_C Z_C <nabla a|1/r_C|b>
    double nbx = 0, nby = 0, nbz = 0;               // sum_C Z_C <nabla b|1/r_C|a>
    for (int c0 = 0; c0 < natm; c0 += 32) {
      const int C = c0 + lane;
      double R0 = 0, Rmu = 0, Ra = 0, Rb = 0, Rab = 0;
      double acx = 0, acy = 0, acz = 0, z = 0;
      if (C < natm) {
        acx = A[0] - R[3 * C]; acy = A[1] - R[3 * C + 1]; acz = A[2] - R[3 * C + 2];
        z = bs.charges[C];
        for (int t = 0; t < m; ++t) {
          const double2 q1 = pra[off + t], q2 = prb[off + t];
      const double4 pr = make_double4(q1.x, q1.y, q2.x, q2.y);
          const int pe = static_cast<int>(__double_as_longlong(pr.w));
          const double p = bs.pe1[4 * pe], gpre = bs.pe1[4 * pe + 2] * pr.z;
          const double px = fma(-pr.x, abx, acx), py = fma(-pr.x, aby, acy), pz = fma(-pr.x, abz, acz);
          double f0, f1;
          boys01(p * fma(px, px, fma(py, py, pz * pz)), boys, f0, f1);
          const double g0 = gpre * f0, g1 = gpre * p * f1;   // g1 = g p F1
          R0 += g0;
          Rmu = fma(pr.y, g0, Rmu);
          Ra = fma(1.0 - pr.x, g1, Ra);                      // alpha = (1 - bh) p
          Rb = fma(pr.x, g1, Rb);                            // beta  = bh p
          Rab = fma(pr.x * (1.0 - pr.x), g1, Rab);           // alpha bh = beta ah
        }
      }
      // <nabla a|1/r_C|b>_x = 2 AB_x (Rmu - Rab) + 2 (A - C)_x Ra
      // <nabla b|1/r_C|a>_x = -2 AB_x (Rmu - Rab) + 2 (B - C)_x Rb
      const double k = 2.0 * (Rmu - Rab);
      const double iax = k * abx + 2.0 * acx * Ra, iay = k * aby + 2.0 * acy * Ra, iaz = k * abz + 2.0 * acz * Ra;
      const double ibx = -k * abx + 2.0 * (acx - abx) * Rb, iby = -k * aby + 2.0 * (acy - aby) * Rb,
                   ibz = -k * abz + 2.0 * (acz - abz) * Rb;
      vsum += z * R0;
      nax += z * iax; nay += z * iay; naz += z * iaz;
      nbx += z * ibx; nby += z * iby; nbz += z * ibz;
      // first part of hcore_generator()(C): -Z_C (iprinv[a,b] + iprinv[b,a]); the aoslice part follows
      if (C < natm) {
        double* h = hd + static_cast<int64_t>(C) * 3 * n2;
        const double hx = -z * (iax + ibx), hy = -z * (iay + iby), hz = -z * (iaz + ibz);
        h[a * n + b] = hx; h[n2 + a * n + b] = hy; h[2 * n2 + a * n + b] = hz;
        h[b * n + a] = hx; h[n2 + b * n + a] = hy; h[2 * n2 + b * n + a] = hz;
      }
    }
    vsum = warp_sum(vsum);
    nax = warp_sum(nax); nay = warp_sum(nay); naz = warp_sum(naz);
    nbx = warp_sum(nbx); nby = warp_sum(nby); nbz = warp_sum(nbz);
    __syncwarp();
    if (lane == 0) {
      S[a * n + b] = s0;
      S[b * n + a] = s0;
      const double h = t0 - vsum;
      Hc[a * n + b] = h;
      Hc[b * n + a] = h;
      // <nabla a|b> = 2 AB mu S ;  <nabla b|a> = -that
      const double ox = 2.0 * abx * s1, oy = 2.0 * aby * s1, oz = 2.0 * abz * s1;
      ipo[a * n + b] = ox; ipo[n2 + a * n + b] = oy; ipo[2 * n2 + a * n + b] = oz;
      ipo[b * n + a] = -ox; ipo[n2 + b * n + a] = -oy; ipo[2 * n2 + b * n + a] = -oz;
      // aoslice part: v[:, p0:p1] -= (ipkin + ipnuc)[:, p0:p1], then v + v^T:
      //   hd[atom(a)][x][a][b] (+ mirror) -= (ipkin + ipnuc)[x][a][b]
      //   hd[atom(b)][x][a][b] (+ mirror) -= (ipkin + ipnuc)[x][b][a]
      // ipkin[x][a][b] = 2 AB_x t1, ipkin[x][b][a] = -2 AB_x t1;  ipnuc = -sum_C Z_C iprinv
      const double kax = 2.0 * abx * t1 - nax, kay = 2.0 * aby * t1 - nay, kaz = 2.0 * abz * t1 - naz;
      const double kbx = -2.0 * abx * t1 - nbx, kby = -2.0 * aby * t1 - nby, kbz = -2.0 * abz * t1 - nbz;
      double* ha = hd + static_cast<int64_t>(atA) * 3 * n2;
      double* hb = hd + static_cast<int64_t>(atB) * 3 * n2;
      if (a != b) {
        ha[a * n + b] -= kax; ha[n2 + a * n + b] -= kay; ha[2 * n2 + a * n + b] -= kaz;
        ha[b * n + a] -= kax; ha[n2 + b * n + a] -= kay; ha[2 * n2 + b * n + a] -= kaz;
        hb[a * n + b] -= kbx; hb[n2 + a * n + b] -= kby; hb[2 * n2 + a * n + b] -= kbz;
        hb[b * n + a] -= kbx; hb[n2 + b * n + a] -= kby; hb[2 * n2 + b * n + a] -= kbz;
      } else {
        // v[x][a][a] -= k;  (v + v^T)[a][a] gets it twice
        ha[a * n + a] -= 2.0 * kax; ha[n2 + a * n + a] -= 2.0 * kay; ha[2 * n2 + a * n + a] -= 2.0 * kaz;
      }
    }
    __syncwarp();
  }

  // ---- nuclear repulsion and its gradient ------------------------------------------------
  if (blockIdx.x == 0 && warp == 0) {
    double e = 0.0;
    for (int c0 = 0; c0 < natm; c0 += 32) {
      const int A = c0 + lane;
      double gx = 0, gy = 0, gz = 0;
      if (A < natm) {
        const double za = bs.charges[A];
        for (int B = 0; B < natm; ++B) {
          if (B == A) continue;
          const double dx = R[3 * A] - R[3 * B], dy = R[3 * A + 1] - R[3 * B + 1], dz = R[3 * A + 2] - R[3 * B + 2];
          const double r2 = dx * dx + dy * dy + dz * dz, ri = rsqrt(r2), zz = za * bs.charges[B];
          if (B < A) e += zz * ri;
          const double f = zz * ri * ri * ri;
          gx -= f * dx; gy -= f * dy; gz -= f * dz;
        }
        double* gn = out.grad_nuc + (static_cast<int64_t>(g) * natm + A) * 3;
        gn[0] = gx; gn[1] = gy; gn[2] = gz;
      }
    }
    e = warp_sum(e);
    if (lane == 0) out.e_nuc[g] = e;
  }
}

size_t sint_smem_bytes(const evc_sbasis* b, bool pp_in_smem, bool pairs_in_smem) {
  size_t d = 2 * static_cast<size_t>(kBoysN) + 2 + (pp_in_smem ? 2 * static_cast<size_t>(b->npe) * b->npe : 0) +
             ((3 * static_cast<size_t>(b->natm) + 1) & ~static_cast<size_t>(1));
  if (!pairs_in_smem) return d * sizeof(double);
  const bool su_in_boys = static_cast<size_t>(b->maxpp) * sizeof(float) <= 2 * kBoysN * sizeof(double);
  return d * sizeof(double) + static_cast<size_t>(b->maxpp) * (2 * sizeof(double2) + (su_in_boys ? 1 : 2) * sizeof(float));
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
  EVC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), std::max<size_t>(1, src.size()) * sizeof(T)));
  if (!src.empty())
    EVC_CHECK_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

extern "C" {

int evc_sbasis_create(evc_ctx* ctx, int natm, const double* charges_host, int nao, const int32_t* ao_atom_host,
                      const int32_t* ao_nprim_host, const double* prim_exp_host, const double* prim_wt_host,
                      evc_sbasis** out) {
  EVC_REQUIRE(ctx && charges_host && ao_atom_host && ao_nprim_host && prim_exp_host && prim_wt_host && out,
              "evc_sbasis_create: NULL argument");
  EVC_REQUIRE(natm >= 1 && natm <= kMaxAtoms && nao >= 1 && nao <= 64, "evc_sbasis_create: natm=%d nao=%d unsupported",
              natm, nao);
  EVC_CHECK_CUDA(cudaSetDevice(ctx->device));
  std::vector<int32_t> ao_atom(ao_atom_host, ao_atom_host + nao), poff(nao + 1, 0), slices(2 * natm, 0);
  for (int a = 0; a < nao; ++a) {
    EVC_REQUIRE(ao_nprim_host[a] >= 1 && ao_nprim_host[a] <= 16, "evc_sbasis_create: AO %d has %d primitives", a,
                ao_nprim_host[a]);
    EVC_REQUIRE(ao_atom[a] >= 0 && ao_atom[a] < natm && (a == 0 || ao_atom[a] >= ao_atom[a - 1]),
                "evc_sbasis_create: AOs must be grouped by atom, in atom order");
    poff[a + 1] = poff[a] + ao_nprim_host[a];
  }
  // aoslice_by_atom: (first AO, one past the last AO) per atom
  {
    int a = 0;
    for (int A = 0; A < natm; ++A) {
      slices[2 * A] = a;
      while (a < nao && ao_atom[a] == A) ++a;
      slices[2 * A + 1] = a;
    }
  }
  const int nprim = poff[nao];
  std::vector<double> ex(prim_exp_host, prim_exp_host + nprim), wt(prim_wt_host, prim_wt_host + nprim);
  std::vector<double> dexp;
  std::vector<int32_t> de(nprim);
  for (int k = 0; k < nprim; ++k) {
    EVC_REQUIRE(ex[k] > 0.0, "evc_sbasis_create: exponent %d is not positive", k);
    int f = -1;
    for (size_t q = 0; q < dexp.size(); ++q)
      if (dexp[q] == ex[k]) f = static_cast<int>(q);
    if (f < 0) {
      f = static_cast<int>(dexp.size());
      dexp.push_back(ex[k]);
    }
    de[k] = f;
  }
  const int nd = static_cast<int>(dexp.size()), npe = nd * (nd + 1) / 2;
  EVC_REQUIRE(nd <= 64, "evc_sbasis_create: %d distinct exponents (max 64)", nd);
  std::vector<double> pe1(4 * static_cast<size_t>(npe)), pp(2 * static_cast<size_t>(npe) * npe);
  const long double pi = 3.141592653589793238462643383279502884L;
  for (int i = 0; i < nd; ++i)
    for (int j = 0; j <= i; ++j) {
      const int e = i * (i + 1) / 2 + j;
      const long double p = static_cast<long double>(dexp[i]) + dexp[j];
      pe1[4 * e] = static_cast<double>(p);
      pe1[4 * e + 1] = static_cast<double>(powl(pi / p, 1.5L));
      pe1[4 * e + 2] = static_cast<double>(2.0L * pi / p);
      // sqrt(f(p, p)) = sqrt(2 pi^2.5 / (p^2 sqrt(2p))): Schwarz factor of a primitive pair
      pe1[4 * e + 3] = static_cast<double>(sqrtl(2.0L * powl(pi, 2.5L) / (p * p * sqrtl(2.0L * p))));
    }
  for (int e = 0; e < npe; ++e)
    for (int f = 0; f < npe; ++f) {
      const long double p = pe1[4 * e], q = pe1[4 * f];
      pp[2 * (static_cast<size_t>(e) * npe + f)] = static_cast<double>(p * q / (p + q));
      pp[2 * (static_cast<size_t>(e) * npe + f) + 1] =
          static_cast<double>(2.0L * powl(pi, 2.5L) / (p * q * sqrtl(p + q)));
    }
  std::vector<double> boys(2 * static_cast<size_t>(kBoysN));
  for (int i = 0; i < kBoysN; ++i) {
    long double f[kBoysTop + 1];
    const long double t0 = static_cast<long double>(i) / kBoysPerUnit;
    boys_host(kBoysTop, t0, f);
    boys[2 * static_cast<size_t>(i)] = static_cast<double>(f[kBoysTop]);
    boys[2 * static_cast<size_t>(i) + 1] = static_cast<double>(expl(-t0));
  }
  const int npc = nao * (nao + 1) / 2;
  std::vector<int32_t> pc_off(npc + 1, 0);
  for (int a = 0; a < nao; ++a)
    for (int b = 0; b <= a; ++b) {
      const int I = a * (a + 1) / 2 + b;
      pc_off[I + 1] = ao_nprim_host[a] * ao_nprim_host[b];
    }
  for (int I = 0; I < npc; ++I) pc_off[I + 1] += pc_off[I];

  evc_sbasis* b = new evc_sbasis();
  b->natm = natm; b->nao = nao; b->nprim = nprim; b->ndexp = nd; b->npe = npe; b->npc = npc;
  b->maxpp = pc_off[npc];
  std::vector<double> ch(charges_host, charges_host + natm);
  int rc = 0;
  if ((rc = upload(&b->ao_atom, ao_atom)) || (rc = upload(&b->ao_poff, poff)) || (rc = upload(&b->aoslices, slices)) ||
      (rc = upload(&b->prim_de, de)) || (rc = upload(&b->prim_exp, ex)) || (rc = upload(&b->prim_wt, wt)) ||
      (rc = upload(&b->charges, ch)) || (rc = upload(&b->pe1, pe1)) || (rc = upload(&b->pp, pp)) ||
      (rc = upload(&b->boys, boys)) || (rc = upload(&b->pc_off, pc_off))) {
    delete b;
    return rc;
  }
  b->pairs_in_smem = sint_smem_bytes(b, false, true) <= 113 * 1024 ? 1 : 0;  // two CTAs per SM
  *out = b;
  return 0;
}

int evc_sbasis_destroy(evc_sbasis* b) {
  if (b) {
    cudaFree(b->ao_atom); cudaFree(b->ao_poff); cudaFree(b->aoslices); cudaFree(b->prim_de); cudaFree(b->prim_exp);
    cudaFree(b->prim_wt); cudaFree(b->charges); cudaFree(b->pe1); cudaFree(b->pp); cudaFree(b->boys);
    cudaFree(b->pc_off);
  }
  delete b;
  return 0;
}

int evc_sbasis_nao(const evc_sbasis* b) { return b ? b->nao : -1; }
int evc_sbasis_natm(const evc_sbasis* b) { return b ? b->natm : -1; }
const int32_t* evc_sbasis_aoslices(const evc_sbasis* b) { return b ? b->aoslices : nullptr; }

int evc_ao_integrals_s_workspace_bytes(const evc_sbasis* b, int nbatch, size_t* bytes) {
  EVC_REQUIRE(b && bytes && nbatch >= 0, "evc_ao_integrals_s_workspace_bytes: bad arguments");
  *bytes = b->pairs_in_smem ? 256 : static_cast<size_t>(nbatch) * pair_table_bytes(b->maxpp) + 256;
  return 0;
}

}  // extern "C"

namespace {
int ao_integrals_s_impl(evc_ctx* ctx, const evc_sbasis* b, int nbatch, const double* coords, double* ovlp,
                        double* hcore, double* eri, double* ipovlp, double* hcore_deriv, double* eri_ip1,
                        double* e_nuc, double* grad_nuc, void* workspace, size_t workspace_bytes, bool packed) {
  EVC_REQUIRE(ctx && b && coords && ovlp && hcore && eri && ipovlp && hcore_deriv && eri_ip1 && e_nuc && grad_nuc,
              "evc_ao_integrals_s: NULL argument");
  if (nbatch <= 0) return 0;
  size_t need = 0;
  evc_ao_integrals_s_workspace_bytes(b, nbatch, &need);
  EVC_REQUIRE(b->pairs_in_smem || (workspace && workspace_bytes >= need),
              "evc_ao_integrals_s: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  BasisView v;
  v.natm = b->natm; v.nao = b->nao; v.nprim = b->nprim; v.npe = b->npe; v.npc = b->npc; v.maxpp = b->maxpp;
  v.ao_atom = b->ao_atom; v.ao_poff = b->ao_poff; v.prim_de = b->prim_de; v.pc_off = b->pc_off;
  v.prim_exp = b->prim_exp; v.prim_wt = b->prim_wt; v.charges = b->charges; v.pe1 = b->pe1; v.pp = b->pp;
  v.boys = b->boys;
  IntOut o{ovlp, hcore, eri, ipovlp, hcore_deriv, eri_ip1, e_nuc, grad_nuc, nullptr, nullptr, 0};
  if (packed) {
    o.erip = eri; o.ip1p = eri_ip1; o.eri = nullptr; o.eri_ip1 = nullptr;
    o.pitch = evc_erip_pitch(b->nao);
  }
  // few geometries: several CTAs per geometry so that the whole GPU works on them
  int split = 1;
  const long long nq = static_cast<long long>(b->npc) * (b->npc + 1) / 2;
  const int nw = kIntThreads / 32;
  while (static_cast<long long>(nbatch) * split < 2LL * ctx->sm_count && static_cast<long long>(split) * nw * 4 <= nq &&
         split < 1024)
    split *= 2;
  const bool pairs_in = b->pairs_in_smem != 0;
  const bool pp_in = sint_smem_bytes(b, true, pairs_in) <= 113 * 1024;  // two CTAs per SM
  const size_t smem = sint_smem_bytes(b, pp_in, pairs_in);
  char* tables = static_cast<char*>(workspace);
  if (!pairs_in) {
    spairs_kernel<<<nbatch, kIntThreads, 0, ctx->stream>>>(v, coords, tables);
    EVC_CHECK_LAUNCH();
  }
  dim3 grid(split, nbatch);
#define EVC_SINT(PP, PR)                                                                                     \
  do {                                                                                                       \
    EVC_CHECK_CUDA(cudaFuncSetAttribute(sint_kernel<PP, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                        static_cast<int>(smem)));                                            \
    sint_kernel<PP, PR><<<grid, kIntThreads, smem, ctx->stream>>>(v, coords, o, tables);                     \
  } while (0)
  if (pairs_in) { if (pp_in) EVC_SINT(true, true); else EVC_SINT(false, true); }
  else { if (pp_in) EVC_SINT(true, false); else EVC_SINT(false, false); }
#undef EVC_SINT
  EVC_CHECK_LAUNCH();
  return 0;
}
}  // namespace

extern "C" {

int evc_ao_integrals_s(evc_ctx* ctx, const evc_sbasis* b, int nbatch, const double* coords, double* ovlp,
                       double* hcore, double* eri, double* ipovlp, double* hcore_deriv, double* eri_ip1,
                       double* e_nuc, double* grad_nuc, void* workspace, size_t workspace_bytes) {
  return ao_integrals_s_impl(ctx, b, nbatch, coords, ovlp, hcore, eri, ipovlp, hcore_deriv, eri_ip1, e_nuc, grad_nuc,
                             workspace, workspace_bytes, false);
}

int evc_ao_integrals_s_packed(evc_ctx* ctx, const evc_sbasis* b, int nbatch, const double* coords, double* ovlp,
                              double* hcore, double* erip, double* ipovlp, double* hcore_deriv, double* eri_ip1p,
                              double* e_nuc, double* grad_nuc, void* workspace, size_t workspace_bytes) {
  return ao_integrals_s_impl(ctx, b, nbatch, coords, ovlp, hcore, erip, ipovlp, hcore_deriv, eri_ip1p, e_nuc,
                             grad_nuc, workspace, workspace_bytes, true);
}

}  // extern "C"
