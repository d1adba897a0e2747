// f4 observables of the predicted one-body density matrix, batched over geometries: the dipole moment
// (about the centre of mass) and atomic charges the reference evaluates in its MD callback
// (scripts/MD/Zundel_thermodynamics/continuation/04_Zundel_continuation_MD.py:71-92 dip_moment:
//  mol.intor_symmetric("int1e_r") with mol.with_common_orig(centre of mass), el_dip = sum_ij r_ij dm_ji,
//  mol_dip = sum_A Z_A (R_A - R_com) - el_dip;  :140-159 callback: dm_ao = X gamma X^T with X = get_basis(mol)).
//
//   evc_int1e_r        r[g][x][i][j] = <i| r_x - O_x |j> over contracted Cartesian s / p Gaussians
//                      (Obara-Saika one-dimensional overlaps: <a|x - O|b> = S(l_a + 1, l_b) + (A_x - O_x) S(l_a, l_b))
//   evc_rdm1_observables   one CTA per geometry: P = X gamma X^T, dipole, populations
//                      Mulliken  n_mu = (P S)_mu,mu          Loewdin  n_mu = (S^1/2 P S^1/2)_mu,mu = gamma_mu,mu
//                      charge_A = Z_A - sum_{mu on A} n_mu
#include <vector>

#include "common.cuh"

struct evc_aotable {
  int natm, nao, nprim;
  int32_t *ao_atom, *ao_pow, *ao_poff;
  double *prim_exp, *prim_wt, *charges, *masses;
};

namespace {

// S[i][j], i <= 2, j <= 1: one-dimensional overlaps of (x - A)^i exp(-a (x-A)^2) and (x - B)^j exp(-b (x-B)^2)
// without the factor exp(-a b / p (A - B)^2)
__device__ __forceinline__ void overlap_1d(double p, double PA, double PB, double (&S)[3][2]) {
  const double h = 0.5 / p;
  S[0][0] = sqrt(3.14159265358979323846 / p);
  S[1][0] = PA * S[0][0];
  S[0][1] = PB * S[0][0];
  S[1][1] = PB * S[1][0] + h * S[0][0];
  S[2][0] = PA * S[1][0] + h * S[0][0];
  S[2][1] = PB * S[2][0] + 2.0 * h * S[1][0];
}

__global__ void int1e_r_kernel(int nbatch, int natm, int nao, const int32_t* __restrict__ ao_atom,
                               const int32_t* __restrict__ ao_pow, const int32_t* __restrict__ ao_poff,
                               const double* __restrict__ prim_exp, const double* __restrict__ prim_wt,
                               const double* __restrict__ coords, const double* __restrict__ origin,
                               double* __restrict__ out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(nbatch) * nao * nao;
  if (t >= total) return;
  const int g = static_cast<int>(t / (nao * nao)), ij = static_cast<int>(t - static_cast<int64_t>(g) * nao * nao);
  const int i = ij / nao, j = ij - i * nao;
  const double* R = coords + static_cast<int64_t>(g) * natm * 3;
  const double* A = R + 3 * ao_atom[i];
  const double* B = R + 3 * ao_atom[j];
  const double* O = origin + static_cast<int64_t>(g) * 3;
  const int la[3] = {ao_pow[3 * i], ao_pow[3 * i + 1], ao_pow[3 * i + 2]};
  const int lb[3] = {ao_pow[3 * j], ao_pow[3 * j + 1], ao_pow[3 * j + 2]};
  const double ab2 = (A[0] - B[0]) * (A[0] - B[0]) + (A[1] - B[1]) * (A[1] - B[1]) + (A[2] - B[2]) * (A[2] - B[2]);
  double acc[3] = {0.0, 0.0, 0.0};
  for (int pa = ao_poff[i]; pa < ao_poff[i + 1]; ++pa)
    for (int pb = ao_poff[j]; pb < ao_poff[j + 1]; ++pb) {
      const double a = prim_exp[pa], b = prim_exp[pb], p = a + b;
      const double w = prim_wt[pa] * prim_wt[pb] * exp(-a * b / p * ab2);
      double s0[3], s1[3];  // per dimension: plain overlap, overlap with the operator x - O
      for (int d = 0; d < 3; ++d) {
        const double P = (a * A[d] + b * B[d]) / p;
        double S[3][2];
        overlap_1d(p, P - A[d], P - B[d], S);
        s0[d] = S[la[d]][lb[d]];
        s1[d] = S[la[d] + 1][lb[d]] + (A[d] - O[d]) * s0[d];
      }
      acc[0] += w * s1[0] * s0[1] * s0[2];
      acc[1] += w * s0[0] * s1[1] * s0[2];
      acc[2] += w * s0[0] * s0[1] * s1[2];
    }
  const int64_t o = static_cast<int64_t>(g) * 3 * nao * nao + ij;
  out[o] = acc[0];
  out[o + static_cast<int64_t>(nao) * nao] = acc[1];
  out[o + 2 * static_cast<int64_t>(nao) * nao] = acc[2];
}

// origin[g] = centre of mass
__global__ void com_kernel(int nbatch, int natm, const double* __restrict__ masses, const double* __restrict__ coords,
                           double* __restrict__ origin) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nbatch) return;
  double m = 0.0, c[3] = {0.0, 0.0, 0.0};
  for (int A = 0; A < natm; ++A) {
    m += masses[A];
    for (int d = 0; d < 3; ++d) c[d] += masses[A] * coords[(static_cast<int64_t>(g) * natm + A) * 3 + d];
  }
  for (int d = 0; d < 3; ++d) origin[static_cast<int64_t>(g) * 3 + d] = c[d] / m;
}

// one CTA per geometry
__global__ void __launch_bounds__(128)
rdm1_observables_kernel(int natm, int n, int method, const int32_t* __restrict__ ao_atom,
                        const double* __restrict__ charges, const double* __restrict__ coords,
                        const double* __restrict__ origin, const double* __restrict__ x,
                        const double* __restrict__ gamma, const double* __restrict__ ovlp,
                        const double* __restrict__ rint, double* __restrict__ dm_ao, double* __restrict__ dipole,
                        double* __restrict__ atom_charges) {
  extern __shared__ double sm[];
  const int ld = n + 1;
  double* X = sm;             // [n][ld]
  double* T = X + n * ld;     // gamma, then P
  double* U = T + n * ld;     // X gamma
  double* pop = U + n * ld;   // [n]
  double* red = pop + n;      // [3][4]
  const int g = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int64_t o2 = static_cast<int64_t>(g) * n * n;
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    X[i * ld + j] = x[o2 + k];
    T[i * ld + j] = gamma[o2 + k];
  }
  __syncthreads();
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += X[i * ld + r] * T[r * ld + j];
    U[i * ld + j] = acc;
  }
  __syncthreads();
  if (method == 1)  // Loewdin populations are the diagonal of the OAO density matrix
    for (int k = tid; k < n; k += nt) pop[k] = T[k * ld + k];
  __syncthreads();
  for (int k = tid; k < n * n; k += nt) {  // P = (X gamma) X^T
    const int i = k / n, j = k - i * n;
    double acc = 0.0;
    for (int r = 0; r < n; ++r) acc += U[i * ld + r] * X[j * ld + r];
    T[i * ld + j] = acc;
    if (dm_ao) dm_ao[o2 + k] = acc;
  }
  __syncthreads();
  // electronic dipole: sum_ij r_ij P_ji, fixed summation order (per-thread strided partials, warp tree, warps in order)
  {
    double e[3] = {0.0, 0.0, 0.0};
    const double* r = rint + static_cast<int64_t>(g) * 3 * n * n;
    for (int k = tid; k < n * n; k += nt) {
      const int i = k / n, j = k - i * n;
      const double pji = T[j * ld + i];
      e[0] += r[k] * pji;
      e[1] += r[n * n + k] * pji;
      e[2] += r[2 * n * n + k] * pji;
    }
    for (int d = 0; d < 3; ++d) {
      for (int off = 16; off > 0; off >>= 1) e[d] += __shfl_xor_sync(0xffffffffu, e[d], off);
      if ((tid & 31) == 0) red[d * 4 + (tid >> 5)] = e[d];
    }
  }
  if (method == 0)  // Mulliken: (P S)_mu,mu
    for (int k = tid; k < n; k += nt) {
      double acc = 0.0;
      for (int r = 0; r < n; ++r) acc += T[k * ld + r] * ovlp[o2 + r * n + k];
      pop[k] = acc;
    }
  __syncthreads();
  if (tid < 3) {
    double nuc = 0.0;
    for (int A = 0; A < natm; ++A)
      nuc += charges[A] * (coords[(static_cast<int64_t>(g) * natm + A) * 3 + tid] - origin[static_cast<int64_t>(g) * 3 + tid]);
    double el = 0.0;
    for (int w = 0; w < nt / 32; ++w) el += red[tid * 4 + w];
    dipole[static_cast<int64_t>(g) * 3 + tid] = nuc - el;
  }
  for (int A = tid; A < natm; A += nt) {
    double q = charges[A];
    for (int k = 0; k < n; ++k)
      if (ao_atom[k] == A) q -= pop[k];
    atom_charges[static_cast<int64_t>(g) * natm + A] = q;
  }
}

template <typename T>
int up(T** dst, const T* src, size_t count) {
  EVC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), (count ? count : 1) * sizeof(T)));
  if (count) EVC_CHECK_CUDA(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

extern "C" {

int evc_aotable_create(evc_ctx* ctx, int natm, const double* charges_host, const double* masses_host, int nao,
                       const int32_t* ao_atom_host, const int32_t* ao_pow_host, const int32_t* ao_nprim_host,
                       const double* prim_exp_host, const double* prim_wt_host, evc_aotable** out) {
  EVC_REQUIRE(ctx && charges_host && masses_host && ao_atom_host && ao_nprim_host && prim_exp_host && prim_wt_host && out,
              "evc_aotable_create: NULL argument");
  EVC_REQUIRE(natm >= 1 && nao >= 1 && nao <= 64, "evc_aotable_create: natm=%d nao=%d unsupported", natm, nao);
  EVC_CHECK_CUDA(cudaSetDevice(ctx->device));
  std::vector<int32_t> poff(nao + 1, 0), pw(3 * static_cast<size_t>(nao), 0);
  for (int a = 0; a < nao; ++a) {
    EVC_REQUIRE(ao_nprim_host[a] >= 1, "evc_aotable_create: AO %d has %d primitives", a, ao_nprim_host[a]);
    EVC_REQUIRE(ao_atom_host[a] >= 0 && ao_atom_host[a] < natm, "evc_aotable_create: AO %d sits on atom %d", a,
                ao_atom_host[a]);
    if (ao_pow_host) {
      for (int d = 0; d < 3; ++d) pw[3 * a + d] = ao_pow_host[3 * a + d];
      const int l = pw[3 * a] + pw[3 * a + 1] + pw[3 * a + 2];
      EVC_REQUIRE(pw[3 * a] >= 0 && pw[3 * a + 1] >= 0 && pw[3 * a + 2] >= 0 && l <= 1,
                  "evc_aotable_create: AO %d has angular momentum %d (s and p only)", a, l);
    }
    poff[a + 1] = poff[a] + ao_nprim_host[a];
  }
  for (int A = 0; A < natm; ++A) EVC_REQUIRE(masses_host[A] > 0.0, "evc_aotable_create: mass of atom %d is not positive", A);
  evc_aotable* t = new evc_aotable();
  t->natm = natm;
  t->nao = nao;
  t->nprim = poff[nao];
  int rc = 0;
  rc |= up(&t->ao_atom, ao_atom_host, nao);
  rc |= up(&t->ao_pow, pw.data(), pw.size());
  rc |= up(&t->ao_poff, poff.data(), poff.size());
  rc |= up(&t->prim_exp, prim_exp_host, t->nprim);
  rc |= up(&t->prim_wt, prim_wt_host, t->nprim);
  rc |= up(&t->charges, charges_host, natm);
  rc |= up(&t->masses, masses_host, natm);
  if (rc) return rc;
  *out = t;
  return 0;
}

int evc_aotable_destroy(evc_aotable* t) {
  if (!t) return 0;
  cudaFree(t->ao_atom); cudaFree(t->ao_pow); cudaFree(t->ao_poff);
  cudaFree(t->prim_exp); cudaFree(t->prim_wt); cudaFree(t->charges); cudaFree(t->masses);
  delete t;
  return 0;
}

int evc_center_of_mass(evc_ctx* ctx, const evc_aotable* t, int nbatch, const double* coords, double* origin) {
  EVC_REQUIRE(ctx && t && coords && origin, "evc_center_of_mass: NULL argument");
  if (nbatch <= 0) return 0;
  com_kernel<<<(nbatch + 127) / 128, 128, 0, ctx->stream>>>(nbatch, t->natm, t->masses, coords, origin);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_int1e_r(evc_ctx* ctx, const evc_aotable* t, int nbatch, const double* coords, const double* origin,
                double* r) {
  EVC_REQUIRE(ctx && t && coords && origin && r, "evc_int1e_r: NULL argument");
  if (nbatch <= 0) return 0;
  const int64_t total = static_cast<int64_t>(nbatch) * t->nao * t->nao;
  int1e_r_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, ctx->stream>>>(
      nbatch, t->natm, t->nao, t->ao_atom, t->ao_pow, t->ao_poff, t->prim_exp, t->prim_wt, coords, origin, r);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_rdm1_observables(evc_ctx* ctx, const evc_aotable* t, int nbatch, int method, const double* coords,
                         const double* origin, const double* x, const double* gamma, const double* ovlp,
                         const double* rint, double* dm_ao, double* dipole, double* atom_charges) {
  EVC_REQUIRE(ctx && t && coords && origin && x && gamma && rint && dipole && atom_charges,
              "evc_rdm1_observables: NULL argument");
  EVC_REQUIRE(method == 0 || method == 1, "evc_rdm1_observables: method=%d (0 Mulliken, 1 Loewdin)", method);
  EVC_REQUIRE(method == 1 || ovlp, "evc_rdm1_observables: Mulliken populations need the overlap matrices");
  if (nbatch <= 0) return 0;
  const int n = t->nao;
  const size_t smem = (3 * static_cast<size_t>(n) * (n + 1) + n + 12) * sizeof(double);
  EVC_CHECK_CUDA(cudaFuncSetAttribute(rdm1_observables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  rdm1_observables_kernel<<<nbatch, 128, smem, ctx->stream>>>(t->natm, n, method, t->ao_atom, t->charges, coords, origin,
                                                              x, gamma, ovlp, rint, dm_ao, dipole, atom_charges);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
