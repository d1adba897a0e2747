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

// nuclear repulsion energy and its gradient: one warp per geometry
__global__ void gnuc_kernel(int natm, const double* __restrict__ charges, const double* __restrict__ coords,
                            double* __restrict__ e_nuc, double* __restrict__ grad_nuc) {
  const int g = blockIdx.x, lane = threadIdx.x;
  const double* Rc = coords + static_cast<int64_t>(g) * natm * 3;
  double e = 0.0;
  for (int c0 = 0; c0 < natm; c0 += 32) {
    const int A = c0 + lane;
    double gx = 0, gy = 0, gz = 0;
    if (A < natm) {
      const double za = charges[A];
      for (int B = 0; B < natm; ++B) {
        if (B == A) continue;
        const double dx = Rc[3 * A] - Rc[3 * B], dy = Rc[3 * A + 1] - Rc[3 * B + 1], dz = Rc[3 * A + 2] - Rc[3 * B + 2];
        const double r2 = dx * dx + dy * dy + dz * dz, ri = rsqrt(r2), zz = za * charges[B];
        if (B < A) e += zz * ri;
        const double f = zz * ri * ri * ri;
        gx -= f * dx; gy -= f * dy; gz -= f * dz;
      }
      double* gn = grad_nuc + (static_cast<int64_t>(g) * natm + A) * 3;
      gn[0] = gx; gn[1] = gy; gn[2] = gz;
    }
  }
  e = gwarp_sum(e);
  if (lane == 0) e_nuc[g] = e;
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
  // one-electron work: ordered shell pairs by class (s|s), (p|s), (s|p), (p|p)
  std::vector<int32_t> plist;
  int p_off[5] = {0, 0, 0, 0, 0};
  for (int cls = 0; cls < 4; ++cls) {
    p_off[cls] = static_cast<int>(plist.size());
    for (int A = 0; A < nshell; ++A)
      for (int B = 0; B < nshell; ++B)
        if (sh_l[A] + 2 * sh_l[B] == cls) plist.push_back(A | (B << 8));
  }
  p_off[4] = static_cast<int>(plist.size());
  evc_gbasis* b = new evc_gbasis();
  for (int c = 0; c < 5; ++c) b->p_off[c] = p_off[c];
  b->natm = natm; b->nao = nao; b->nprim = nprim; b->nshell = nshell;
  for (int c = 0; c <= kGClasses; ++c) { b->cq_off[c] = cq_off[c]; b->cunit_off[c] = cunit_off[c]; }
  int rc = 0;
  if ((rc = gupload(&b->ao_atom, ao_atom)) || (rc = gupload(&b->ao_pow, pw)) || (rc = gupload(&b->ao_poff, poff)) ||
      (rc = gupload(&b->aoslices, slices)) || (rc = gupload(&b->prim_exp, ex)) || (rc = gupload(&b->prim_wt, wt)) ||
      (rc = gupload(&b->charges, ch)) || (rc = gupload(&b->boys, boys)) ||
      (rc = gupload(&b->sh_atom, sh_atom)) || (rc = gupload(&b->sh_ao0, sh_ao0)) ||
      (rc = gupload(&b->sh_p0, sh_p0)) || (rc = gupload(&b->sh_np, sh_np)) || (rc = gupload(&b->cq, cq)) ||
      (rc = gupload(&b->cunits, cunits)) || (rc = gupload(&b->plist, plist))) {
    delete b;
    return rc;
  }
  for (int k = 0; k < kGAux; ++k) {
    EVC_CHECK_CUDA(cudaStreamCreateWithFlags(&b->aux[k], cudaStreamNonBlocking));
    EVC_CHECK_CUDA(cudaEventCreateWithFlags(&b->ev_join[k], cudaEventDisableTiming));
  }
  EVC_CHECK_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
  *out = b;
  return 0;
}

int evc_gbasis_destroy(evc_gbasis* b) {
  if (b) {
    for (int k = 0; k < kGAux; ++k) {
      if (b->aux[k]) cudaStreamDestroy(b->aux[k]);
      if (b->ev_join[k]) cudaEventDestroy(b->ev_join[k]);
    }
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    cudaFree(b->ao_atom); cudaFree(b->ao_pow); cudaFree(b->ao_poff); cudaFree(b->aoslices); cudaFree(b->prim_exp);
    cudaFree(b->prim_wt); cudaFree(b->charges); cudaFree(b->boys);
    cudaFree(b->sh_atom); cudaFree(b->sh_ao0); cudaFree(b->sh_p0); cudaFree(b->sh_np); cudaFree(b->cq); cudaFree(b->cunits); cudaFree(b->plist);
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
  // Few geometries: every class kernel is bounded by its longest unit, not by throughput, so the 21 launches run
  // concurrently on side streams (fork / join through events: capturable in a CUDA graph); many geometries: one stream.
  cudaStream_t ring[kGAux + 1];
  ring[0] = ctx->stream;
  int nst = 1;
  if (nbatch <= 16) {
    EVC_CHECK_CUDA(cudaEventRecord(b->ev_fork, ctx->stream));
    for (int k = 0; k < kGAux; ++k) {
      EVC_CHECK_CUDA(cudaStreamWaitEvent(b->aux[k], b->ev_fork, 0));
      ring[1 + k] = b->aux[k];
    }
    nst = kGAux + 1;
  }
  {
    int rc;
    // heaviest kernels first on distinct streams: part 0 = ssss, psss, ppss, psps
    if ((rc = launch_gclass_part0(ring, nst, 0, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_gclass_part1(ring, nst, 4, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_g1e(ring, nst, 7, ctx->sm_count, nbatch, v, b->plist, b->p_off, coords, o))) return rc;
    if ((rc = launch_gclass_part2(ring, nst, 3, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_gclass_part3(ring, nst, 6, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
    if ((rc = launch_gclass_part4(ring, nst, 1, ctx->sm_count, nbatch, v, b->cq, b->cunits, b->cunit_off, coords, o))) return rc;
  }
  gnuc_kernel<<<nbatch, 32, 0, ring[nst - 1]>>>(b->natm, b->charges, coords, e_nuc, grad_nuc);
  EVC_CHECK_LAUNCH();
  if (nst > 1) {
    for (int k = 0; k < kGAux; ++k) {
      EVC_CHECK_CUDA(cudaEventRecord(b->ev_join[k], b->aux[k]));
      EVC_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, b->ev_join[k], 0));
    }
  }
  const int64_t total = static_cast<int64_t>(nbatch) * b->natm * 3 * b->nao * b->nao;
  ghd_sym_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctx->stream>>>(total, b->nao, o.vtmp, hcore_deriv);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
