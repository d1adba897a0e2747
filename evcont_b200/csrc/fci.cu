// FCI sigma vector and diagonal on the device: what
// cisolver.kernel (pyscf.fci.direct_spin0 / direct_spin1 Davidson) needs from the
// Hamiltonian, called by the reference at evcont/FCI_EVCont.py:70.
//
//   H = sum_pq h'_pq E_pq + 1/2 sum_pqrs (pq|rs) E_pq E_rs,   h'_ps = h_ps - 1/2 sum_q (pq|qs)
//   sigma = H c in three steps over the same link tables as the transition RDMs (K0):
//     D[K,(rs)] = <K|E_rs|c>                     gather through the links of K
//     G[K,(pq)] = sum_rs 1/2 (pq|rs) D[K,(rs)]   ndet x n^2 x n^2 DMMA GEMM (evc_rows_axpy)
//     sigma[J]  = sum_pq h'_pq D[J,(pq)] + sum_{links (a,i,J',s) of J} s G[J',(i a)]
// Link record (8 bytes, string-major table): addr:32 | a:8 | i:8 | sign:8 with
// a^+_a a_i |str> = sign |addr>, hence <str|E_ia|addr> = sign.
#include "common.cuh"

namespace {

__device__ __forceinline__ void unpack_link(uint64_t rec, int& addr, int& a, int& i, double& sign) {
  addr = static_cast<int>(rec & 0xffffffffu);
  a = static_cast<int>((rec >> 32) & 0xff);
  i = static_cast<int>((rec >> 40) & 0xff);
  sign = static_cast<double>(static_cast<signed char>((rec >> 48) & 0xff));
}

// one warp per determinant K = (ia, ib)
__global__ void fci_t1_kernel(int n, int n2p, int64_t na, int64_t nb, const uint64_t* __restrict__ link_a, int nlink_a,
                              const uint64_t* __restrict__ link_b, int nlink_b, const double* __restrict__ c,
                              double* __restrict__ D) {
  const int64_t K = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (K >= na * nb) return;
  const int64_t ia = K / nb, ib = K - ia * nb;
  double* row = D + K * n2p;
  for (int k = lane; k < n2p; k += 32) row[k] = 0.0;
  __syncwarp();
  for (int l = lane; l < nlink_a; l += 32) {
    int addr, a, i;
    double s;
    unpack_link(link_a[ia * nlink_a + l], addr, a, i, s);
    row[i * n + a] += s * c[static_cast<int64_t>(addr) * nb + ib];
  }
  __syncwarp();
  for (int l = lane; l < nlink_b; l += 32) {
    int addr, a, i;
    double s;
    unpack_link(link_b[ib * nlink_b + l], addr, a, i, s);
    row[i * n + a] += s * c[ia * nb + addr];
  }
}

__global__ void fci_sigma_kernel(int n, int n2p, int64_t na, int64_t nb, const uint64_t* __restrict__ link_a,
                                 int nlink_a, const uint64_t* __restrict__ link_b, int nlink_b,
                                 const double* __restrict__ h1eff, const double* __restrict__ D,
                                 const double* __restrict__ G, double* __restrict__ sigma) {
  const int64_t J = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (J >= na * nb) return;
  const int64_t ia = J / nb, ib = J - ia * nb;
  const double* row = D + J * n2p;
  double acc = 0.0;
  for (int k = lane; k < n2p; k += 32) acc = fma(h1eff[k], row[k], acc);
  for (int l = lane; l < nlink_a; l += 32) {
    int addr, a, i;
    double s;
    unpack_link(link_a[ia * nlink_a + l], addr, a, i, s);
    acc = fma(s, G[(static_cast<int64_t>(addr) * nb + ib) * n2p + i * n + a], acc);
  }
  for (int l = lane; l < nlink_b; l += 32) {
    int addr, a, i;
    double s;
    unpack_link(link_b[ib * nlink_b + l], addr, a, i, s);
    acc = fma(s, G[(ia * nb + addr) * n2p + i * n + a], acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) sigma[J] = acc;
}

// <K|H|K> from the occupation strings (pyscf.fci.direct_spin1.make_hdiag)
__global__ void fci_hdiag_kernel(int n, int64_t na, int64_t nb, const int64_t* __restrict__ strs_a,
                                 const int64_t* __restrict__ strs_b, const double* __restrict__ h1,
                                 const double* __restrict__ eri, double* __restrict__ hdiag) {
  const int64_t K = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (K >= na * nb) return;
  const int64_t sa = strs_a[K / nb], sb = strs_b[K % nb];
  const int64_t n2 = static_cast<int64_t>(n) * n, n3 = n2 * n;
  double e = 0.0;
  for (int i = 0; i < n; ++i) {
    const bool ai = (sa >> i) & 1, bi = (sb >> i) & 1;
    if (!ai && !bi) continue;
    e += (ai ? 1.0 : 0.0) * h1[i * n + i] + (bi ? 1.0 : 0.0) * h1[i * n + i];
    for (int j = 0; j < n; ++j) {
      const bool aj = (sa >> j) & 1, bj = (sb >> j) & 1;
      const double Jij = eri[i * n3 + i * n2 + j * n + j], Kij = eri[i * n3 + j * n2 + j * n + i];
      if (ai && aj) e += 0.5 * (Jij - Kij);
      if (bi && bj) e += 0.5 * (Jij - Kij);
      if (ai && bj) e += Jij;
    }
  }
  hdiag[K] = e;
}


// ---- transform_ci: the same state in a rotated one-particle basis -------------------------
// pyscf.fci.addons.transform_ci, called at evcont/FCI_EVCont.py:79-85:
//   ci_new = Ta^T ci Tb,   T[I][J] = det(u[occ(I), occ(J)])   (rows: old strings, columns: new strings)
// One thread per minor: the k x k sub-matrix is gathered from u (shared memory) and eliminated with
// partial pivoting.  transpose != 0 stores T^T (element (r, c) of the output is T[c][r]).
constexpr int kMaxOcc = 12;

__global__ void __launch_bounds__(128)
fci_minors_kernel(int n, int k, int64_t ns, const int64_t* __restrict__ strs, const double* __restrict__ u,
                  int transpose, double* __restrict__ out, int64_t ld) {
  extern __shared__ double us[];
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) us[t] = u[t];
  __syncthreads();
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (c >= ns) return;
  const int64_t sI = strs[transpose ? c : r], sJ = strs[transpose ? r : c];
  unsigned char oi[kMaxOcc], oj[kMaxOcc];
  int ki = 0, kj = 0;
  for (int o = 0; o < n; ++o) {
    if ((sI >> o) & 1) oi[ki++] = static_cast<unsigned char>(o);
    if ((sJ >> o) & 1) oj[kj++] = static_cast<unsigned char>(o);
  }
  double a[kMaxOcc * kMaxOcc];
  for (int x = 0; x < k; ++x)
    for (int y = 0; y < k; ++y) a[x * kMaxOcc + y] = us[oi[x] * n + oj[y]];
  double det = 1.0;
  for (int p = 0; p < k; ++p) {
    int piv = p;
    double best = fabs(a[p * kMaxOcc + p]);
    for (int x = p + 1; x < k; ++x) {
      const double v = fabs(a[x * kMaxOcc + p]);
      if (v > best) { best = v; piv = x; }
    }
    if (best == 0.0) { det = 0.0; break; }
    if (piv != p) {
      for (int y = p; y < k; ++y) {
        const double t = a[p * kMaxOcc + y];
        a[p * kMaxOcc + y] = a[piv * kMaxOcc + y];
        a[piv * kMaxOcc + y] = t;
      }
      det = -det;
    }
    const double d = a[p * kMaxOcc + p], inv = 1.0 / d;
    det *= d;
    for (int x = p + 1; x < k; ++x) {
      const double f = a[x * kMaxOcc + p] * inv;
      for (int y = p + 1; y < k; ++y) a[x * kMaxOcc + y] = fma(-f, a[p * kMaxOcc + y], a[x * kMaxOcc + y]);
    }
  }
  out[r * ld + c] = det;
}

__global__ void fci_zero_pad_kernel(int64_t rows, int64_t cols, int64_t ld, double* __restrict__ m) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int64_t c = cols; c < ld; ++c) m[r * ld + c] = 0.0;
}

}  // namespace

extern "C" {

int evc_fci_hdiag(evc_ctx* ctx, int norb, int64_t na, int64_t nb, const int64_t* strs_a, const int64_t* strs_b,
                  const double* h1, const double* eri, double* hdiag) {
  EVC_REQUIRE(ctx && strs_a && strs_b && h1 && eri && hdiag, "evc_fci_hdiag: NULL argument");
  EVC_REQUIRE(norb >= 1 && norb <= 62, "evc_fci_hdiag: norb=%d unsupported", norb);
  const int64_t nd = na * nb;
  fci_hdiag_kernel<<<static_cast<unsigned>((nd + 127) / 128), 128, 0, ctx->stream>>>(norb, na, nb, strs_a, strs_b, h1,
                                                                                    eri, hdiag);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_fci_contract_workspace_bytes(int norb, int64_t na, int64_t nb, size_t* bytes) {
  EVC_REQUIRE(bytes && norb >= 1 && na >= 1 && nb >= 1, "evc_fci_contract_workspace_bytes: bad arguments");
  const int64_t n2p = (static_cast<int64_t>(norb) * norb + 1) & ~static_cast<int64_t>(1);
  const size_t dg = evc_align_up(static_cast<size_t>(na * nb) * n2p * sizeof(double), 256);
  *bytes = 2 * dg + evc_align_up(evc_rows_axpy_ws_bytes(n2p, static_cast<int>(n2p), static_cast<int>(na * nb)), 256);
  return 0;
}

// h1eff: [n2p] (h' row-major n x n, zero padded); w2: [n2p][n2p] with w2[(rs)][(pq)] = 1/2 (pq|rs)
int evc_fci_contract_2e(evc_ctx* ctx, int norb, int64_t na, int64_t nb, const uint64_t* link_a, int nlink_a,
                        const uint64_t* link_b, int nlink_b, const double* h1eff, const double* w2,
                        const double* civec, double* sigma, void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && link_a && link_b && h1eff && w2 && civec && sigma && workspace, "evc_fci_contract_2e: NULL argument");
  EVC_REQUIRE(na * nb < (int64_t(1) << 31), "evc_fci_contract_2e: %lld determinants unsupported", (long long)(na * nb));
  size_t need = 0;
  int rc = evc_fci_contract_workspace_bytes(norb, na, nb, &need);
  if (rc) return rc;
  EVC_REQUIRE(workspace_bytes >= need, "evc_fci_contract_2e: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  const int n2p = (norb * norb + 1) & ~1;
  const int64_t nd = na * nb;
  evc_arena ar(workspace, workspace_bytes);
  double* D = ar.take<double>(static_cast<size_t>(nd) * n2p);
  double* G = ar.take<double>(static_cast<size_t>(nd) * n2p);
  const size_t gws = evc_rows_axpy_ws_bytes(n2p, n2p, static_cast<int>(nd));
  char* gw = ar.take<char>(gws);
  EVC_REQUIRE(D && G && gw, "evc_fci_contract_2e: workspace too small");
  const unsigned blocks = static_cast<unsigned>((nd + 7) / 8);
  fci_t1_kernel<<<blocks, 256, 0, ctx->stream>>>(norb, n2p, na, nb, link_a, nlink_a, link_b, nlink_b, civec, D);
  EVC_CHECK_LAUNCH();
  if ((rc = evc_rows_axpy(ctx, w2, n2p, n2p, D, static_cast<int>(nd), G, gw, gws))) return rc;
  fci_sigma_kernel<<<blocks, 256, 0, ctx->stream>>>(norb, n2p, na, nb, link_a, nlink_a, link_b, nlink_b, h1eff, D, G,
                                                     sigma);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_transform_ci_workspace_bytes(int norb, int64_t na, int64_t nb, size_t* bytes) {
  EVC_REQUIRE(bytes && norb >= 1 && na >= 1 && nb >= 1, "evc_transform_ci_workspace_bytes: bad arguments");
  const int64_t ldb = (nb + 1) & ~static_cast<int64_t>(1);
  size_t tot = evc_align_up(static_cast<size_t>(na) * na * sizeof(double), 256);        // Ta^T
  tot += evc_align_up(static_cast<size_t>(nb) * ldb * sizeof(double), 256);              // Tb
  tot += 2 * evc_align_up(static_cast<size_t>(na) * ldb * sizeof(double), 256);          // ci Tb, result
  tot += evc_align_up(evc_rows_axpy_ws_bytes(ldb, static_cast<int>(nb), static_cast<int>(na)), 256);
  tot += evc_align_up(evc_rows_axpy_ws_bytes(ldb, static_cast<int>(na), static_cast<int>(na)), 256);
  *bytes = tot + 256;
  return 0;
}

// u: [norb][norb] row-major, rows = old orbitals, columns = new orbitals; strs_a / strs_b: occupation
// strings (device int64, ascending); ci_in / ci_out: [na][nb] contiguous.
int evc_transform_ci(evc_ctx* ctx, int norb, int nelec_a, int nelec_b, int64_t na, int64_t nb, const int64_t* strs_a,
                     const int64_t* strs_b, const double* u, const double* ci_in, double* ci_out, void* workspace,
                     size_t workspace_bytes) {
  EVC_REQUIRE(ctx && strs_a && strs_b && u && ci_in && ci_out && workspace, "evc_transform_ci: NULL argument");
  EVC_REQUIRE(norb >= 1 && norb <= 62 && nelec_a >= 0 && nelec_b >= 0 && nelec_a <= kMaxOcc && nelec_b <= kMaxOcc &&
                  nelec_a <= norb && nelec_b <= norb,
              "evc_transform_ci: norb=%d nelec=(%d,%d) unsupported (at most %d electrons per spin)", norb, nelec_a,
              nelec_b, kMaxOcc);
  EVC_REQUIRE(na < (int64_t(1) << 31) / (nb > 0 ? nb : 1) && na <= 65535 && nb <= 65535,
              "evc_transform_ci: %lld x %lld strings unsupported", (long long)na, (long long)nb);
  size_t need = 0;
  int rc = evc_transform_ci_workspace_bytes(norb, na, nb, &need);
  if (rc) return rc;
  EVC_REQUIRE(workspace_bytes >= need, "evc_transform_ci: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  const int64_t ldb = (nb + 1) & ~static_cast<int64_t>(1);
  evc_arena ar(workspace, workspace_bytes);
  double* TaT = ar.take<double>(static_cast<size_t>(na) * na);
  double* Tb = ar.take<double>(static_cast<size_t>(nb) * ldb);
  double* tmp = ar.take<double>(static_cast<size_t>(na) * ldb);
  double* res = ar.take<double>(static_cast<size_t>(na) * ldb);
  const size_t w1 = evc_rows_axpy_ws_bytes(ldb, static_cast<int>(nb), static_cast<int>(na));
  const size_t w2 = evc_rows_axpy_ws_bytes(ldb, static_cast<int>(na), static_cast<int>(na));
  char* ws1 = ar.take<char>(w1);
  char* ws2 = ar.take<char>(w2);
  EVC_REQUIRE(TaT && Tb && tmp && res && ws1 && ws2, "evc_transform_ci: workspace too small");
  const size_t smem = static_cast<size_t>(norb) * norb * sizeof(double);
  fci_minors_kernel<<<dim3(static_cast<unsigned>((na + 127) / 128), static_cast<unsigned>(na)), 128, smem, ctx->stream>>>(
      norb, nelec_a, na, strs_a, u, 1, TaT, na);
  EVC_CHECK_LAUNCH();
  fci_minors_kernel<<<dim3(static_cast<unsigned>((nb + 127) / 128), static_cast<unsigned>(nb)), 128, smem, ctx->stream>>>(
      norb, nelec_b, nb, strs_b, u, 0, Tb, ldb);
  EVC_CHECK_LAUNCH();
  if (ldb != nb) {
    fci_zero_pad_kernel<<<static_cast<unsigned>((nb + 127) / 128), 128, 0, ctx->stream>>>(nb, nb, ldb, Tb);
    EVC_CHECK_LAUNCH();
  }
  // tmp[Ia][Jb] = sum_Ib ci[Ia][Ib] Tb[Ib][Jb];  res[Ja][Jb] = sum_Ia TaT[Ja][Ia] tmp[Ia][Jb]
  if ((rc = evc_rows_axpy(ctx, Tb, ldb, static_cast<int>(nb), ci_in, static_cast<int>(na), tmp, ws1, w1))) return rc;
  if ((rc = evc_rows_axpy(ctx, tmp, ldb, static_cast<int>(na), TaT, static_cast<int>(na), res, ws2, w2))) return rc;
  EVC_CHECK_CUDA(cudaMemcpy2DAsync(ci_out, static_cast<size_t>(nb) * sizeof(double), res,
                                   static_cast<size_t>(ldb) * sizeof(double), static_cast<size_t>(nb) * sizeof(double),
                                   static_cast<size_t>(na), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

}  // extern "C"
