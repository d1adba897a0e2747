// Closed-shell Fock matrix on the device: what the RHF behind get_basis(mol, "canonical")
// (evcont/electron_integral_utils.py:103-106, scf.RHF(mol).scf() -> mo_coeff; the reference's default
// cibasis, evcont/FCI_EVCont.py:15-21) needs per SCF cycle.
//
//   F = hcore + J - K / 2,   J[p,q] = sum_rs (pq|rs) D[r,s],   K[p,s] = sum_qr (pq|rs) D[q,r]
//
// One CTA per first index p streams the slab eri[p][:][:][:] (n^3 doubles) once: J rows as warp dot
// products over (r,s), K rows as column sums over (q,r) with threads on the contiguous index s.
#include "common.cuh"

namespace {

constexpr int kFockThreads = 256;

__global__ void __launch_bounds__(kFockThreads)
fock_rhf_kernel(int n, const double* __restrict__ hcore, const double* __restrict__ eri, const double* __restrict__ dm,
                double* __restrict__ fock) {
  extern __shared__ double sm[];
  double* D = sm;                 // [n*n]
  double* Kp = D + n * n;         // [kFockThreads / 32][n] partial column sums... sized [nslice][n]
  const int p = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n2 = n * n;
  for (int k = tid; k < n2; k += kFockThreads) D[k] = dm[k];
  __syncthreads();
  const double* slab = eri + static_cast<int64_t>(p) * n2 * n;
  // J[p][q]: one warp per q
  for (int q = warp; q < n; q += kFockThreads / 32) {
    const double* row = slab + static_cast<int64_t>(q) * n2;
    double acc = 0.0;
    for (int k = lane; k < n2; k += 32) acc = fma(row[k], D[k], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) fock[p * n + q] = hcore[p * n + q] + acc;
  }
  // K[p][s] = sum_{(qr)} D[(qr)] slab[(qr)][s]: thread = (slice of (qr), s)
  const int ns = kFockThreads / 64;          // slices (n <= 64)
  const int s = tid & 63, sl = tid >> 6;
  double acc = 0.0;
  if (s < n)
    for (int k = sl; k < n2; k += ns) acc = fma(D[k], slab[static_cast<int64_t>(k) * n + s], acc);
  if (s < n) Kp[sl * n + s] = acc;
  __syncthreads();
  if (tid < n) {
    double k = 0.0;
    for (int j = 0; j < ns; ++j) k += Kp[j * n + tid];
    fock[p * n + tid] -= 0.5 * k;
  }
}

}  // namespace

extern "C" {

int evc_fock_rhf(evc_ctx* ctx, int n, const double* hcore, const double* eri, const double* dm, double* fock) {
  EVC_REQUIRE(ctx && hcore && eri && dm && fock, "evc_fock_rhf: NULL argument");
  EVC_REQUIRE(n >= 1 && n <= 64, "evc_fock_rhf: n=%d unsupported (1..64)", n);
  const size_t smem = (static_cast<size_t>(n) * n + static_cast<size_t>(kFockThreads / 64) * n) * sizeof(double);
  fock_rhf_kernel<<<n, kFockThreads, smem, ctx->stream>>>(n, hcore, eri, dm, fock);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
