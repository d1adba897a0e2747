// K8: electronic nuclear gradient in the Loewdin basis, and the fused
// prediction step (K3 -> K4 -> K5 -> K6 -> K7 -> K8).
//
// Replaces get_grad_elec_OAO and callees (evcont/ab_initio_gradients_loewdin.py:
// 13-305) and get_energy_with_grad (:308-379).  The reference forms the
// (n,n,natm,3) derivative of the Loewdin transform and of h1, and contracts a
// six-operand einsum per call; here the derivative never materialises.  With
// S = V diag(s) V^T, X = V s^-1/2 V^T, for xi = (atom A, x):
//
//   grad_xi = sum_{mu in A, nu} -<d_x mu|nu> (Omega + Omega^T)[mu,nu]      (dX/dR terms)
//           + sum_{mu nu} (d h_core / d xi)[mu,nu] (X gamma X^T)[mu,nu]    (d hcore)
//           - 1/2 sum_{m in A} sum_{bcd} (d_x m b|c d) GammaAO_s[m,b,c,d]  (d eri)
//   Omega   = V (G o (V^T Z V)) V^T,  G_pq = -1/(sqrt(s_p) sqrt(s_q)(sqrt(s_p)+sqrt(s_q)))
//   Z       = hcore X (gamma + gamma^T) + 1/2 Y
//   Y[a,i]  = sum_{jkl} g[a,jkl] Gamma_s[i,jkl],  g = three-quarter transformed ERIs
//             (= t3 of the h2 transform, index-reversed), Gamma_s = Gamma + its
//             (1,0,2,3), (3,2,1,0), (2,3,0,1) transposes
//   GammaAO = X Gamma X X X, GammaAO_s = GammaAO + (1,0,3,2) + (2,3,0,1) + (3,2,1,0)
//
// (derivation and the numpy cross-check against the reference: DESIGN.md).
#include "common.cuh"

namespace {

__device__ __forceinline__ double block_reduce_sum(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double tot = 0.0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) tot += scratch[w];
  return tot;
}

// Gsp[i,l,k,j] = Gamma_s[i,j,k,l]
__global__ void gamma_sym_perm_kernel(int n, const double* __restrict__ Gamma, double* __restrict__ Gsp) {
  const int g = blockIdx.y;
  const int64_t n4 = static_cast<int64_t>(n) * n * n * n;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n4) return;
  const int j = static_cast<int>(k % n);
  const int kk = static_cast<int>((k / n) % n);
  const int l = static_cast<int>((k / (static_cast<int64_t>(n) * n)) % n);
  const int i = static_cast<int>(k / (static_cast<int64_t>(n) * n * n));
  const double* G = Gamma + static_cast<int64_t>(g) * n4;
  auto at = [&](int a, int b, int c, int d) { return G[((static_cast<int64_t>(a) * n + b) * n + c) * n + d]; };
  Gsp[static_cast<int64_t>(g) * n4 + k] = at(i, j, kk, l) + at(j, i, kk, l) + at(l, kk, j, i) + at(kk, l, i, j);
}

// Y[g][a][i] = sum_m T3[g][a][m] Gsp[g][i][m],  m < n^3
__global__ void __launch_bounds__(128)
y_contract_kernel(int n, const double* __restrict__ T3, const double* __restrict__ Gsp,
                  double* __restrict__ Y) {
  __shared__ double scratch[4];
  const int g = blockIdx.y;
  const int a = blockIdx.x / n, i = blockIdx.x - a * n;
  const int64_t n3 = static_cast<int64_t>(n) * n * n;
  const double* t = T3 + (static_cast<int64_t>(g) * n + a) * n3;
  const double* s = Gsp + (static_cast<int64_t>(g) * n + i) * n3;
  double acc = 0.0;
  for (int64_t m = threadIdx.x; m < n3; m += 128) acc += t[m] * s[m];
  const double tot = block_reduce_sum(acc, scratch);
  if (threadIdx.x == 0) Y[(static_cast<int64_t>(g) * n + a) * n + i] = tot;
}

// ---- forms for an 8-fold symmetric Gamma (the packed prediction step, n > 13) ----------------------
// Gamma_s = 4 Gamma and GammaAO_s = 4 GammaAO when Gamma has the permutational symmetry of the
// integrals, so the four-fold gathers above collapse to one stream each.

// Gsp[i,l,k,j] = 4 Gamma[i,j,k,l]: per (i, k) a transpose of the (j, l) plane through shared memory
__global__ void __launch_bounds__(256)
gamma_perm4_kernel(int n, const double* __restrict__ Gamma, double* __restrict__ Gsp) {
  extern __shared__ double tile[];  // [n][n + 1]
  const int g = blockIdx.y, i = blockIdx.x / n, k = blockIdx.x - i * n;
  const int64_t n2 = static_cast<int64_t>(n) * n, n4 = n2 * n2;
  const double* G = Gamma + static_cast<int64_t>(g) * n4 + static_cast<int64_t>(i) * n2 * n + static_cast<int64_t>(k) * n;
  double* O = Gsp + static_cast<int64_t>(g) * n4 + static_cast<int64_t>(i) * n2 * n + static_cast<int64_t>(k) * n;
  for (int t = threadIdx.x; t < n * n; t += 256) {
    const int j = t / n, l = t - j * n;
    tile[j * (n + 1) + l] = G[j * n2 + l];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n * n; t += 256) {
    const int l = t / n, j = t - l * n;
    O[l * n2 + j] = 4.0 * tile[j * (n + 1) + l];
  }
}

// Y[g][a][i] = sum_m T3[g][a][m] Gsp[g][i][m], m < n^3, as a (n x n^3)(n^3 x n) product on the FP64
// tensor cores: CTA = (chunk of m, geometry); every warp owns the m = 4-element steps
// warp, warp + 8, ... of the chunk for all (n/8)^2 output tiles, operands straight from global memory
// (each element is read once); the eight warps are combined in a fixed order.  n <= 32.
__global__ void __launch_bounds__(256)
y_contract_mma_kernel(int n, int nchunk, const double* __restrict__ T3, const double* __restrict__ Gsp,
                      double* __restrict__ part) {
  extern __shared__ double red_dyn[];  // [8][32 * 33]
  double (*red)[32 * 33] = reinterpret_cast<double (*)[32 * 33]>(red_dyn);
  const int g = blockIdx.y, ch = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tg = lane & 3;
  const int64_t n3 = static_cast<int64_t>(n) * n * n;
  const int64_t steps = (n3 + 3) / 4;
  const int64_t s0 = steps * ch / nchunk, s1 = steps * (ch + 1) / nchunk;
  const double* A = T3 + static_cast<int64_t>(g) * n * n3;
  const double* B = Gsp + static_cast<int64_t>(g) * n * n3;
  const int nt8 = (n + 7) >> 3;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  // two steps per iteration: 16 independent 8-byte loads in flight per lane before the 32 DMMAs
  for (int64_t s = s0 + warp; s < s1; s += 16) {
    double af[2][4], bf[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t su = s + 8 * u;
      const int64_t m = 4 * su + tg;
      const bool mok = su < s1 && m < n3;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int r = t * 8 + gq;
        const bool ok = mok && r < n && t < nt8;
        af[u][t] = ok ? __ldg(A + static_cast<int64_t>(r) * n3 + m) : 0.0;
        bf[u][t] = ok ? __ldg(B + static_cast<int64_t>(r) * n3 + m) : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (a < nt8 && b < nt8) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[u][a], bf[u][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      red[warp][(a * 8 + gq) * 33 + b * 8 + tg * 2] = acc[a][b][0];
      red[warp][(a * 8 + gq) * 33 + b * 8 + tg * 2 + 1] = acc[a][b][1];
    }
  __syncthreads();
  double* dst = part + (static_cast<int64_t>(g) * nchunk + ch) * n * n;
  for (int t = threadIdx.x; t < n * n; t += 256) {
    const int a = t / n, i = t - a * n;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][a * 33 + i];
    dst[t] = v;
  }
}

__global__ void y_reduce_kernel(int n2, int nchunk, const double* __restrict__ part, double* __restrict__ Y) {
  const int g = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n2) return;
  const double* p = part + static_cast<int64_t>(g) * nchunk * n2 + t;
  double v = 0.0;
  for (int c = 0; c < nchunk; ++c) v += p[static_cast<int64_t>(c) * n2];
  Y[static_cast<int64_t>(g) * n2 + t] = v;
}

// per geometry: OmS = Omega + Omega^T, Pao = X gamma X^T
__global__ void one_el_adjoint_kernel(int n, const double* __restrict__ evals,
                                      const double* __restrict__ evecs, const double* __restrict__ x,
                                      const double* __restrict__ hcore, const double* __restrict__ gamma,
                                      const double* __restrict__ Y, double* __restrict__ OmS,
                                      double* __restrict__ Pao) {
  extern __shared__ double sm[];
  const int ld = n + 1;
  double* V = sm;
  double* X = V + n * ld;
  double* Hc = X + n * ld;
  double* Gm = Hc + n * ld;
  double* Z = Gm + n * ld;
  double* A = Z + n * ld;
  double* B = A + n * ld;
  double* rs = B + n * ld;  // sqrt(s), 0 if cut
  const int g = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int64_t o = static_cast<int64_t>(g) * n * n;
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    V[i * ld + j] = evecs[o + k];
    X[i * ld + j] = x[o + k];
    Hc[i * ld + j] = hcore[o + k];
    Gm[i * ld + j] = gamma[o + k];
    Z[i * ld + j] = 0.5 * Y[o + k];
  }
  for (int k = tid; k < n; k += nt) {
    const double s = evals[static_cast<int64_t>(g) * n + k];
    rs[k] = s > 1.0e-15 ? sqrt(s) : 0.0;
  }
  __syncthreads();
  auto matmul = [&](double* C, const double* P, bool tp, const double* Q, bool tq) {
    for (int k = tid; k < n * n; k += nt) {
      const int i = k / n, j = k - i * n;
      double acc = 0.0;
      for (int r = 0; r < n; ++r)
        acc += (tp ? P[r * ld + i] : P[i * ld + r]) * (tq ? Q[j * ld + r] : Q[r * ld + j]);
      C[i * ld + j] = acc;
    }
    __syncthreads();
  };
  // B = gamma + gamma^T
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    B[i * ld + j] = Gm[i * ld + j] + Gm[j * ld + i];
  }
  __syncthreads();
  matmul(A, X, false, B, false);   // A = X (gamma + gamma^T)
  matmul(B, Hc, false, A, false);  // B = hcore A = Q
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    Z[i * ld + j] += B[i * ld + j];
  }
  __syncthreads();
  matmul(A, V, true, Z, false);    // A = V^T Z
  matmul(B, A, false, V, false);   // B = V^T Z V
  for (int k = tid; k < n * n; k += nt) {
    const int p = k / n, q = k - p * n;
    const double rp = rs[p], rq = rs[q];
    double gpq = 0.0;
    if (rp > 0.0 && rq > 0.0) {
      gpq = -1.0 / (rp * rq * (rp + rq));
    } else if ((rp > 0.0) != (rq > 0.0)) {
      const double sp = evals[static_cast<int64_t>(g) * n + p], sq = evals[static_cast<int64_t>(g) * n + q];
      if (sp != sq) gpq = ((rp > 0.0 ? 1.0 / rp : 0.0) - (rq > 0.0 ? 1.0 / rq : 0.0)) / (sp - sq);
    }
    B[p * ld + q] *= gpq;
  }
  __syncthreads();
  matmul(A, V, false, B, false);   // A = V B
  matmul(Z, A, false, V, true);    // Z = Omega = V B V^T
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    OmS[o + k] = Z[i * ld + j] + Z[j * ld + i];
  }
  matmul(A, X, false, Gm, false);  // A = X gamma
  matmul(B, A, false, X, true);    // B = X gamma X^T
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k - i * n;
    Pao[o + k] = B[i * ld + j];
  }
}

// T2[g][x][m] = sum_{bcd} ip1[g][x][m][bcd] * GAO_s[m][bcd]
__global__ void __launch_bounds__(256)
ip1_dot_kernel(int n, int sym8, const double* __restrict__ ip1, const double* __restrict__ GAO,
               double* __restrict__ T2) {
  __shared__ double scratch[8];
  const int g = blockIdx.y, m = blockIdx.x;
  const int64_t n3 = static_cast<int64_t>(n) * n * n, n4 = n3 * n;
  const double* G = GAO + static_cast<int64_t>(g) * n4;
  const double* ip = ip1 + static_cast<int64_t>(g) * 3 * n4 + static_cast<int64_t>(m) * n3;
  auto at = [&](int a, int b, int c, int d) { return G[((static_cast<int64_t>(a) * n + b) * n + c) * n + d]; };
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (sym8) {  // GammaAO_s = 4 GammaAO: a pure stream, 8 loads in flight per thread
    const double* Gm = G + static_cast<int64_t>(m) * n3;
#pragma unroll 2
    for (int64_t k = threadIdx.x; k < n3; k += 256) {
      const double gs = 4.0 * __ldg(Gm + k);
      a0 = fma(__ldg(ip + k), gs, a0);
      a1 = fma(__ldg(ip + n4 + k), gs, a1);
      a2 = fma(__ldg(ip + 2 * n4 + k), gs, a2);
    }
  } else
  for (int64_t k = threadIdx.x; k < n3; k += 256) {
    const int d = static_cast<int>(k % n);
    const int c = static_cast<int>((k / n) % n);
    const int b = static_cast<int>(k / (static_cast<int64_t>(n) * n));
    const double gs = G[static_cast<int64_t>(m) * n3 + k] + at(b, m, d, c) + at(c, d, m, b) + at(d, c, b, m);
    a0 += ip[k] * gs;
    a1 += ip[n4 + k] * gs;
    a2 += ip[2 * n4 + k] * gs;
  }
  a0 = block_reduce_sum(a0, scratch);
  a1 = block_reduce_sum(a1, scratch);
  a2 = block_reduce_sum(a2, scratch);
  if (threadIdx.x == 0) {
    double* t = T2 + static_cast<int64_t>(g) * 3 * n;
    t[m] = a0; t[n + m] = a1; t[2 * n + m] = a2;
  }
}

// grad[g][A][x]
__global__ void __launch_bounds__(128)
grad_final_kernel(int n, int natm, const int32_t* __restrict__ aoslices, const double* __restrict__ ipovlp,
                  const double* __restrict__ hcore_deriv, const double* __restrict__ OmS,
                  const double* __restrict__ Pao, const double* __restrict__ T2,
                  const double* __restrict__ grad_nuc, double* __restrict__ grad) {
  __shared__ double scratch[4];
  const int g = blockIdx.y, A = blockIdx.x;
  const int p0 = aoslices[2 * A], p1 = aoslices[2 * A + 1];
  const int n2 = n * n;
  const double* om = OmS + static_cast<int64_t>(g) * n2;
  const double* pa = Pao + static_cast<int64_t>(g) * n2;
  const double* ip = ipovlp + static_cast<int64_t>(g) * 3 * n2;
  const double* hd = hcore_deriv + (static_cast<int64_t>(g) * natm + A) * 3 * n2;
  const double* t2 = T2 + static_cast<int64_t>(g) * 3 * n;
  for (int x = 0; x < 3; ++x) {
    double acc = 0.0;
    const int cnt = (p1 - p0) * n;
    for (int k = threadIdx.x; k < cnt; k += 128) {
      const int idx = p0 * n + k;
      acc -= ip[x * n2 + idx] * om[idx];
    }
    for (int k = threadIdx.x; k < n2; k += 128) acc += hd[x * n2 + k] * pa[k];
    for (int m = p0 + threadIdx.x; m < p1; m += 128) acc -= 0.5 * t2[x * n + m];
    const double tot = block_reduce_sum(acc, scratch);
    if (threadIdx.x == 0) {
      const int64_t o = (static_cast<int64_t>(g) * natm + A) * 3 + x;
      grad[o] = tot + (grad_nuc ? grad_nuc[o] : 0.0);
    }
  }
}

__global__ void add_enuc_kernel(int G, const double* __restrict__ e0, const double* __restrict__ e_nuc,
                                double* __restrict__ E) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) E[g] = e0[g] + (e_nuc ? e_nuc[g] : 0.0);
}

// chunks of the n^3 contraction index per geometry in y_contract_mma_kernel: about two CTAs per SM
// in all, at least 512 elements per chunk (independent of the device: it sizes the workspace)
inline int y_nchunk(int n, int nbatch) {
  const long long n3 = static_cast<long long>(n) * n * n;
  long long c = (2 * 148 + nbatch - 1) / nbatch;
  const long long cap = n3 / 512 > 0 ? n3 / 512 : 1;
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  return static_cast<int>(c);
}

int grad_elec_impl(evc_ctx* ctx, int nbatch, int n, int natm, const int32_t* aoslices,
                   const double* evals, const double* evecs, const double* x, const double* hcore,
                   const double* t3, const double* gamma, const double* Gamma, const double* ipovlp,
                   const double* hcore_deriv, const double* eri_ip1, const double* grad_nuc,
                   double* grad, void* workspace, size_t workspace_bytes, int sym8) {
  const size_t n2 = static_cast<size_t>(n) * n, n4 = n2 * n2;
  evc_arena ar(workspace, workspace_bytes);
  double* bufA = ar.take<double>(nbatch * n4);  // Gsp, then rot scratch
  double* bufB = ar.take<double>(nbatch * n4);  // rot scratch
  double* GAO = ar.take<double>(nbatch * n4);
  double* Y = ar.take<double>(nbatch * n2);
  double* OmS = ar.take<double>(nbatch * n2);
  double* Pao = ar.take<double>(nbatch * n2);
  double* T2 = ar.take<double>(static_cast<size_t>(nbatch) * 3 * n);
  EVC_REQUIRE(bufA && bufB && GAO && Y && OmS && Pao && T2, "evc_grad_elec: workspace too small (%zu bytes)",
              workspace_bytes);
  cudaStream_t st = ctx->stream;
  if (sym8 && n <= 32) {
    // Gamma is 8-fold symmetric (packed step): one transposing stream, Y on the tensor cores
    const int nchunk = y_nchunk(n, nbatch);
    double* ypart = ar.take<double>(static_cast<size_t>(nbatch) * nchunk * n2);
    EVC_REQUIRE(ypart, "evc_grad_elec: workspace too small (%zu bytes)", workspace_bytes);
    dim3 grid(n * n, nbatch);
    gamma_perm4_kernel<<<grid, 256, static_cast<size_t>(n) * (n + 1) * sizeof(double), st>>>(n, Gamma, bufA);
    EVC_CHECK_LAUNCH();
    dim3 g2(nchunk, nbatch);
    constexpr size_t kYSmem = 8 * 32 * 33 * sizeof(double);
    EVC_CHECK_CUDA(cudaFuncSetAttribute(y_contract_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kYSmem)));
    y_contract_mma_kernel<<<g2, 256, kYSmem, st>>>(n, nchunk, t3, bufA, ypart);
    EVC_CHECK_LAUNCH();
    dim3 g3(static_cast<unsigned>((n2 + 127) / 128), nbatch);
    y_reduce_kernel<<<g3, 128, 0, st>>>(static_cast<int>(n2), nchunk, ypart, Y);
    EVC_CHECK_LAUNCH();
  } else {
    dim3 grid(static_cast<unsigned>((n4 + 255) / 256), nbatch);
    gamma_sym_perm_kernel<<<grid, 256, 0, st>>>(n, Gamma, bufA);
    EVC_CHECK_LAUNCH();
    dim3 g2(n * n, nbatch);
    y_contract_kernel<<<g2, 128, 0, st>>>(n, t3, bufA, Y);
    EVC_CHECK_LAUNCH();
  }
  {
    const size_t smem = (7 * n * (n + 1) + n) * sizeof(double);
    EVC_CHECK_CUDA(cudaFuncSetAttribute(one_el_adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    one_el_adjoint_kernel<<<nbatch, n <= 16 ? 128 : 256, smem, st>>>(n, evals, evecs, x, hcore, gamma, Y, OmS, Pao);
    EVC_CHECK_LAUNCH();
  }
  int rc;
  // GammaAO[abcd] = sum Gamma[ijkl] X[a,i] X[b,j] X[c,k] X[d,l]: four passes with X^T
  if ((rc = evc_launch_rot_pass(st, nbatch, n, Gamma, x, 1, bufA))) return rc;
  if ((rc = evc_launch_rot_pass(st, nbatch, n, bufA, x, 1, bufB))) return rc;
  if ((rc = evc_launch_rot_pass(st, nbatch, n, bufB, x, 1, bufA))) return rc;
  if ((rc = evc_launch_rot_pass(st, nbatch, n, bufA, x, 1, GAO))) return rc;
  {
    dim3 grid(n, nbatch);
    ip1_dot_kernel<<<grid, 256, 0, st>>>(n, sym8 ? 1 : 0, eri_ip1, GAO, T2);
    EVC_CHECK_LAUNCH();
  }
  {
    dim3 grid(natm, nbatch);
    grad_final_kernel<<<grid, 128, 0, st>>>(n, natm, aoslices, ipovlp, hcore_deriv, OmS, Pao, T2, grad_nuc, grad);
    EVC_CHECK_LAUNCH();
  }
  return 0;
}

size_t grad_ws_bytes(int n, int natm, int nbatch) {
  const size_t n2 = static_cast<size_t>(n) * n, n4 = n2 * n2;
  (void)natm;
  return 3 * evc_align_up(nbatch * n4 * 8, 256) + 3 * evc_align_up(nbatch * n2 * 8, 256) +
         evc_align_up(static_cast<size_t>(nbatch) * 3 * n * 8, 256) +
         evc_align_up(static_cast<size_t>(nbatch) * y_nchunk(n, nbatch) * n2 * 8, 256);
}

}  // namespace

int evc_grad_elec_full(evc_ctx* ctx, int nbatch, int n, int natm, const int32_t* aoslices,
                       const double* evals, const double* evecs, const double* x, const double* hcore,
                       const double* t3, const double* gamma, const double* Gamma, const double* ipovlp,
                       const double* hcore_deriv, const double* eri_ip1, const double* grad_nuc,
                       double* grad, void* workspace, size_t workspace_bytes, int sym8) {
  return grad_elec_impl(ctx, nbatch, n, natm, aoslices, evals, evecs, x, hcore, t3, gamma, Gamma, ipovlp,
                        hcore_deriv, eri_ip1, grad_nuc, grad, workspace, workspace_bytes, sym8);
}

extern "C" {

int evc_grad_workspace_bytes(int n, int natm, int nbatch, size_t* bytes) {
  EVC_REQUIRE(bytes != nullptr, "evc_grad_workspace_bytes: bytes is NULL");
  *bytes = grad_ws_bytes(n, natm, nbatch);
  return 0;
}

int evc_grad_elec(evc_ctx* ctx, int nbatch, int n, int natm, const int32_t* aoslices,
                  const double* evals, const double* evecs, const double* x, const double* hcore,
                  const double* t3, const double* gamma, const double* Gamma, const double* ipovlp,
                  const double* hcore_deriv, const double* eri_ip1, double* grad_elec,
                  void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && aoslices && evals && evecs && x && hcore && t3 && gamma && Gamma && ipovlp &&
                  hcore_deriv && eri_ip1 && grad_elec && workspace,
              "evc_grad_elec: NULL argument");
  EVC_REQUIRE(n >= 1 && n <= 32 && natm >= 1, "evc_grad_elec: n=%d natm=%d unsupported", n, natm);
  if (nbatch <= 0) return 0;
  return grad_elec_impl(ctx, nbatch, n, natm, aoslices, evals, evecs, x, hcore, t3, gamma, Gamma, ipovlp,
                        hcore_deriv, eri_ip1, nullptr, grad_elec, workspace, workspace_bytes, 0);
}

int evc_energy_with_grad_workspace_bytes(int layout, int N, int n, int natm, int nbatch, size_t* bytes) {
  EVC_REQUIRE(bytes != nullptr, "evc_energy_with_grad_workspace_bytes: bytes is NULL");
  const size_t n2 = static_cast<size_t>(n) * n, n4 = n2 * n2;
  size_t sub = 0, pred = 0;
  int rc;
  if ((rc = evc_subspace_workspace_bytes(layout, N, n, nbatch, &sub))) return rc;
  if ((rc = evc_predict_workspace_bytes(layout, N, n, nbatch, &pred))) return rc;
  size_t tot = 0;
  tot += 4 * evc_align_up(nbatch * n2 * 8, 256);                 // X, evecs, h1, gamma
  tot += evc_align_up(static_cast<size_t>(nbatch) * n * 8, 256);  // evals
  tot += 4 * evc_align_up(nbatch * n4 * 8, 256);                 // h2, t3, rot scratch, Gamma
  tot += evc_align_up(static_cast<size_t>(nbatch) * N * N * 8, 256);  // H
  tot += evc_align_up(static_cast<size_t>(nbatch) * 8, 256);          // E0
  tot += evc_align_up(static_cast<size_t>(nbatch) * N * 8, 256);      // C
  tot += sub + pred + grad_ws_bytes(n, natm, nbatch);
  *bytes = tot;
  return 0;
}

int evc_energy_with_grad(evc_ctx* ctx, int layout, int N, int n, int natm, const double* one_rdm,
                         const double* two_rdm, const double* Linv, int nbatch, const evc_ao_bundle* ao,
                         double* E, double* grad, double* gamma_out, double* Gamma_out, double* Cvec,
                         void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && one_rdm && two_rdm && Linv && ao && E && grad && workspace,
              "evc_energy_with_grad: NULL argument");
  EVC_REQUIRE(ao->ovlp && ao->hcore && ao->eri && ao->ipovlp && ao->hcore_deriv && ao->eri_ip1 && ao->aoslices,
              "evc_energy_with_grad: incomplete AO bundle");
  EVC_REQUIRE(n >= 1 && n <= 32, "evc_energy_with_grad: n=%d unsupported (1..32)", n);
  if (nbatch <= 0) return 0;
  const size_t n2 = static_cast<size_t>(n) * n, n4 = n2 * n2;
  size_t sub_b = 0, pred_b = 0;
  int rc;
  if ((rc = evc_subspace_workspace_bytes(layout, N, n, nbatch, &sub_b))) return rc;
  if ((rc = evc_predict_workspace_bytes(layout, N, n, nbatch, &pred_b))) return rc;
  const size_t grad_b = grad_ws_bytes(n, natm, nbatch);
  evc_arena ar(workspace, workspace_bytes);
  double* X = ar.take<double>(nbatch * n2);
  double* evecs = ar.take<double>(nbatch * n2);
  double* h1 = ar.take<double>(nbatch * n2);
  double* gamma = ar.take<double>(nbatch * n2);
  double* evals = ar.take<double>(static_cast<size_t>(nbatch) * n);
  double* h2 = ar.take<double>(nbatch * n4);
  double* t3 = ar.take<double>(nbatch * n4);
  double* scratch = ar.take<double>(nbatch * n4);
  double* Gamma = ar.take<double>(nbatch * n4);
  double* H = ar.take<double>(static_cast<size_t>(nbatch) * N * N);
  double* E0 = ar.take<double>(nbatch);
  double* C = ar.take<double>(static_cast<size_t>(nbatch) * N);
  char* sub_ws = ar.take<char>(sub_b);
  char* pred_ws = ar.take<char>(pred_b);
  char* grad_ws = ar.take<char>(grad_b);
  EVC_REQUIRE(X && evecs && h1 && gamma && evals && h2 && t3 && scratch && Gamma && H && E0 && C &&
                  (sub_ws || sub_b == 0) && (pred_ws || pred_b == 0) && grad_ws,
              "evc_energy_with_grad: workspace too small (%zu bytes)", workspace_bytes);
  if (gamma_out) gamma = gamma_out;
  if (Gamma_out) Gamma = Gamma_out;
  if (Cvec) C = Cvec;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_LOEWDIN))) return rc;
  if ((rc = evc_loewdin(ctx, nbatch, n, ao->ovlp, X, evals, evecs))) return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_AO2OAO))) return rc;
  if ((rc = evc_ao2oao(ctx, nbatch, n, ao->hcore, ao->eri, X, 0, h1, h2, t3, scratch, evc_align_up(nbatch * n4 * 8, 256)))) return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_SUBSPACE_H))) return rc;
  if ((rc = evc_subspace_H(ctx, layout, N, n, one_rdm, two_rdm, nbatch, h1, h2, H, sub_ws, sub_b))) return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_GENEIG))) return rc;
  if ((rc = evc_geneig(ctx, nbatch, N, H, Linv, 1, E0, C))) return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_PREDICT))) return rc;
  if ((rc = evc_predict_rdm(ctx, layout, N, n, one_rdm, two_rdm, nbatch, C, N, gamma, Gamma, pred_ws, pred_b))) return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_GRAD))) return rc;
  if ((rc = grad_elec_impl(ctx, nbatch, n, natm, ao->aoslices, evals, evecs, X, ao->hcore, t3, gamma, Gamma,
                           ao->ipovlp, ao->hcore_deriv, ao->eri_ip1, ao->grad_nuc, grad, grad_ws, grad_b, 0)))
    return rc;
  if ((rc = evc_stage_mark(ctx, EVC_STAGE_GRAD_STREAM))) return rc;
  add_enuc_kernel<<<(nbatch + 127) / 128, 128, 0, ctx->stream>>>(nbatch, E0, ao->e_nuc, E);
  EVC_CHECK_LAUNCH();
  return evc_stage_mark(ctx, EVC_NSTAGE);
}

}  // extern "C"
