// K3 for large batches of small overlap matrices (n <= 16): the Loewdin transformation with the whole
// Jacobi eigensolver in registers.
//
// Reference anchor: evcont/electron_integral_utils.py:6-18 (get_loewdin_trafo: eigh of the overlap matrix,
// X = V s^-1/2 V^T over the eigenvalues above 1e-15).
//
// The shared-memory kernel of dense.cu is instruction bound at large batches (21.7 k warp instructions per
// 10 x 10 problem, most of them index arithmetic, shared-memory traffic and barriers around ~10 % floating
// point).  Here lane i of a GROUP of n lanes owns row i of A and row i of V in registers, 32 / n matrices
// share a warp (three 10 x 10 problems), the tournament schedule is unrolled at compile time so that every
// register index is a constant, and the only communication is warp shuffles:
//   per round (n/2 disjoint rotations):   one rotation angle per lane (the pair's two lanes compute the same
//   numbers from the same inputs), row update through the partner lane's row (n double shuffles), column
//   update of A and V with the n/2 broadcast (c, s) pairs.
// About 4.8 k warp instructions per 10 x 10 problem.
//
// Each lane's own diagonal element lives in a scalar (`dii`, updated by the closed form app -+ t apq); its slot
// in the row array (whose register index would depend on the lane) is kept at zero, and so is the element the
// rotation annihilates -- both are written where their indices are static, in the column update of the lane's
// own pair.
#include <cstdlib>

#include "common.cuh"

namespace {

template <int NP>
__host__ __device__ constexpr int sched_p(int r, int k) {  // circle method, player NP-1 fixed
  return k == 0 ? r : (r + k) % (NP - 1);
}
template <int NP>
__host__ __device__ constexpr int sched_q(int r, int k) {
  return k == 0 ? NP - 1 : (r - k + 2 * (NP - 1)) % (NP - 1);
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <int N>
struct LwGeom {
  static constexpr int NP = N + (N & 1);  // players of the tournament (odd N: player N is the bye)
  static constexpr int MPW = 32 / N;      // matrices per warp
  static constexpr int kWarps = 4;        // warps per CTA
  static constexpr int vs = N * (N + 1);  // doubles of one sorted eigenvector matrix in shared memory
  static constexpr size_t smem = static_cast<size_t>(kWarps) * MPW * (vs + N) * sizeof(double);
};

template <int N>
__global__ void __launch_bounds__(LwGeom<N>::kWarps * 32)
loewdin_reg_kernel(int nbatch, const double* __restrict__ s_ao, double* __restrict__ x,
                   double* __restrict__ evals, double* __restrict__ evecs) {
  using G = LwGeom<N>;
  constexpr int NP = G::NP, MPW = G::MPW, HALF = NP / 2;
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // lanes beyond the last group shadow lane 0 (same shuffles, same numbers, no stores)
  const bool spare = lane >= MPW * N;
  const int grp = spare ? 0 : lane / N, idx = spare ? 0 : lane - grp * N, base = grp * N;
  const int64_t b0 = (static_cast<int64_t>(blockIdx.x) * G::kWarps + warp) * MPW;
  if (b0 >= nbatch) return;  // whole warp
  const int64_t b = b0 + grp;
  const bool valid = !spare && b < nbatch;
  const int64_t bl = b < nbatch ? b : nbatch - 1;  // a finite matrix for the lanes without one

  double a[N], v[N], dii = 0.0;
  {
    // numpy.linalg.eigh reads the lower triangle only
    const double* S = s_ao + bl * (N * N);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double lo = __ldg(S + idx * N + j), up = __ldg(S + j * N + idx);
      a[j] = idx == j ? 0.0 : (idx > j ? lo : up);  // the diagonal element lives in dii
      v[j] = idx == j ? 1.0 : 0.0;
      if (idx == j) dii = lo;
    }
  }

  const double tol = (static_cast<double>(N) * 2.3e-16) * (static_cast<double>(N) * 2.3e-16);
  bool active = true;
  for (int sweep = 0; sweep < 30; ++sweep) {
    // squared off-diagonal norm against the squared Frobenius norm, summed in lane order by every lane of
    // the group (one decision per matrix)
    {
      double offr = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) offr = fma(a[j], a[j], offr);
      const double totr = fma(dii, dii, offr);
      double off = 0.0, tot = 0.0;
#pragma unroll
      for (int l = 0; l < N; ++l) {
        off += shfl_d(offr, base + l);
        tot += shfl_d(totr, base + l);
      }
      if (off <= tol * tot) active = false;
    }
    if (!__any_sync(0xffffffffu, active)) break;

#pragma unroll
    for (int r = 0; r < NP - 1; ++r) {
      // this lane's partner in round r
      int partner;
      if (idx == NP - 1) partner = r;
      else if (idx == r) partner = NP - 1;
      else {
        partner = 2 * r - idx + (NP - 1);
        partner -= partner >= 2 * (NP - 1) ? 2 * (NP - 1) : 0;
        partner -= partner >= (NP - 1) ? (NP - 1) : 0;
      }
      const bool bye = partner >= N;  // odd N only
      const int plane = base + (bye ? idx : partner);
      const bool is_p = idx < partner;
      // rotation angle: (app, aqq, apq) with apq taken from the p lane by both lanes
      double arow = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) arow = j == partner ? a[j] : arow;
      const double dother = shfl_d(dii, plane), aother = shfl_d(arow, plane);
      const double app = is_p ? dii : dother, aqq = is_p ? dother : dii, apq = is_p ? arow : aother;
      // Rotation angle without divisions (the dependent chain of a round is what a warp waits for): with d = aqq - app,
      // b = 2 apq and theta = d / b the classical t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) is tan(phi) with
      // cos(2 phi) = |d| / r, sin(2 phi) = |b| / r, r = sqrt(d^2 + b^2):  c = sqrt((1 + cos 2phi) / 2),
      // |s| = sin(2 phi) / (2 c) -- two reciprocal square roots and a few products.
      double c = 1.0, s = 0.0, dnew = dii;
      {
        const double d = aqq - app, bq = 2.0 * apq;
        const double r2 = fma(d, d, bq * bq);
        if (active && !bye && fabs(apq) > 1.0e-300 && r2 > 1.0e-290) {
          const double rinv = rsqrt(r2);
          const double h = fma(0.5 * fabs(d), rinv, 0.5);   // (1 + cos 2phi) / 2 = c^2, in [1/2, 1]
          const double cinv = rsqrt(h);
          c = h * cinv;
          const double sabs = 0.5 * fabs(bq) * rinv * cinv;
          s = ((d >= 0.0) == (bq >= 0.0)) ? sabs : -sabs;   // sgn(theta), theta = 0 counts as positive
          const double t = s * cinv;
          dnew = is_p ? app - t * apq : aqq + t * apq;
        }
      }
      dii = dnew;
      // rows: row p <- c row p - s row q, row q <- s row p + c row q; for the q lane that is
      // row <- c row + s' other with s' = +s, for the p lane s' = -s
      {
        const double sp = is_p ? -s : s;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const double o = shfl_d(a[j], plane);
          a[j] = fma(sp, o, c * a[j]);
        }
      }
      // columns of A and of V, pair by pair (static indices); the lane's own pair leaves an exact zero
#pragma unroll
      for (int k = 0; k < HALF; ++k) {
        const int p0 = sched_p<NP>(r, k), q0 = sched_q<NP>(r, k);
        const int p = p0 < q0 ? p0 : q0, q = p0 < q0 ? q0 : p0;
        if (q >= N) continue;  // the bye
        const double ck = shfl_d(c, base + p), sk = shfl_d(s, base + p);
        const double ax = a[p], ay = a[q];
        const double nx = fma(-sk, ay, ck * ax), ny = fma(sk, ax, ck * ay);
        const bool own = idx == p || idx == q;  // (p, p), (p, q) / (q, p), (q, q) of the lane's own pair
        a[p] = own ? 0.0 : nx;
        a[q] = own ? 0.0 : ny;
        const double vx = v[p], vy = v[q];
        v[p] = fma(-sk, vy, ck * vx);
        v[q] = fma(sk, vx, ck * vy);
      }
    }
  }

  // ascending order (ties by index), s^-1/2 with the reference's cut-off
  int rk = 0;
#pragma unroll
  for (int l = 0; l < N; ++l) {
    const double wl = shfl_d(dii, base + l);
    rk += (wl < dii) || (wl == dii && l < idx);
  }
  double* Vs = sm + static_cast<size_t>(warp * MPW + grp) * (G::vs + N);
  double* fs = Vs + G::vs;
#pragma unroll
  for (int cidx = 0; cidx < N; ++cidx) {
    const int rc = __shfl_sync(0xffffffffu, rk, base + cidx);
    if (!spare) Vs[idx * (N + 1) + rc] = v[cidx];
  }
  if (!spare) fs[rk] = dii > 1.0e-15 ? 1.0 / sqrt(dii) : 0.0;
  if (valid) evals[b * N + rk] = dii;
  __syncwarp();
  const int nmat = (nbatch - b0) < MPW ? static_cast<int>(nbatch - b0) : MPW;
  const double* Vw = sm + static_cast<size_t>(warp * MPW) * (G::vs + N);
  for (int e = lane; e < nmat * N * N; e += 32) {
    const int g = e / (N * N), k = e - g * (N * N), i = k / N, j = k - i * N;
    const double* Vg = Vw + g * (G::vs + N);
    const double* fg = Vg + G::vs;
    double acc = 0.0;
#pragma unroll
    for (int rr = 0; rr < N; ++rr) acc += Vg[i * (N + 1) + rr] * fg[rr] * Vg[j * (N + 1) + rr];
    const int64_t o = (b0 + g) * (N * N) + k;
    evecs[o] = Vg[i * (N + 1) + j];
    x[o] = acc;
  }
}

template <int N>
int launch_loewdin_reg(cudaStream_t stream, int nbatch, const double* s_ao, double* x, double* evals, double* evecs) {
  using G = LwGeom<N>;
  const int per_cta = G::kWarps * G::MPW;
  const int grid = (nbatch + per_cta - 1) / per_cta;
  loewdin_reg_kernel<N><<<grid, G::kWarps * 32, G::smem, stream>>>(nbatch, s_ao, x, evals, evecs);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool evc_loewdin_reg_supported(int n) { return n >= 2 && n <= 16; }

// smallest batch that takes this kernel; EVC_LOEWDIN_REG_MIN in the environment overrides it (A/B timing of the
// two forms in one build; development aid)
int evc_loewdin_reg_min_batch() {
  static const int v = [] {
    const char* e = getenv("EVC_LOEWDIN_REG_MIN");
    return e ? atoi(e) : 1;
  }();
  return v;
}

int evc_loewdin_reg(evc_ctx* ctx, int nbatch, int n, const double* s_ao, double* x, double* evals, double* evecs) {
  switch (n) {
#define EVC_LW_CASE(N_) case N_: return launch_loewdin_reg<N_>(ctx->stream, nbatch, s_ao, x, evals, evecs);
    EVC_LW_CASE(2) EVC_LW_CASE(3) EVC_LW_CASE(4) EVC_LW_CASE(5) EVC_LW_CASE(6) EVC_LW_CASE(7) EVC_LW_CASE(8)
    EVC_LW_CASE(9) EVC_LW_CASE(10) EVC_LW_CASE(11) EVC_LW_CASE(12) EVC_LW_CASE(13) EVC_LW_CASE(14)
    EVC_LW_CASE(15) EVC_LW_CASE(16)
#undef EVC_LW_CASE
    default: break;
  }
  EVC_REQUIRE(false, "evc_loewdin_reg: n=%d unsupported (2..16)", n);
  return -1;
}
