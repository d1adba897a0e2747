// Persistent, warp-specialised forms of the two per-geometry GEMM kernels of the packed
// prediction step (algebra: packed.cu header; reference arithmetic:
// evcont/ab_initio_gradients_loewdin.py:190-252, 338-339), for n <= 10 orbitals.
//
// One CTA per SM walks its share of the geometries.  Its 16 warps have fixed roles:
//
//   MMA   (warps 0-7)   nothing but the FP64 tensor-core GEMMs (DMMA m8n8k4), one rectangle of
//                       2 x 4 output tiles per warp with the accumulators in registers;
//   FRONT (K8a: warps 8-11, K4p: warps 8-15)  the inputs of geometry g+1: ONE 1-D bulk copy
//                       (TMA unit, completion on an mbarrier) brings the np x np operand image from
//                       HBM; the small n x n matrices come through ordinary loads; builds the
//                       second GEMM operand (Gm / Q) and the first half of the one-electron chain;
//   MID   (K8a: warps 12-15)  between the GEMMs of a geometry: the Y contraction of U0 with X,
//                       the P0 operand, the second half of the one-electron chain.
//
// The roles meet only through mbarriers (per ring stage: load, operand-ready, u0, p0, free), so
// the GEMMs of geometry g run while FRONT stages g+1 and MID finishes g-1: the tensor pipe no
// longer idles through the load / build / store phases that took ~65 % of a CTA's life in the
// one-CTA-per-geometry kernels (profiles/r01d_packed_phase_clocks.txt).
//
// MMA schedule of K8a:  M1(g) | M2(g-1) M3(g-1) | M1(g+1) | M2(g) M3(g) ...  with
//   M1: U0 = T Gm           (U0 overwrites T)        -> MID: Y, Z, P0 (P0 overwrites U0)
//   M2: R  = Gm P0^T        (R overwrites Gm)
//   M3: W  = P0 R  (lower)  -> HBM
// so MID has a whole period to turn U0 into P0, and three stages of two np x np images
// (3 x 2 x 26.9 KB at n = 10) are alive at once.
#include <cstdlib>
#include <mutex>

#include "packed.cuh"

using namespace evcp;

namespace {

constexpr int kMmaWarps = 8;
constexpr int kStages = 3;        // K8a ring
constexpr int kAoStages = 2;      // K4p ring (two CTAs per SM)
constexpr int kAoPipeThreads = 384;  // K4p: 8 MMA + 4 FRONT warps

struct Rect {
  unsigned char r0, nr, c0, nc;  // in units of 8 x 8 tiles; nr == 0: no work
};
struct RectMap {
  Rect it[kMmaWarps];
};

// The output tiles of a per-geometry GEMM are dealt to the 8 MMA warps as rectangles of at most
// 8 tiles (accumulators in registers), no per-tile predicates inside the k-loop (a predicated
// mma.sync costs a WARPSYNC + NOP + predicate bookkeeping per DMMA: a warp's k-step took ~350
// cycles instead of 128, profiles/r02c).  One warp issues a DMMA only every ~27 cycles while a
// sub-partition's pipe takes one every ~15, so a GEMM phase lasts as long as its LARGEST
// rectangle: the partition minimises the largest area (guillotine search), then the rectangles
// are dealt so that the four SM sub-partitions (warp % 4) carry about the same number of tiles.
struct RectList {
  Rect r[kMmaWarps];
  int n;
};

// guillotine partition of the h x w tile block at (r0, c0) into at most k rectangles of area <= cap that
// minimises the number of fragment loads per k-step (sum of rows + columns); returns that sum (or a
// large number if impossible) and fills out[0 .. k) (unused entries have nr == 0)
int guillotine(int r0, int c0, int h, int w, int k, int cap, Rect* out) {
  constexpr int kInf = 1 << 20;
  for (int i = 0; i < k; ++i) out[i] = Rect{0, 0, 0, 0};
  int best = kInf;
  if (h * w <= cap) {
    out[0] = Rect{static_cast<unsigned char>(r0), static_cast<unsigned char>(h), static_cast<unsigned char>(c0),
                  static_cast<unsigned char>(w)};
    best = h + w;
  }
  if (k == 1) return best;
  Rect a[kMmaWarps], b[kMmaWarps];
  for (int vertical = 0; vertical < 2; ++vertical) {
    const int len = vertical ? w : h;
    for (int cut = 1; cut < len; ++cut)
      for (int k1 = 1; k1 < k; ++k1) {
        const int m1 = vertical ? guillotine(r0, c0, h, cut, k1, cap, a) : guillotine(r0, c0, cut, w, k1, cap, a);
        if (m1 >= best) continue;
        const int m2 = vertical ? guillotine(r0, c0 + cut, h, w - cut, k - k1, cap, b)
                                : guillotine(r0 + cut, c0, h - cut, w, k - k1, cap, b);
        if (m1 + m2 >= best) continue;
        best = m1 + m2;
        int n = 0;
        for (int i = 0; i < k1; ++i)
          if (a[i].nr) out[n++] = a[i];
        for (int i = 0; i < k - k1; ++i)
          if (b[i].nr) out[n++] = b[i];
        for (; n < k; ++n) out[n] = Rect{0, 0, 0, 0};
      }
  }
  return best;
}

RectMap build_rects_uncached(int M8, bool lower) {
  RectMap m;
  for (int w = 0; w < kMmaWarps; ++w) m.it[w] = Rect{0, 0, 0, 0};
  Rect items[kMmaWarps];
  int nitems = 0;
  if (!lower) {
    // smallest cap on the largest rectangle for which 8 rectangles suffice, then the fewest fragment loads
    Rect out[kMmaWarps];
    for (int cap = 1; cap <= 8; ++cap)
      if (guillotine(0, 0, M8, M8, kMmaWarps, cap, out) < (1 << 20)) break;
    for (int i = 0; i < kMmaWarps; ++i)
      if (out[i].nr) items[nitems++] = out[i];
  } else {
    // tiles on or below the diagonal, covered by FULL rectangles: rows in groups of `rh`, columns up to the
    // diagonal in chunks of `cw`; the (rh, cw) pair with the smallest largest area that needs <= 8 rectangles
    int best = 1 << 20, brh = 2, bcw = 4;
    for (int rh = 1; rh <= 2; ++rh)
      for (int cw = 1; cw <= 4; ++cw) {
        int cnt = 0, mx = 0;
        for (int r0 = 0; r0 < M8; r0 += rh) {
          const int nr = M8 - r0 < rh ? M8 - r0 : rh, cend = r0 + nr < M8 ? r0 + nr : M8;
          for (int c0 = 0; c0 < cend; c0 += cw) {
            const int nc = cend - c0 < cw ? cend - c0 : cw;
            ++cnt;
            mx = nr * nc > mx ? nr * nc : mx;
          }
        }
        if (cnt <= kMmaWarps && mx < best) { best = mx; brh = rh; bcw = cw; }
      }
    if (M8 == 7) {  // 2 x 2 blocks for the row pairs, the last row as 1 x 4 + 1 x 3: eight rectangles of <= 4 tiles
      const Rect t[8] = {{0, 2, 0, 2}, {2, 2, 0, 2}, {2, 2, 2, 2}, {4, 2, 0, 2}, {4, 2, 2, 2}, {4, 2, 4, 2}, {6, 1, 0, 4}, {6, 1, 4, 3}};
      for (int i = 0; i < 8; ++i) items[nitems++] = t[i];
    } else {
      for (int r0 = 0; r0 < M8; r0 += brh) {
        const int nr = M8 - r0 < brh ? M8 - r0 : brh, cend = r0 + nr < M8 ? r0 + nr : M8;
        for (int c0 = 0; c0 < cend; c0 += bcw) {
          const int nc = cend - c0 < bcw ? cend - c0 : bcw;
          if (nitems < kMmaWarps)
            items[nitems++] = Rect{static_cast<unsigned char>(r0), static_cast<unsigned char>(nr),
                                   static_cast<unsigned char>(c0), static_cast<unsigned char>(nc)};
        }
      }
    }
  }
  int order[kMmaWarps];
  for (int i = 0; i < nitems; ++i) order[i] = i;
  auto area = [&](int i) { return items[i].nr * items[i].nc; };
  for (int i = 0; i < nitems; ++i)
    for (int j = i + 1; j < nitems; ++j)
      if (area(order[j]) > area(order[i])) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  int sload[4] = {0, 0, 0, 0};
  bool used[kMmaWarps] = {false};
  for (int q = 0; q < nitems; ++q) {
    int best = -1;
    for (int w = 0; w < kMmaWarps; ++w) {
      if (used[w]) continue;
      if (best < 0 || sload[w & 3] < sload[best & 3]) best = w;
    }
    used[best] = true;
    sload[best & 3] += area(order[q]);
    m.it[best] = items[order[q]];
  }
  return m;
}

// the partition search is exponential in the tile count: done once per (M8, lower)
RectMap build_rects(int M8, bool lower) {
  static RectMap cache[9][2];
  static bool have[9][2] = {};
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!have[M8][lower ? 1 : 0]) {
    cache[M8][lower ? 1 : 0] = build_rects_uncached(M8, lower);
    have[M8][lower ? 1 : 0] = true;
  }
  return cache[M8][lower ? 1 : 0];
}

// A GEMM operand in shared memory: element (row, k) at p[row * rs + k * ks].  The A operand of
// C = A B uses row = output row; the B operand uses row = output column.
struct Opnd {
  const double* p;
  int rs, ks;
};

// accumulate one NR x NC rectangle of 8 x 8 tiles (NR * NC <= 8) over K4 / 4 k-steps; the fragments of
// step k + 1 are fetched while the DMMAs of step k issue.  No predicates inside the loop.
template <int NR, int NC>
__device__ __forceinline__ void rect_acc_fixed(double (&acc)[8][2], int r0, int c0, int K4, const Opnd A,
                                               const Opnd B) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const double* pa = A.p + (r0 * 8 + g) * A.rs + tg * A.ks;
  const double* pb = B.p + (c0 * 8 + g) * B.rs + tg * B.ks;
  const int sa = 8 * A.rs, sb = 8 * B.rs, ka = 4 * A.ks, kb = 4 * B.ks;
  double a[NR], b[NC];
#pragma unroll
  for (int i = 0; i < NR; ++i) a[i] = pa[i * sa];
#pragma unroll
  for (int j = 0; j < NC; ++j) b[j] = pb[j * sb];
  const int nk = K4 >> 2;
#pragma unroll 2
  for (int kk = 0; kk < nk; ++kk) {
    const bool more = kk + 1 < nk;  // the last step re-reads its own fragments (stays inside the image)
    pa += more ? ka : 0;
    pb += more ? kb : 0;
    double an[NR], bn[NC];
#pragma unroll
    for (int i = 0; i < NR; ++i) an[i] = pa[i * sa];
#pragma unroll
    for (int j = 0; j < NC; ++j) bn[j] = pb[j * sb];
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) dmma8x8x4(acc[i * NC + j][0], acc[i * NC + j][1], a[i], b[j]);
#pragma unroll
    for (int i = 0; i < NR; ++i) a[i] = an[i];
#pragma unroll
    for (int j = 0; j < NC; ++j) b[j] = bn[j];
  }
}

// tile t of the rectangle (row-major within it) is accumulated in acc[t]
__device__ __forceinline__ void rect_acc(double (&acc)[8][2], const Rect rc, int K4, const Opnd A, const Opnd B) {
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
#define EVC_RECT_CASE(NR_, NC_) \
  case NR_ * 16 + NC_: rect_acc_fixed<NR_, NC_>(acc, rc.r0, rc.c0, K4, A, B); break;
  switch (rc.nr * 16 + rc.nc) {  // warp-uniform
    EVC_RECT_CASE(1, 1) EVC_RECT_CASE(1, 2) EVC_RECT_CASE(1, 3) EVC_RECT_CASE(1, 4)
    EVC_RECT_CASE(1, 5) EVC_RECT_CASE(1, 6) EVC_RECT_CASE(1, 7) EVC_RECT_CASE(1, 8)
    EVC_RECT_CASE(2, 1) EVC_RECT_CASE(2, 2) EVC_RECT_CASE(2, 3) EVC_RECT_CASE(2, 4)
    EVC_RECT_CASE(3, 1) EVC_RECT_CASE(3, 2) EVC_RECT_CASE(4, 1) EVC_RECT_CASE(4, 2)
    EVC_RECT_CASE(5, 1) EVC_RECT_CASE(6, 1) EVC_RECT_CASE(7, 1) EVC_RECT_CASE(8, 1)
    default: break;  // no work for this warp
  }
#undef EVC_RECT_CASE
}

// st(row, col, v0, v1): the accumulator pair of (row, col) and (row, col + 1); col is even.  Dispatched on the
// shape like rect_acc, so that the tile coordinates are compile-time offsets (computed at run time they were
// spilled to local memory and every store waited for a local load: profiles/r02x).
template <int NR, int NC, typename ST>
__device__ __forceinline__ void rect_store_fixed(const double (&acc)[8][2], int r0, int c0, ST st) {
  const int lane = threadIdx.x & 31, row = r0 * 8 + (lane >> 2), col = c0 * 8 + (lane & 3) * 2;
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < NC; ++j) st(row + 8 * i, col + 8 * j, acc[i * NC + j][0], acc[i * NC + j][1]);
}

template <typename ST>
__device__ __forceinline__ void rect_store(const double (&acc)[8][2], const Rect rc, ST st) {
#define EVC_RECT_CASE(NR_, NC_) \
  case NR_ * 16 + NC_: rect_store_fixed<NR_, NC_>(acc, rc.r0, rc.c0, st); break;
  switch (rc.nr * 16 + rc.nc) {  // warp-uniform
    EVC_RECT_CASE(1, 1) EVC_RECT_CASE(1, 2) EVC_RECT_CASE(1, 3) EVC_RECT_CASE(1, 4)
    EVC_RECT_CASE(1, 5) EVC_RECT_CASE(1, 6) EVC_RECT_CASE(1, 7) EVC_RECT_CASE(1, 8)
    EVC_RECT_CASE(2, 1) EVC_RECT_CASE(2, 2) EVC_RECT_CASE(2, 3) EVC_RECT_CASE(2, 4)
    EVC_RECT_CASE(3, 1) EVC_RECT_CASE(3, 2) EVC_RECT_CASE(4, 1) EVC_RECT_CASE(4, 2)
    EVC_RECT_CASE(5, 1) EVC_RECT_CASE(6, 1) EVC_RECT_CASE(7, 1) EVC_RECT_CASE(8, 1)
    default: break;
  }
#undef EVC_RECT_CASE
}

// C = op(P) op(Q) for n x n matrices stored [n][ld], by ONE warp on the tensor cores (tp / tq: use the
// transpose of P / Q).  The one-electron chain of K8a: a scalar FP64 instruction of a helper warp waits ~50-300
// cycles for the FP64 datapath behind the streaming DMMAs, so the chain is issued as few, wide instructions.
__device__ __forceinline__ void warp_mm(int n, int ld, double* C, const double* P, bool tp, const double* Q, bool tq) {
  const int lane = threadIdx.x & 31, gq = lane >> 2, tq4 = lane & 3;
  const int n8 = (n + 7) >> 3, k4 = (n + 3) & ~3;
  for (int mt = 0; mt < n8; ++mt)
    for (int nt = 0; nt < n8; ++nt) {
      double c0 = 0.0, c1 = 0.0;
      const int row = mt * 8 + gq, col = nt * 8 + gq;
      for (int k0 = 0; k0 < k4; k0 += 4) {
        const int kk = k0 + tq4;
        const double av = (row < n && kk < n) ? (tp ? P[kk * ld + row] : P[row * ld + kk]) : 0.0;
        const double bv = (col < n && kk < n) ? (tq ? Q[col * ld + kk] : Q[kk * ld + col]) : 0.0;
        dmma8x8x4(c0, c1, av, bv);
      }
      const int cc = nt * 8 + tq4 * 2;
      if (row < n) {
        if (cc < n) C[row * ld + cc] = c0;
        if (cc + 1 < n) C[row * ld + cc + 1] = c1;
      }
    }
  __syncwarp();
}

__device__ __forceinline__ void st2(double* p, double v0, double v1) {  // 16-byte aligned pair
  *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
}

// development aid: clock64 stamps of CTA 0's roles (evc_debug_pipe_clocks); slot layout
// [kernel 0/1][role 0..3][iteration 0..15][event 0..7]
__device__ long long g_pipe_clk[2][4][16][8];  // role 3: further stamps of the MMA role
__device__ int g_pipe_clk_on = 0;
__device__ long long g_pipe_cta[2][512][3];  // development aid: (smid, globaltimer at start, at end) of every CTA
#define PIPE_STAMP(kern, role, it, ev)                                                   \
  do {                                                                                   \
    if (stamp && (it) < 16 && (threadIdx.x & 31) == 0) g_pipe_clk[kern][role][it][ev] = clock64(); \
  } while (0)

struct AoBars {
  uint64_t load[kAoStages], ready[kAoStages], free_[kAoStages];
};

struct PipeBars {
  uint64_t load[kStages], ready[kStages], u0[kStages], p0[kStages], free_[kStages];
  // CHAIN warp c handles every 4th geometry: it gets its own barriers (one phase per geometry it handles), because
  // a parity wait is only meaningful for a waiter that observes EVERY phase of a barrier
  uint64_t cready[4], cyz[4];
};
static_assert(sizeof(PipeBars) <= 256 && sizeof(AoBars) <= 256, "the barrier block of the pipelined kernels");

template <int NC>
struct PipeSizes {
  static constexpr int n = NC, n2 = NC * NC, np = NC * (NC + 1) / 2, ld = NC | 1;
  static constexpr int M8 = (np + 7) / 8, rows8 = M8 * 8, K4 = (np + 3) & ~3;
  static constexpr int pA = ((K4 & 7) == 4) ? K4 : K4 + 4;       // K4 % 4 == 0: next value == 4 (mod 8)
  static constexpr int pB = rows8 + 4;                           // rows8 % 8 == 0
  static constexpr int szA = rows8 * pA, szB = K4 * pB;
  static constexpr int SZ = ((szA > szB ? szA : szB) + 15) & ~15;  // doubles per image buffer (128-byte multiple)
  static constexpr int mat = n * ld;                              // doubles of one n x n matrix
  static constexpr int ncol = (np + 31) & ~31;                    // pair columns rounded to whole warps
};

__device__ __forceinline__ void build_pij(int n, unsigned short* pij, int tid, int nthreads) {
  for (int k = tid; k < n * n; k += nthreads) {
    const int i = k / n, j = k - i * n;
    if (i >= j) pij[tri_idx(i, j)] = static_cast<unsigned short>(i | (j << 8));
  }
}

// ---------------------------------------------------------------------------------------------
// K4p: hvec[g] = [X^T hcore X | tril(Q^T ERIp Q)],  T[g] = ERIp Q      (persistent, pipelined)
// ---------------------------------------------------------------------------------------------
template <int NC>
struct AoSmem {
  using S = PipeSizes<NC>;
  static constexpr size_t bars = 256;  // >= sizeof(PipeBars)
  static constexpr size_t big = static_cast<size_t>(kAoStages) * 2 * S::SZ * sizeof(double);
  static constexpr size_t smalls = static_cast<size_t>(3) * S::mat * sizeof(double);  // X, Hc, T1 (front-private)
  static constexpr size_t tables = (static_cast<size_t>(S::np) * sizeof(unsigned short) + 15) / 16 * 16;
  static constexpr size_t total = bars + big + smalls + tables;
};

template <int NC>
__global__ void __launch_bounds__(kAoPipeThreads, 2)
ao2oao_pipe_kernel(const __grid_constant__ RectMap mfull, const __grid_constant__ RectMap mlow, int nbatch, int64_t L8, const double* __restrict__ x,
                   const double* __restrict__ hcore, const double* __restrict__ erip, double* __restrict__ hvec,
                   double* __restrict__ Tout) {
  using S = PipeSizes<NC>;
  constexpr int n = S::n, n2 = S::n2, np = S::np, ld = S::ld, pA = S::pA, pB = S::pB, K4 = S::K4, rows8 = S::rows8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AoBars* bars = reinterpret_cast<AoBars*>(smem_raw);
  double* big = reinterpret_cast<double*>(smem_raw + AoSmem<NC>::bars);
  double* Xs = big + static_cast<size_t>(kAoStages) * 2 * S::SZ;
  double* Hs = Xs + S::mat;
  double* T1 = Hs + S::mat;
  unsigned short* pij = reinterpret_cast<unsigned short*>(T1 + S::mat);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kFrontThreads = kAoPipeThreads - kMmaWarps * 32;  // 128
  constexpr int kFrontWarps = kFrontThreads / 32, kRowGroups = kFrontWarps / 2;

  if (tid == 0) {
    for (int s = 0; s < kAoStages; ++s) {
      mbar_init(&bars->load[s], 1);
      mbar_init(&bars->ready[s], kFrontWarps);
      mbar_init(&bars->free_[s], kMmaWarps);
    }
    mbar_init_fence();
  }
  build_pij(n, pij, tid, kAoPipeThreads);
  __syncthreads();
  if (g_pipe_clk_on && tid == 0 && blockIdx.x < 512) {
    unsigned smid;
    long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_pipe_cta[0][blockIdx.x][0] = smid;
    g_pipe_cta[0][blockIdx.x][1] = t;
  }

  const int nloc = (nbatch > static_cast<int>(blockIdx.x))
                       ? (nbatch - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;
  constexpr unsigned kLoadBytes = static_cast<unsigned>(np) * pA * sizeof(double);

  if (warp >= kMmaWarps) {
    // ------------------------------ FRONT ------------------------------
    const int ftid = tid - kMmaWarps * 32, fwarp = warp - kMmaWarps;
    const bool stamp = g_pipe_clk_on && blockIdx.x == 0 && fwarp == 0;
    for (int it = 0; it < nloc; ++it) {
      const int64_t g = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x;
      const int s = it % kAoStages, use = it / kAoStages;
      double* B1 = big + static_cast<size_t>(s) * 2 * S::SZ;  // ERIp (A-type) -> T (B-type)
      double* B2 = B1 + S::SZ;                                // Q (B-type)
      PIPE_STAMP(0, 1, it, 0);
      if (use > 0) mbar_wait(&bars->free_[s], (use - 1) & 1);
      PIPE_STAMP(0, 1, it, 1);
      if (ftid == 0) {
        fence_proxy_async();
        mbar_arrive_expect_tx(&bars->load[s], kLoadBytes);
        bulk_g2s(B1, erip + g * (static_cast<int64_t>(np) * pA), kLoadBytes, &bars->load[s]);
      }
      named_sync(2, kFrontThreads);  // the previous geometry's readers of X / Hc / T1 are done
      for (int k = ftid; k < n2; k += kFrontThreads) {
        const int i = k / n, j = k - i * n;
        Xs[i * ld + j] = __ldg(x + g * n2 + k);
        Hs[i * ld + j] = __ldg(hcore + g * n2 + k);
      }
      named_sync(2, kFrontThreads);
      PIPE_STAMP(0, 1, it, 2);
      // Q[CD][K] = (X_ck X_dl + X_dk X_cl) / s_CD: a thread owns column K = (k, l) and keeps the two
      // columns of X it needs in registers
      {
        const int col = (fwarp & 1) * 32 + lane, rg = fwarp >> 1;  // rows CD with CD % kRowGroups == rg
        if (col < np) {
          const int k = pij[col] & 0xff, l = pij[col] >> 8;
          double xk[NC], xl[NC];
#pragma unroll
          for (int a = 0; a < NC; ++a) { xk[a] = Xs[a * ld + k]; xl[a] = Xs[a * ld + l]; }
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int d = 0; d <= c; ++d) {
              const int CD = c * (c + 1) / 2 + d;
              if (CD % kRowGroups == rg)
                B2[CD * pB + col] = (c == d ? 0.5 : 1.0) * (xk[c] * xl[d] + xk[d] * xl[c]);
            }
        }
        // zero padding of Q: columns [np, pB) of the data rows, rows [np, K4)
        constexpr int padq = pB - np;  // >= 4
        for (int k = ftid; k < np * padq; k += kFrontThreads) {
          const int r = k / padq, c = np + (k - r * padq);
          B2[r * pB + c] = 0.0;
        }
        for (int k = np * pB + ftid; k < K4 * pB; k += kFrontThreads) B2[k] = 0.0;
      }
      PIPE_STAMP(0, 1, it, 3);
      // h1 = X^T (hcore X)
      for (int k = ftid; k < n2; k += kFrontThreads) {
        const int i = k / n, j = k - i * n;
        double acc1 = 0.0;
#pragma unroll
        for (int r = 0; r < NC; ++r) acc1 += Hs[i * ld + r] * Xs[r * ld + j];
        T1[i * ld + j] = acc1;
      }
      named_sync(2, kFrontThreads);
      double* hv = hvec + g * L8;
      for (int k = ftid; k < n2; k += kFrontThreads) {
        const int i = k / n, j = k - i * n;
        double acc1 = 0.0;
#pragma unroll
        for (int r = 0; r < NC; ++r) acc1 += Xs[r * ld + i] * T1[r * ld + j];
        hv[k] = acc1;
      }
      for (int64_t k = n2 + static_cast<int64_t>(np) * (np + 1) / 2 + ftid; k < L8; k += kFrontThreads) hv[k] = 0.0;
      // the bulk copy has landed: zero the padding of the ERIp image (the array in HBM has none)
      PIPE_STAMP(0, 1, it, 4);
      mbar_wait(&bars->load[s], use & 1);
      PIPE_STAMP(0, 1, it, 5);
      {
        constexpr int padc = pA - np;
        if constexpr (padc > 0)
          for (int k = ftid; k < np * padc; k += kFrontThreads) {
            const int r = k / padc, c = np + (k - r * padc);
            B1[r * pA + c] = 0.0;
          }
        for (int k = np * pA + ftid; k < rows8 * pA; k += kFrontThreads) B1[k] = 0.0;
      }
      mbar_arrive_warp(&bars->ready[s]);
      PIPE_STAMP(0, 1, it, 6);
    }
  } else {
    // ------------------------------ MMA ------------------------------
    const Rect rf = mfull.it[warp], rl = mlow.it[warp];
    const bool stamp = g_pipe_clk_on && blockIdx.x == 0 && warp == 0;
    double acc[8][2];
    for (int it = 0; it < nloc; ++it) {
      const int64_t g = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x;
      const int s = it % kAoStages, use = it / kAoStages;
      double* B1 = big + static_cast<size_t>(s) * 2 * S::SZ;
      double* B2 = B1 + S::SZ;
      PIPE_STAMP(0, 0, it, 0);
      mbar_wait(&bars->load[s], use & 1);
      mbar_wait(&bars->ready[s], use & 1);
      PIPE_STAMP(0, 0, it, 1);
      // T = ERIp Q
      rect_acc(acc, rf, K4, Opnd{B1, pA, 1}, Opnd{B2, 1, pB});
      PIPE_STAMP(0, 0, it, 2);
      named_sync(1, kMmaWarps * 32);  // every warp is done reading ERIp: T may overwrite it
      PIPE_STAMP(0, 0, it, 3);
      double* Tg = Tout + g * (static_cast<int64_t>(np) * pA);
      rect_store(acc, rf, [&](int m, int c, double v0, double v1) {
        if (m < K4) st2(B1 + m * pB + c, v0, v1);
        if (m < np && c < np) st2(Tg + m * pA + c, v0, v1);  // c + 1 <= np - 1 or a padding column of the row
      });
      PIPE_STAMP(0, 0, it, 4);
      named_sync(1, kMmaWarps * 32);
      PIPE_STAMP(0, 0, it, 5);
      // h2p = Q^T T, lower triangle only
      rect_acc(acc, rl, K4, Opnd{B2, 1, pB}, Opnd{B1, 1, pB});
      PIPE_STAMP(0, 0, it, 6);
      mbar_arrive_warp(&bars->free_[s]);
      double* hv = hvec + g * L8;
      rect_store(acc, rl, [&](int m, int c, double v0, double v1) {
        if (m < np) {
          if (c <= m) hv[n2 + tri_idx(m, c)] = v0;
          if (c + 1 <= m) hv[n2 + tri_idx(m, c + 1)] = v1;
        }
      });
      PIPE_STAMP(0, 0, it, 7);
    }
    if (g_pipe_clk_on && tid == 0 && blockIdx.x < 512) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      g_pipe_cta[0][blockIdx.x][2] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K8a: W[g] = P0 Gm P0^T (lower), OmS[g], Pao[g]                       (persistent, pipelined)
// ---------------------------------------------------------------------------------------------
template <int NC>
struct GradSmem {
  using S = PipeSizes<NC>;
  static constexpr int kChainWarps = 4;
  static constexpr size_t bars = 256;
  static constexpr size_t big = static_cast<size_t>(kStages) * 2 * S::SZ * sizeof(double);
  // per stage: X, V, Hc, Gam, Ysum, rs, sv | MID-private: Ypart | per CHAIN warp: A, Bm, Z, Qh, Gs, Ps
  static constexpr size_t vec = (S::n + 1) & ~1;
  static constexpr size_t per_stage = static_cast<size_t>(5) * S::mat + 2 * vec;
  static constexpr size_t ypart = static_cast<size_t>(2) * S::ncol * 2 * S::n;  // 2 row groups
  static constexpr size_t chain_priv = static_cast<size_t>(6) * S::mat;
  static constexpr size_t smalls = (kStages * per_stage + ypart + kChainWarps * chain_priv) * sizeof(double);
  static constexpr size_t ntri = static_cast<size_t>(S::np) * (S::np + 1) / 2;
  static constexpr size_t tables = ((S::np + ntri) * sizeof(unsigned short) + 15) / 16 * 16;
  static constexpr size_t total = bars + big + smalls + tables;
};

constexpr int kGradPipeThreads = 640;  // 8 MMA + 4 FRONT + 4 MID + 4 CHAIN warps

template <int NC>
__global__ void __launch_bounds__(kGradPipeThreads, 1)
grad_pipe_kernel(const __grid_constant__ RectMap mfull, const __grid_constant__ RectMap mlow, int nbatch, int64_t L8, const double* __restrict__ x,
                 const double* __restrict__ evals, const double* __restrict__ evecs,
                 const double* __restrict__ hcore, const double* __restrict__ Timg,
                 const double* __restrict__ out7, double* __restrict__ Wout, double* __restrict__ OmSout,
                 double* __restrict__ PaoOut) {
  using S = PipeSizes<NC>;
  using M = GradSmem<NC>;
  constexpr int n = S::n, n2 = S::n2, np = S::np, ld = S::ld, pA = S::pA, pB = S::pB, K4 = S::K4, rows8 = S::rows8;
  constexpr int ntri = static_cast<int>(M::ntri);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PipeBars* bars = reinterpret_cast<PipeBars*>(smem_raw);
  double* big = reinterpret_cast<double*>(smem_raw + M::bars);
  double* stage_small = big + static_cast<size_t>(kStages) * 2 * S::SZ;  // [kStages][X, V, Hc, Gam, Ysum, rs, sv]
  double* Ypart = stage_small + kStages * M::per_stage;                  // MID-private
  double* chain_small = Ypart + M::ypart;                                // [kChainWarps][A, Bm, Z, Qh, Gs, Ps]
  unsigned short* pij = reinterpret_cast<unsigned short*>(chain_small + M::kChainWarps * M::chain_priv);
  unsigned short* trc = pij + np;  // (row | col << 8) of packed triangle entry t
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kFrontWarps = 4, kMidWarps = 4, kChainWarps = M::kChainWarps, kFrontThreads = 128, kMidThreads = 128;
  static_assert(kGradPipeThreads == (kMmaWarps + kFrontWarps + kMidWarps + kChainWarps) * 32, "role split");

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->load[s], 1);
      mbar_init(&bars->ready[s], kFrontWarps);
      mbar_init(&bars->u0[s], kMmaWarps);
      mbar_init(&bars->p0[s], kMidWarps);
      mbar_init(&bars->free_[s], kMmaWarps + kMidWarps + 1);  // + the CHAIN warp of the geometry
    }
    for (int c = 0; c < kChainWarps; ++c) {
      mbar_init(&bars->cready[c], kFrontWarps);
      mbar_init(&bars->cyz[c], kMidWarps);
    }
    mbar_init_fence();
  }
  build_pij(n, pij, tid, kGradPipeThreads);
  for (int r = tid; r < np; r += kGradPipeThreads)
    for (int c = 0; c <= r; ++c) trc[tri_idx(r, c)] = static_cast<unsigned short>(r | (c << 8));
  __syncthreads();

  const int nloc = (nbatch > static_cast<int>(blockIdx.x))
                       ? (nbatch - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;
  constexpr unsigned kLoadBytes = static_cast<unsigned>(np) * pA * sizeof(double);

  if (warp >= kMmaWarps + kFrontWarps + kMidWarps) {
    // ------------------------------ CHAIN ------------------------------
    // The one-electron chain of a geometry, by ONE warp on the tensor cores (warp_mm): Pao = X gamma X^T,
    // Qh = hcore X (gamma + gamma^T), then, once MID has summed Y, Z = Y/2 + Qh,
    // Omega = V (G o (V^T Z V)) V^T, OmS = Omega + Omega^T.  Eight dependent small products: ~10 k cycles of
    // latency under the streaming DMMAs, hidden by giving consecutive geometries to different warps.
    const int cw = warp - (kMmaWarps + kFrontWarps + kMidWarps);
    double* A = chain_small + cw * M::chain_priv;
    double* Bm = A + S::mat, *Z = Bm + S::mat, *Qh = Z + S::mat, *Gs = Qh + S::mat, *Ps = Gs + S::mat;
    static_assert(kChainWarps == 4, "cready / cyz are indexed with it & 3");
    for (int it = cw, round = 0; it < nloc; it += kChainWarps, ++round) {
      const int64_t g = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x;
      const int s = it % kStages;
      const double* Xs = stage_small + s * M::per_stage;
      const double* Vs = Xs + S::mat;
      const double* Hc = Vs + S::mat;
      const double* Gam = Hc + S::mat;
      const double* Ysum = Gam + S::mat;
      const double* rs = Ysum + S::mat;
      const double* sv = rs + M::vec;
      mbar_wait(&bars->cready[cw], round & 1);
      warp_mm(n, ld, A, Xs, false, Gam, false);      // A = X gamma
      warp_mm(n, ld, Ps, A, false, Xs, true);        // Pao = A X^T
      for (int k = lane; k < n2; k += 32) {
        const int i = k / n, j = k - i * n;
        PaoOut[g * n2 + k] = Ps[i * ld + j];
        Bm[i * ld + j] = Gam[i * ld + j] + Gam[j * ld + i];
        // G_pq = -1/(sqrt(s_p) sqrt(s_q) (sqrt(s_p) + sqrt(s_q))), exact divided differences of s^-1/2
        const double rp = rs[i], rq = rs[j];
        double gpq = 0.0;
        if (rp > 0.0 && rq > 0.0) {
          gpq = -1.0 / (rp * rq * (rp + rq));
        } else if ((rp > 0.0) != (rq > 0.0)) {
          const double sp = sv[i], sq = sv[j];
          if (sp != sq) gpq = ((rp > 0.0 ? 1.0 / rp : 0.0) - (rq > 0.0 ? 1.0 / rq : 0.0)) / (sp - sq);
        }
        Gs[i * ld + j] = gpq;
      }
      __syncwarp();
      warp_mm(n, ld, A, Xs, false, Bm, false);       // A = X (gamma + gamma^T)
      warp_mm(n, ld, Qh, Hc, false, A, false);       // Qh = hcore A
      mbar_wait(&bars->cyz[cw], round & 1);          // MID has summed Y
      for (int k = lane; k < n2; k += 32) {
        const int i = k / n, j = k - i * n;
        Z[i * ld + j] = Ysum[i * ld + j] + Qh[i * ld + j];
      }
      __syncwarp();
      warp_mm(n, ld, A, Vs, true, Z, false);         // A = V^T Z
      warp_mm(n, ld, Bm, A, false, Vs, false);       // Bm = A V
      for (int k = lane; k < n2; k += 32) {
        const int i = k / n, j = k - i * n;
        Bm[i * ld + j] *= Gs[i * ld + j];
      }
      __syncwarp();
      warp_mm(n, ld, A, Vs, false, Bm, false);       // A = V Bm
      warp_mm(n, ld, Z, A, false, Vs, true);         // Omega = A V^T
      for (int k = lane; k < n2; k += 32) {
        const int i = k / n, j = k - i * n;
        OmSout[g * n2 + k] = Z[i * ld + j] + Z[j * ld + i];
      }
      mbar_arrive_warp(&bars->free_[s]);
    }
  } else if (warp >= kMmaWarps + kFrontWarps) {
    // ------------------------------ MID ------------------------------
    // between the GEMMs of a geometry: Y from U0, then the P0 operand (the MMA warps wait for it), then the
    // sum of the Y partials for the CHAIN warp.
    const int mtid = tid - (kMmaWarps + kFrontWarps) * 32, mwarp = warp - (kMmaWarps + kFrontWarps);
    const int col = (mwarp & 1) * 32 + lane, rg = mwarp >> 1;  // pair column; pair rows AB with (AB & 1) == rg
    const bool active = col < np;
    const int ci = active ? (pij[col] & 0xff) : 0, cj = active ? (pij[col] >> 8) : 0;
    const bool stamp = g_pipe_clk_on && blockIdx.x == 0 && mwarp == 0;
    const int e = mtid >> 1, half = mtid & 1;
    for (int it = 0; it < nloc; ++it) {
      const int s = it % kStages, use = it / kStages;
      double* B1 = big + static_cast<size_t>(s) * 2 * S::SZ;  // U0 -> P0 (A-type)
      double* Xs = stage_small + s * M::per_stage;
      double* Ysum = Xs + 4 * S::mat;
      PIPE_STAMP(1, 2, it, 0);
      mbar_wait(&bars->ready[s], use & 1);
      mbar_wait(&bars->u0[s], use & 1);
      PIPE_STAMP(1, 2, it, 1);
      // Y[a, i] = 2 sum_{b j} X_bj s_(ij) U0[(ab), (ij)]: the owner of pair column (i, j) walks its half of the
      // pair rows (a, b) and keeps its contributions to Y[:, i] and Y[:, j] in registers
      if (active) {
        double yi[NC], yj[NC], xjs[NC], xis[NC];
        const double sd = (ci == cj) ? 2.0 : 1.0, so = (ci == cj) ? 0.0 : 1.0;
#pragma unroll
        for (int a = 0; a < NC; ++a) {
          yi[a] = yj[a] = 0.0;
          xjs[a] = sd * Xs[a * ld + cj];
          xis[a] = so * Xs[a * ld + ci];
        }
#pragma unroll
        for (int a = 0; a < NC; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            const int AB = a * (a + 1) / 2 + b;
            if ((AB & 1) == rg) {
              const double u = B1[AB * pA + col];
              yi[a] = fma(u, xjs[b], yi[a]);
              yj[a] = fma(u, xis[b], yj[a]);
              if (a != b) {
                yi[b] = fma(u, xjs[a], yi[b]);
                yj[b] = fma(u, xis[a], yj[b]);
              }
            }
          }
        // layout [row group][yi | yj][a][pair column]: a warp's lanes store to consecutive words (with the pair
        // column outermost every store was an 8-way bank conflict: profiles/r02_pipe_ncu_full.txt)
        double* yp = Ypart + static_cast<size_t>(rg) * (2 * n) * S::ncol + col;
#pragma unroll
        for (int a = 0; a < NC; ++a) { yp[a * S::ncol] = yi[a]; yp[(n + a) * S::ncol] = yj[a]; }
      }
      PIPE_STAMP(1, 2, it, 2);
      named_sync(3, kMidThreads);  // every reader of U0 is done; the Y partials are visible
      // P0[AB][I] = X_ai X_bj + X_aj X_bi  (overwrites U0)
      if (active) {
        double xi[NC], xj[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a) { xi[a] = Xs[a * ld + ci]; xj[a] = Xs[a * ld + cj]; }
#pragma unroll
        for (int a = 0; a < NC; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            const int AB = a * (a + 1) / 2 + b;
            if ((AB & 1) == rg) B1[AB * pA + col] = fma(xi[a], xj[b], xj[a] * xi[b]);
          }
      }
      mbar_arrive_warp(&bars->p0[s]);
      PIPE_STAMP(1, 2, it, 3);
      // Y = sum of the partials in a fixed order (thread `half` sums row group `half`, one shuffle)
      for (int e0 = 0; e0 < n2; e0 += kMidThreads / 2) {
        const int ee = e0 + e, ea = ee < n2 ? ee / n : 0, ec = ee < n2 ? ee - ea * n : 0;
        const double* yh = Ypart + static_cast<size_t>(half) * S::ncol * (2 * n);
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int q = 0; q < NC; ++q) {  // q <= ec: column (ec, q) holds Y[:, ec] in its yi; else column (q, ec) in its yj
          const int off = q <= ec ? ea * S::ncol + tri_idx(ec, q) : (n + ea) * S::ncol + tri_idx(q, ec);
          if (q & 1) t1 += yh[off]; else t0 += yh[off];
        }
        double t = t0 + t1;
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        if (ee < n2 && half == 0) Ysum[ea * ld + ec] = t;
      }
      mbar_arrive_warp(&bars->cyz[it & 3]);
      PIPE_STAMP(1, 2, it, 4);
      named_sync(3, kMidThreads);  // the Y partials are consumed before the next geometry overwrites them
      mbar_arrive_warp(&bars->free_[s]);
      PIPE_STAMP(1, 2, it, 5);
    }
  } else if (warp >= kMmaWarps) {
    // ------------------------------ FRONT ------------------------------
    // the inputs of geometry it: T by one bulk copy, the small matrices and the packed pair block of out7 by
    // ordinary loads (all issued before the first dependent store: one round trip), Gm expanded to its image.
    const int ftid = tid - kMmaWarps * 32;
    const bool stamp = g_pipe_clk_on && blockIdx.x == 0 && warp == kMmaWarps;
    for (int it = 0; it < nloc; ++it) {
      const int64_t g = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x;
      const int s = it % kStages, use = it / kStages;
      double* B1 = big + static_cast<size_t>(s) * 2 * S::SZ;  // T (A-type)
      double* B2 = B1 + S::SZ;                                // Gm (A-type)
      double* Xs = stage_small + s * M::per_stage;
      double* Vs = Xs + S::mat;
      double* Hc = Vs + S::mat;
      double* Gam = Hc + S::mat;
      double* rs = Gam + 2 * S::mat;
      double* sv = rs + M::vec;
      PIPE_STAMP(1, 1, it, 0);
      if (use > 0) mbar_wait(&bars->free_[s], (use - 1) & 1);
      PIPE_STAMP(1, 1, it, 1);
      if (ftid == 0) {
        fence_proxy_async();
        mbar_arrive_expect_tx(&bars->load[s], kLoadBytes);
        bulk_g2s(B1, Timg + g * (static_cast<int64_t>(np) * pA), kLoadBytes, &bars->load[s]);
      }
      const double* o7 = out7 + g * L8;
      {
        constexpr int kPer = (ntri + kFrontThreads - 1) / kFrontThreads;
        double v[kPer], sx = 0.0, sv_ = 0.0, sh = 0.0, sg = 0.0, se = 0.0;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int t = ftid + q * kFrontThreads;
          v[q] = t < ntri ? __ldg(o7 + n2 + t) : 0.0;
        }
        if (ftid < n2) {
          sx = __ldg(x + g * n2 + ftid);
          sv_ = __ldg(evecs + g * n2 + ftid);
          sh = __ldg(hcore + g * n2 + ftid);
          sg = __ldg(o7 + ftid);
        }
        if (ftid < n) se = __ldg(evals + g * n + ftid);
        static_assert(n2 <= kFrontThreads, "one small-matrix element per FRONT thread");
        if (ftid < n2) {
          const int i = ftid / n, j = ftid - i * n;
          Xs[i * ld + j] = sx;
          Vs[i * ld + j] = sv_;
          Hc[i * ld + j] = sh;
          Gam[i * ld + j] = sg;
        }
        if (ftid < n) {
          sv[ftid] = se;
          rs[ftid] = se > 1.0e-15 ? sqrt(se) : 0.0;
        }
        // Gm = sym(out7 pair block) as an A-type image: every packed entry is read once (coalesced)
        // and written to (r, c) and (c, r)
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int t = ftid + q * kFrontThreads;
          if (t < ntri) {
            const int r = trc[t] & 0xff, c = trc[t] >> 8;
            B2[r * pA + c] = v[q];
            B2[c * pA + r] = v[q];
          }
        }
        constexpr int padc = pA - np;
        if constexpr (padc > 0)
          for (int k = ftid; k < np * padc; k += kFrontThreads) {
            const int r = k / padc, c = np + (k - r * padc);
            B2[r * pA + c] = 0.0;
          }
        for (int k = np * pA + ftid; k < rows8 * pA; k += kFrontThreads) B2[k] = 0.0;
      }
      PIPE_STAMP(1, 1, it, 2);
      // the bulk copy has landed: zero the padding of the T image (the array in HBM has none)
      mbar_wait(&bars->load[s], use & 1);
      PIPE_STAMP(1, 1, it, 5);
      {
        constexpr int padc = pA - np;
        if constexpr (padc > 0)
          for (int k = ftid; k < np * padc; k += kFrontThreads) {
            const int r = k / padc, c = np + (k - r * padc);
            B1[r * pA + c] = 0.0;
          }
        for (int k = np * pA + ftid; k < rows8 * pA; k += kFrontThreads) B1[k] = 0.0;
      }
      mbar_arrive_warp(&bars->ready[s]);
      mbar_arrive_warp(&bars->cready[it & 3]);
      PIPE_STAMP(1, 1, it, 6);
    }
  } else {
    // ------------------------------ MMA ------------------------------
    const Rect rf = mfull.it[warp], rl = mlow.it[warp];
    const bool stamp = g_pipe_clk_on && blockIdx.x == 0 && warp == 0;
    constexpr int kMmaThreads = kMmaWarps * 32;
    double acc[8][2];
    // M1(it) | M2(it - 1) M3(it - 1): iteration nloc only finishes the last geometry
    for (int itx = 0; itx <= nloc; ++itx) {
      if (itx < nloc) {
        const int it = itx;
        const int s = it % kStages, use = it / kStages;
        double* B1 = big + static_cast<size_t>(s) * 2 * S::SZ;
        double* B2 = B1 + S::SZ;
        PIPE_STAMP(1, 0, it, 0);
        mbar_wait(&bars->load[s], use & 1);
        mbar_wait(&bars->ready[s], use & 1);
        PIPE_STAMP(1, 0, it, 1);
        // U0 = T Gm  (Gm symmetric: B(k, c) = Gm[c][k])
        rect_acc(acc, rf, K4, Opnd{B1, pA, 1}, Opnd{B2, pA, 1});
        PIPE_STAMP(1, 3, it, 0);
        named_sync(1, kMmaThreads);  // every warp is done reading T: U0 may overwrite it
        PIPE_STAMP(1, 3, it, 1);
        rect_store(acc, rf, [&](int m, int c, double v0, double v1) {
          if (c < K4) st2(B1 + m * pA + c, v0, v1);
        });
        mbar_arrive_warp(&bars->u0[s]);
        PIPE_STAMP(1, 0, it, 2);
      }
      if (itx > 0) {
        const int it = itx - 1;
        const int64_t g = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x;
        const int s = it % kStages, use = it / kStages;
        double* B1 = big + static_cast<size_t>(s) * 2 * S::SZ;
        double* B2 = B1 + S::SZ;
        PIPE_STAMP(1, 0, it, 3);
        mbar_wait(&bars->p0[s], use & 1);
        PIPE_STAMP(1, 0, it, 4);
        // R = Gm P0^T
        rect_acc(acc, rf, K4, Opnd{B2, pA, 1}, Opnd{B1, pA, 1});
        PIPE_STAMP(1, 3, it, 2);
        named_sync(1, kMmaThreads);  // every warp is done reading Gm: R may overwrite it (B-type)
        PIPE_STAMP(1, 3, it, 3);
        rect_store(acc, rf, [&](int m, int c, double v0, double v1) {
          if (m < K4) st2(B2 + m * pB + c, v0, v1);
        });
        PIPE_STAMP(1, 3, it, 4);
        named_sync(1, kMmaThreads);
        PIPE_STAMP(1, 0, it, 5);
        // W = P0 R  (symmetric: lower triangle only, the reader takes (max, min))
        rect_acc(acc, rl, K4, Opnd{B1, pA, 1}, Opnd{B2, 1, pB});
        PIPE_STAMP(1, 0, it, 7);
        mbar_arrive_warp(&bars->free_[s]);
        constexpr int npw = (np + 1) & ~1;  // w_pitch_of(n): the pair (c, c + 1), c even, always fits the padded row
        double* Wg = Wout + g * (static_cast<int64_t>(np) * npw);
        rect_store(acc, rl, [&](int m, int c, double v0, double v1) {
          if (m < np && c < np) st2(Wg + m * npw + c, v0, v1);
        });
        PIPE_STAMP(1, 0, it, 6);
      }
    }
  }
}

template <int NC>
int launch_ao2oao_pipe(evc_ctx* ctx, int nbatch, const double* x, const double* hcore, const double* erip,
                       double* hvec, double* Tout) {
  using S = PipeSizes<NC>;
  constexpr size_t smem = AoSmem<NC>::total;
  EVC_REQUIRE(smem <= ctx->smem_optin, "packed_ao2oao (pipelined): needs %zu bytes of shared memory", smem);
  const RectMap mf = build_rects(S::M8, false), ml = build_rects(S::M8, true);
  auto kern = ao2oao_pipe_kernel<NC>;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // two CTAs per SM: the tensor pipe is fed by one CTA while the other stores T / hvec or waits at a barrier
  const int grid = nbatch < 2 * ctx->sm_count ? nbatch : 2 * ctx->sm_count;
  kern<<<grid, kAoPipeThreads, smem, ctx->stream>>>(mf, ml, nbatch, packed_len(NC), x, hcore, erip, hvec, Tout);
  EVC_CHECK_LAUNCH();
  return 0;
}

template <int NC>
int launch_grad_pipe(evc_ctx* ctx, int nbatch, const double* x, const double* evals, const double* evecs,
                     const double* hcore, const double* Timg, const double* out7, double* Wg, double* OmS,
                     double* Pao) {
  using S = PipeSizes<NC>;
  constexpr size_t smem = GradSmem<NC>::total;
  EVC_REQUIRE(smem <= ctx->smem_optin, "packed_grad (pipelined): needs %zu bytes of shared memory", smem);
  const RectMap mf = build_rects(S::M8, false), ml = build_rects(S::M8, true);
  auto kern = grad_pipe_kernel<NC>;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = nbatch < ctx->sm_count ? nbatch : ctx->sm_count;
  kern<<<grid, kGradPipeThreads, smem, ctx->stream>>>(mf, ml, nbatch, packed_len(NC), x, evals, evecs, hcore, Timg,
                                                    out7, Wg, OmS, Pao);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// EVC_PACKED_PIPE=0 in the environment routes n <= 10 to the one-CTA-per-geometry kernels of packed.cu
// (A/B timing of the two forms in one build; development aid)
bool evc_packed_pipe_supported(int n) {
  static const bool enabled = [] {
    const char* e = getenv("EVC_PACKED_PIPE");
    return !(e && e[0] == '0');
  }();
  return enabled && n >= 2 && n <= kPackedPipeMaxNorb;
}

#define EVC_PIPE_DISPATCH(FN, ...)                 \
  switch (n) {                                     \
    case 2: return FN<2>(__VA_ARGS__);             \
    case 3: return FN<3>(__VA_ARGS__);             \
    case 4: return FN<4>(__VA_ARGS__);             \
    case 5: return FN<5>(__VA_ARGS__);             \
    case 6: return FN<6>(__VA_ARGS__);             \
    case 7: return FN<7>(__VA_ARGS__);             \
    case 8: return FN<8>(__VA_ARGS__);             \
    case 9: return FN<9>(__VA_ARGS__);             \
    case 10: return FN<10>(__VA_ARGS__);           \
    default: break;                                \
  }

int evc_packed_ao2oao_pipe(evc_ctx* ctx, int nbatch, int n, const double* x, const double* hcore,
                           const double* erip, double* hvec, double* Tout) {
  EVC_PIPE_DISPATCH(launch_ao2oao_pipe, ctx, nbatch, x, hcore, erip, hvec, Tout)
  evc_set_error("packed_ao2oao (pipelined): n=%d unsupported", n);
  return -1;
}

int evc_packed_grad_pipe(evc_ctx* ctx, int nbatch, int n, const double* x, const double* evals,
                         const double* evecs, const double* hcore, const double* Timg, const double* out7,
                         double* Wg, double* OmS, double* Pao) {
  EVC_PIPE_DISPATCH(launch_grad_pipe, ctx, nbatch, x, evals, evecs, hcore, Timg, out7, Wg, OmS, Pao)
  evc_set_error("packed_grad (pipelined): n=%d unsupported", n);
  return -1;
}

extern "C" {
// development aid: resident CTAs per SM of the n = 10 pipelined kernels (K4p, K8a)
int evc_debug_pipe_occupancy(int* ao2oao_ctas, int* grad_ctas) {
  auto k1 = ao2oao_pipe_kernel<10>;
  auto k2 = grad_pipe_kernel<10>;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(AoSmem<10>::total)));
  EVC_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(GradSmem<10>::total)));
  EVC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ao2oao_ctas, k1, kAoPipeThreads, AoSmem<10>::total));
  EVC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(grad_ctas, k2, kGradPipeThreads, GradSmem<10>::total));
  return 0;
}
// development aid: switch the clock64 stamps of CTA 0 on/off, read them back ([2][3][16][8] int64)
int evc_debug_pipe_ctas(long long* out_host) {  // [2][512][3]
  EVC_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_pipe_cta, sizeof(long long) * 2 * 512 * 3));
  return 0;
}
int evc_debug_pipe_clocks(int enable, long long* out_host) {
  if (out_host) EVC_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_pipe_clk, sizeof(long long) * 2 * 4 * 16 * 8));
  EVC_CHECK_CUDA(cudaMemcpyToSymbol(g_pipe_clk_on, &enable, sizeof(int)));
  return 0;
}
}
