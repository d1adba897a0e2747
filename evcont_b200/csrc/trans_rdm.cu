// K1+K2: transition 1-/2-RDMs between FCI vectors, fused t1 gather + FP64 DMMA.
//
// Replaces cisolver.trans_rdm12(cibra, ciket, norb, nelec) at
// evcont/FCI_EVCont.py:117-127 (PySCF FCIrdm12_drv / FCItdm12kern_sf + reorder_rdm,
// SURVEY.md Appendix A).  Definitions:
//   t1_v[K, p*n+q] = <K|E_pq|v>                                   (gathered)
//   C[x, y] = sum_K braT[x, K] ketT[y, K]                          (DMMA)
//     braT[(p,q)] = t1_bra[:, (q,p)], braT[n^2] = bra        (index swap folded
//     ketT[(r,s)] = t1_ket[:, (r,s)], ketT[n^2] = ket         into the gather)
//   => C[(p,q),(r,s)] = <bra|E_pq E_rs|ket>, C[n^2,(r,s)] = <bra|E_rs|ket>,
//      C[n^2,n^2] = <bra|ket>.
// Only macro-blocks on/below the diagonal of C are computed; the rest follows
// from [E_pq, E_rs] = d_qr E_ps - d_ps E_rq in the finalize kernel, which also
// applies PySCF's reorder (dm2[p,q,r,s] = C[pq,rs] - d_qr <bra|p^+ s|ket>) and the
// dm1 transpose.
//
// Work decomposition: CTA = (pair, slice of alpha strings).  For each alpha string
// Ia of the slice and each tile of Bt beta strings the CTA builds the two
// [W x Bt] tiles in shared memory (K contiguous, row pitch Bp = Bt with
// Bt % 8 == 4 so that both the scatter and the DMMA fragment loads are
// conflict-free), then every warp multiplies its macro-blocks (16x16 outputs
// = 2x2 DMMA tiles) with accumulators in registers.  Partial sums per CTA go to
// the workspace and are reduced in slice order => deterministic.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

struct TrdmParams {
  int norb;
  int n2;          // norb*norb
  int W;           // padded number of columns (n2+1 rounded up to 16)
  int64_t na, nb;
  const double* civecs;
  int64_t vec_stride;
  const int32_t* pairs;
  int nsplit;
  const uint64_t* link_a;  // string-major [na][nlink_a]
  int nlink_a;
  const uint64_t* link_b;  // link-major   [nlink_b][nb]
  int nlink_b;
  int Bt;           // beta tile width, Bt % 8 == 4
  int ntile;        // ceil(nb / Bt)
  double* partial;  // [npairs][nsplit][T][256]
  int T;            // number of lower-triangular macro-blocks
  unsigned char blk_start[16];
  unsigned char blk_count[16];
};

__device__ __forceinline__ void unpack_link(uint64_t rec, int& addr, int& a, int& i, double& sgn) {
  addr = static_cast<int>(rec & 0xffffffffu);
  a = static_cast<int>((rec >> 32) & 0xff);
  i = static_cast<int>((rec >> 40) & 0xff);
  sgn = static_cast<double>(static_cast<signed char>((rec >> 48) & 0xff));
}

// sign of a link as a mask on the IEEE sign bit (the sign byte is +1 or -1): v ^ mask == sgn * v bit for bit, and no
// FP64 instruction (a scalar FP64 operation waits behind the DMMAs that stream through the same datapath)
__device__ __forceinline__ unsigned long long link_sign_mask(uint64_t rec) { return ((rec >> 55) & 1ull) << 63; }
__device__ __forceinline__ double flip(double v, unsigned long long m) {
  return __longlong_as_double(static_cast<long long>(static_cast<unsigned long long>(__double_as_longlong(v)) ^ m));
}

// Build the two [W x Bp] tiles (bra: E_pq stored in column (q,p); ket: column
// (p,q)) for alpha string Ia and beta strings [b0, b0+Bt).  Three phases separated
// by block barriers: zero + stage c[Ia,:]; beta links (pure stores of +-c[Ia, Jb],
// every (column, x) is written by at most one link); alpha links (added to their
// rows with contiguous 16-byte accesses, again one link per (column, x)).
// Thread map of the link phases: one link per warp iteration, the lanes cover the
// beta strings of the tile.  The link record of an alpha link is warp-uniform (one
// decode serves the bra and the ket row copy) and no index is split by a division
// (the flat (link, x) map spent most of the kernel's instructions on that).
template <int NWARPS>
__device__ __forceinline__ void build_tiles(const TrdmParams& P, const double* __restrict__ cbra,
                                            const double* __restrict__ cket, int64_t Ia, int b0,
                                            double* __restrict__ braT, double* __restrict__ ketT,
                                            double* __restrict__ crow_bra,
                                            double* __restrict__ crow_ket) {
  constexpr int nthreads = NWARPS * 32;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = P.norb, Bp = P.Bt;
  const int nb = static_cast<int>(P.nb);
  {
    // braT and ketT are contiguous: one vectorised zero fill
    double2* t2 = reinterpret_cast<double2*>(braT);
    const int tot = P.W * Bp;
    for (int k = tid; k < tot; k += nthreads) t2[k] = make_double2(0.0, 0.0);
    if (b0 == 0) {  // c[Ia, :] rows change only with Ia
      const double* sb = cbra + Ia * P.nb;
      const double* sk = cket + Ia * P.nb;
      for (int k = tid; k < nb; k += nthreads) {
        crow_bra[k] = __ldg(sb + k);
        crow_ket[k] = __ldg(sk + k);
      }
    }
  }
  __syncthreads();
  const int width = min(P.Bt, nb - b0);
  // beta links first, as pure stores into the zeroed tiles (one link per (column, x)); the alpha links then ADD
  // their rows with contiguous 16-byte accesses.  (beta as read-modify-write cost four scattered shared-memory
  // accesses per element instead of two; tile = alpha + beta either way, one rounding.)
  for (int l = warp; l < P.nlink_b; l += NWARPS) {
    const uint64_t* lb = P.link_b + static_cast<int64_t>(l) * P.nb + b0;
    for (int x = lane; x < width; x += 32) {
      const uint64_t rec = __ldg(lb + x);
      int Jb, a, i; double sg;
      unpack_link(rec, Jb, a, i, sg);
      const unsigned long long m = link_sign_mask(rec);
      braT[(a * n + i) * Bp + x] = flip(crow_bra[Jb], m);
      ketT[(i * n + a) * Bp + x] = flip(crow_ket[Jb], m);
    }
  }
  // identity column: the CI coefficients themselves
  for (int x = tid; x < width; x += nthreads) {
    braT[P.n2 * Bp + x] = crow_bra[b0 + x];
    ketT[P.n2 * Bp + x] = crow_ket[b0 + x];
  }
  __syncthreads();
  const uint64_t* la = P.link_a + Ia * P.nlink_a;
  const bool vec2 = ((nb & 1) == 0) && ((b0 & 1) == 0);
  for (int l = warp; l < P.nlink_a; l += NWARPS) {
    const uint64_t rec = __ldg(la + l);
    int Ja, a, i; double sg;
    unpack_link(rec, Ja, a, i, sg);
    const unsigned long long m = link_sign_mask(rec);
    const double* sb = cbra + static_cast<int64_t>(Ja) * P.nb + b0;
    const double* sk = cket + static_cast<int64_t>(Ja) * P.nb + b0;
    double* db = braT + (a * n + i) * Bp;
    double* dk = ketT + (i * n + a) * Bp;
    if (vec2) {
      for (int x = 2 * lane; x < width; x += 64) {
        if (x + 1 < width) {
          const double2 vb = __ldg(reinterpret_cast<const double2*>(sb + x));
          const double2 vk = __ldg(reinterpret_cast<const double2*>(sk + x));
          const double2 tb = *reinterpret_cast<const double2*>(db + x);
          const double2 tk = *reinterpret_cast<const double2*>(dk + x);
          *reinterpret_cast<double2*>(db + x) = make_double2(tb.x + flip(vb.x, m), tb.y + flip(vb.y, m));
          *reinterpret_cast<double2*>(dk + x) = make_double2(tk.x + flip(vk.x, m), tk.y + flip(vk.y, m));
        } else {
          db[x] += flip(__ldg(sb + x), m);
          dk[x] += flip(__ldg(sk + x), m);
        }
      }
    } else {
      for (int x = lane; x < width; x += 32) {
        db[x] += flip(__ldg(sb + x), m);
        dk[x] += flip(__ldg(sk + x), m);
      }
    }
  }
}

// The same build with NL link records (and their gathers) in flight per warp: for the instances whose accumulators
// leave the registers for it (at most three macro-blocks per warp).
template <int NWARPS, int NL>
__device__ __forceinline__ void build_tiles_nl(const TrdmParams& P, const double* __restrict__ cbra,
                                            const double* __restrict__ cket, int64_t Ia, int b0,
                                            double* __restrict__ braT, double* __restrict__ ketT,
                                            double* __restrict__ crow_bra,
                                            double* __restrict__ crow_ket) {
  constexpr int nthreads = NWARPS * 32;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = P.norb, Bp = P.Bt;
  const int nb = static_cast<int>(P.nb);
  {
    // braT and ketT are contiguous: one vectorised zero fill
    double2* t2 = reinterpret_cast<double2*>(braT);
    const int tot = P.W * Bp;
    for (int k = tid; k < tot; k += nthreads) t2[k] = make_double2(0.0, 0.0);
    if (b0 == 0) {  // c[Ia, :] rows change only with Ia
      const double* sb = cbra + Ia * P.nb;
      const double* sk = cket + Ia * P.nb;
      for (int k = tid; k < nb; k += nthreads) {
        crow_bra[k] = __ldg(sb + k);
        crow_ket[k] = __ldg(sk + k);
      }
    }
  }
  __syncthreads();
  const int width = min(P.Bt, nb - b0);
  // beta links first, as pure stores into the zeroed tiles (one link per (column, x)); the alpha links then ADD
  // their rows with contiguous 16-byte accesses.  (beta as read-modify-write cost four scattered shared-memory
  // accesses per element instead of two; tile = alpha + beta either way, one rounding.)
  for (int l0 = warp; l0 < P.nlink_b; l0 += NL * NWARPS) {
    for (int x = lane; x < width; x += 32) {
      uint64_t rec[NL];
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        const int l = l0 + j * NWARPS;
        rec[j] = l < P.nlink_b ? __ldg(P.link_b + static_cast<int64_t>(l) * P.nb + b0 + x) : ~0ull;
      }
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        if (rec[j] != ~0ull) {
          int Jb, a, i; double sg;
          unpack_link(rec[j], Jb, a, i, sg);
          const unsigned long long m = link_sign_mask(rec[j]);
          braT[(a * n + i) * Bp + x] = flip(crow_bra[Jb], m);
          ketT[(i * n + a) * Bp + x] = flip(crow_ket[Jb], m);
        }
      }
    }
  }
  // identity column: the CI coefficients themselves
  for (int x = tid; x < width; x += nthreads) {
    braT[P.n2 * Bp + x] = crow_bra[b0 + x];
    ketT[P.n2 * Bp + x] = crow_ket[b0 + x];
  }
  __syncthreads();
  const uint64_t* la = P.link_a + Ia * P.nlink_a;
  const bool vec2 = ((nb & 1) == 0) && ((b0 & 1) == 0) && ((width & 1) == 0);
  for (int l0 = warp; l0 < P.nlink_a; l0 += NL * NWARPS) {
    uint64_t rec[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int l = l0 + j * NWARPS;
      rec[j] = l < P.nlink_a ? __ldg(la + l) : ~0ull;
    }
    if (vec2) {
      for (int x = 2 * lane; x < width; x += 64) {
        double2 vb[NL], vk[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) {
          if (rec[j] != ~0ull) {
            const int64_t off = static_cast<int64_t>(rec[j] & 0xffffffffu) * P.nb + b0 + x;
            vb[j] = __ldg(reinterpret_cast<const double2*>(cbra + off));
            vk[j] = __ldg(reinterpret_cast<const double2*>(cket + off));
          }
        }
#pragma unroll
        for (int j = 0; j < NL; ++j) {
          if (rec[j] != ~0ull) {
            int Ja, a, i; double sg;
            unpack_link(rec[j], Ja, a, i, sg);
            const unsigned long long m = link_sign_mask(rec[j]);
            double2* db = reinterpret_cast<double2*>(braT + (a * n + i) * Bp + x);
            double2* dk = reinterpret_cast<double2*>(ketT + (i * n + a) * Bp + x);
            const double2 tb = *db, tk = *dk;
            *db = make_double2(tb.x + flip(vb[j].x, m), tb.y + flip(vb[j].y, m));
            *dk = make_double2(tk.x + flip(vk[j].x, m), tk.y + flip(vk[j].y, m));
          }
        }
      }
    } else {
      for (int x = lane; x < width; x += 32) {
        double vb[NL], vk[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) {
          if (rec[j] != ~0ull) {
            const int64_t off = static_cast<int64_t>(rec[j] & 0xffffffffu) * P.nb + b0 + x;
            vb[j] = __ldg(cbra + off);
            vk[j] = __ldg(cket + off);
          }
        }
#pragma unroll
        for (int j = 0; j < NL; ++j) {
          if (rec[j] != ~0ull) {
            int Ja, a, i; double sg;
            unpack_link(rec[j], Ja, a, i, sg);
            const unsigned long long m = link_sign_mask(rec[j]);
            braT[(a * n + i) * Bp + x] += flip(vb[j], m);
            ketT[(i * n + a) * Bp + x] += flip(vk[j], m);
          }
        }
      }
    }
  }
}

// DMMA phase of one tile: NB macro-blocks of this warp, straight-line (all fragment loads of a k-step ahead of its
// DMMAs; a per-block `s < my_count` branch kept every block's loads next to its own DMMAs, their latency exposed).
template <int NB, int MAXBLK>
__device__ __forceinline__ void mma_tile(double (&acc)[MAXBLK][4][2], const double* __restrict__ braT,
                                         const double* __restrict__ ketT, const int (&xoff)[MAXBLK],
                                         const int (&yoff)[MAXBLK], int Bt, int Bp) {
#pragma unroll 1
  for (int k0 = 0; k0 < Bt; k0 += 4) {
    double a0[NB > 0 ? NB : 1], a1[NB > 0 ? NB : 1], b0v[NB > 0 ? NB : 1], b1v[NB > 0 ? NB : 1];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      a0[s] = braT[xoff[s] + k0];
      a1[s] = braT[xoff[s] + 8 * Bp + k0];
      b0v[s] = ketT[yoff[s] + k0];
      b1v[s] = ketT[yoff[s] + 8 * Bp + k0];
    }
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      dmma8x8x4(acc[s][0][0], acc[s][0][1], a0[s], b0v[s]);
      dmma8x8x4(acc[s][1][0], acc[s][1][1], a0[s], b1v[s]);
      dmma8x8x4(acc[s][2][0], acc[s][2][1], a1[s], b0v[s]);
      dmma8x8x4(acc[s][3][0], acc[s][3][1], a1[s], b1v[s]);
    }
  }
}

template <int NWARPS, int MAXBLK, int MINCTA>
__global__ void __launch_bounds__(NWARPS * 32, MINCTA)
trdm_fused_kernel(const __grid_constant__ TrdmParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Bp = P.Bt;
  const int nbpad = (static_cast<int>(P.nb) + 1) & ~1;
  double* braT = reinterpret_cast<double*>(smem_raw);
  double* ketT = braT + P.W * Bp;
  double* crow_bra = ketT + P.W * Bp;
  double* crow_ket = crow_bra + nbpad;

  const int item = blockIdx.x;
  const int pair = item / P.nsplit;
  const int split = item - pair * P.nsplit;
  const int ibra = P.pairs[2 * pair], iket = P.pairs[2 * pair + 1];
  const double* cbra = P.civecs + ibra * P.vec_stride;
  const double* cket = P.civecs + iket * P.vec_stride;
  const int64_t ia_lo = P.na * split / P.nsplit;
  const int64_t ia_hi = P.na * (split + 1) / P.nsplit;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int my_start = P.blk_start[warp], my_count = P.blk_count[warp];

  // macro-block coordinates of this warp's slots
  int xoff[MAXBLK], yoff[MAXBLK];
#pragma unroll
  for (int s = 0; s < MAXBLK; ++s) {
    int t = my_start + (s < my_count ? s : 0);
    int r = 0;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    const int cc = t - r * (r + 1) / 2;
    xoff[s] = (r * 16 + g) * Bp + tg;
    yoff[s] = (cc * 16 + g) * Bp + tg;
  }
  double acc[MAXBLK][4][2];
#pragma unroll
  for (int s = 0; s < MAXBLK; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[s][q][0] = acc[s][q][1] = 0.0;

  for (int64_t Ia = ia_lo; Ia < ia_hi; ++Ia) {
    for (int tile = 0; tile < P.ntile; ++tile) {
      const int b0 = tile * P.Bt;
      __syncthreads();  // previous MMA phase done with the tiles
      // two link records in flight per warp where the registers allow (norb = 9: 2.16 -> 2.06 ms; with four or five
      // macro-blocks per warp the second set spills and H2O sizes lose: those keep the one-link build)
      if constexpr (MAXBLK <= 3 && MINCTA * NWARPS <= 16)
        build_tiles_nl<NWARPS, 2>(P, cbra, cket, Ia, b0, braT, ketT, crow_bra, crow_ket);
      else
        build_tiles<NWARPS>(P, cbra, cket, Ia, b0, braT, ketT, crow_bra, crow_ket);
      __syncthreads();
      // a warp has MAXBLK or MAXBLK - 1 macro-blocks (fewer: its spare slots repeat block 0 and are never stored)
      if (my_count == MAXBLK) mma_tile<MAXBLK, MAXBLK>(acc, braT, ketT, xoff, yoff, P.Bt, Bp);
      else mma_tile<MAXBLK - 1, MAXBLK>(acc, braT, ketT, xoff, yoff, P.Bt, Bp);
    }
  }
  // write partial sums: [item][t][tile(2x2)][8x8 row-major]
  double* out = P.partial + static_cast<int64_t>(item) * P.T * 256;
#pragma unroll
  for (int s = 0; s < MAXBLK; ++s) {
    if (s < my_count) {
      double* blk = out + (my_start + s) * 256;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        reinterpret_cast<double2*>(blk + q * 64)[lane] = make_double2(acc[s][q][0], acc[s][q][1]);
    }
  }
}

// ---- producer / consumer form of the fused kernel (7 or 8 macro-block rows: norb = 10, 11) ------------------------
// The fused kernel above alternates a t1 tile build (latency-bound gathers) and a DMMA phase in every warp, two
// CTAs per SM filling each other's gaps: the DMMA pipe is busy 51 % of the time.  Here ONE CTA per SM (16 or 20 warps) holds
// two tile stages: the 8 builder warps build stage (it + 1) while the NMMA (8 or 12) MMA warps multiply stage it.  The builders hold no
// accumulators, so they keep four link records and their eight gathers in flight per lane.  Hand-over by named
// barriers (bar.arrive by the side that is done, bar.sync by the side that waits; the PTX producer / consumer
// pattern).  Same tiles, same k order and same alpha slices as the fused kernel => bit-identical results.
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

constexpr int kPipeBuild = 8;                                    // builder warps (the MMA warps: template NMMA)
constexpr int kBarFull = 1, kBarEmpty = 3, kBarBuild = 5;        // named barriers: full[2], empty[2], builders

template <int NMMA>
__device__ __forceinline__ void pipe_build(const TrdmParams& P, const double* __restrict__ cbra,
                                           const double* __restrict__ cket, int64_t Ia, int b0,
                                           double* __restrict__ braT, double* __restrict__ ketT,
                                           double* __restrict__ crow_bra, double* __restrict__ crow_ket) {
  constexpr int nthreads = kPipeBuild * 32;
  const int tid = threadIdx.x - NMMA * 32;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = P.norb, Bp = P.Bt;
  const int nb = static_cast<int>(P.nb);
  if (b0 == 0) {  // c[Ia, :] rows change only with Ia; the beta phase of the previous step still reads them
    bar_sync_n(kBarBuild, nthreads);
    const double* sb = cbra + Ia * P.nb;
    const double* sk = cket + Ia * P.nb;
    for (int k = tid; k < nb; k += nthreads) {
      crow_bra[k] = __ldg(sb + k);
      crow_ket[k] = __ldg(sk + k);
    }
  }
  {
    double2* t2 = reinterpret_cast<double2*>(braT);  // braT and ketT of a stage are contiguous
    const int tot = P.W * Bp;
    for (int k = tid; k < tot; k += nthreads) t2[k] = make_double2(0.0, 0.0);
  }
  bar_sync_n(kBarBuild, nthreads);
  const int width = min(P.Bt, nb - b0);
  // beta links (pure stores, see build_tiles): four link rows per warp pass, the lanes over the beta strings
  for (int l0 = warp; l0 < P.nlink_b; l0 += 4 * kPipeBuild) {
    for (int x = lane; x < width; x += 32) {
      uint64_t rec[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = l0 + j * kPipeBuild;
        rec[j] = l < P.nlink_b ? __ldg(P.link_b + static_cast<int64_t>(l) * P.nb + b0 + x) : ~0ull;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (rec[j] != ~0ull) {
          int Jb, a, i; double sg;
          unpack_link(rec[j], Jb, a, i, sg);
          const unsigned long long m = link_sign_mask(rec[j]);
          braT[(a * n + i) * Bp + x] = flip(crow_bra[Jb], m);
          ketT[(i * n + a) * Bp + x] = flip(crow_ket[Jb], m);
        }
      }
    }
  }
  // identity column: the CI coefficients themselves
  for (int x = tid; x < width; x += nthreads) {
    braT[P.n2 * Bp + x] = crow_bra[b0 + x];
    ketT[P.n2 * Bp + x] = crow_ket[b0 + x];
  }
  bar_sync_n(kBarBuild, nthreads);
  // alpha links (added to the rows): four links per warp pass, all records, then all gathers and row reads in
  // flight before the first addition
  const uint64_t* la = P.link_a + Ia * P.nlink_a;
  const bool vec2 = ((nb & 1) == 0) && ((b0 & 1) == 0) && ((width & 1) == 0);
  for (int l0 = warp; l0 < P.nlink_a; l0 += 4 * kPipeBuild) {
    uint64_t rec[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = l0 + j * kPipeBuild;
      rec[j] = l < P.nlink_a ? __ldg(la + l) : ~0ull;
    }
    if (vec2) {
      for (int x = 2 * lane; x < width; x += 64) {
        double2 vb[4], vk[4], tb[4], tk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rec[j] != ~0ull) {
            int Ja, a, i; double sg;
            unpack_link(rec[j], Ja, a, i, sg);
            const int64_t off = static_cast<int64_t>(Ja) * P.nb + b0 + x;
            vb[j] = __ldg(reinterpret_cast<const double2*>(cbra + off));
            vk[j] = __ldg(reinterpret_cast<const double2*>(cket + off));
            tb[j] = *reinterpret_cast<const double2*>(braT + (a * n + i) * Bp + x);
            tk[j] = *reinterpret_cast<const double2*>(ketT + (i * n + a) * Bp + x);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rec[j] != ~0ull) {
            int Ja, a, i; double sg;
            unpack_link(rec[j], Ja, a, i, sg);
            const unsigned long long m = link_sign_mask(rec[j]);
            *reinterpret_cast<double2*>(braT + (a * n + i) * Bp + x) =
                make_double2(tb[j].x + flip(vb[j].x, m), tb[j].y + flip(vb[j].y, m));
            *reinterpret_cast<double2*>(ketT + (i * n + a) * Bp + x) =
                make_double2(tk[j].x + flip(vk[j].x, m), tk[j].y + flip(vk[j].y, m));
          }
        }
      }
    } else {
      for (int x = lane; x < width; x += 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rec[j] != ~0ull) {
            int Ja, a, i; double sg;
            unpack_link(rec[j], Ja, a, i, sg);
            const unsigned long long m = link_sign_mask(rec[j]);
            const int64_t off = static_cast<int64_t>(Ja) * P.nb + b0 + x;
            braT[(a * n + i) * Bp + x] += flip(__ldg(cbra + off), m);
            ketT[(i * n + a) * Bp + x] += flip(__ldg(cket + off), m);
          }
        }
      }
    }
  }
}

template <int NMMA, int MAXBLK>
__global__ void __launch_bounds__((NMMA + kPipeBuild) * 32, 1)
trdm_pipe_kernel(const __grid_constant__ TrdmParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Bp = P.Bt;
  const int nbpad = (static_cast<int>(P.nb) + 1) & ~1;
  const int tile_elems = P.W * Bp;
  double* stage0 = reinterpret_cast<double*>(smem_raw);          // stage s: braT | ketT
  double* crow_bra = stage0 + 4 * tile_elems;
  double* crow_ket = crow_bra + nbpad;

  const int item = blockIdx.x;
  const int pair = item / P.nsplit;
  const int split = item - pair * P.nsplit;
  const int ibra = P.pairs[2 * pair], iket = P.pairs[2 * pair + 1];
  const double* cbra = P.civecs + ibra * P.vec_stride;
  const double* cket = P.civecs + iket * P.vec_stride;
  const int64_t ia_lo = P.na * split / P.nsplit;
  const int64_t ia_hi = P.na * (split + 1) / P.nsplit;
  const int nsteps = static_cast<int>(ia_hi - ia_lo) * P.ntile;
  constexpr int kAll = (NMMA + kPipeBuild) * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp >= NMMA) {
    // ---- builders ----
    int64_t Ia = ia_lo;
    int tile = 0;
    for (int it = 0; it < nsteps; ++it) {
      const int s = it & 1;
      if (it >= 2) bar_sync_n(kBarEmpty + s, kAll);  // the MMA warps are done with this stage
      double* braT = stage0 + 2 * s * tile_elems;
      pipe_build<NMMA>(P, cbra, cket, Ia, tile * P.Bt, braT, braT + tile_elems, crow_bra, crow_ket);
      __threadfence_block();
      bar_arrive_n(kBarFull + s, kAll);
      if (++tile == P.ntile) { tile = 0; ++Ia; }
    }
    return;
  }

  // ---- MMA warps ----
  const int g = lane >> 2, tg = lane & 3;
  const int my_start = P.blk_start[warp], my_count = P.blk_count[warp];
  int xoff[MAXBLK], yoff[MAXBLK];
#pragma unroll
  for (int s = 0; s < MAXBLK; ++s) {
    int t = my_start + (s < my_count ? s : 0);
    int r = 0;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    const int cc = t - r * (r + 1) / 2;
    xoff[s] = (r * 16 + g) * Bp + tg;
    yoff[s] = (cc * 16 + g) * Bp + tg;
  }
  double acc[MAXBLK][4][2];
#pragma unroll
  for (int s = 0; s < MAXBLK; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[s][q][0] = acc[s][q][1] = 0.0;

  for (int it = 0; it < nsteps; ++it) {
    const int st = it & 1;
    const double* braT = stage0 + 2 * st * tile_elems;
    const double* ketT = braT + tile_elems;
    bar_sync_n(kBarFull + st, kAll);
    if (my_count == MAXBLK) mma_tile<MAXBLK, MAXBLK>(acc, braT, ketT, xoff, yoff, P.Bt, Bp);
    else mma_tile<MAXBLK - 1, MAXBLK>(acc, braT, ketT, xoff, yoff, P.Bt, Bp);
    if (it + 2 < nsteps) bar_arrive_n(kBarEmpty + st, kAll);
  }
  double* out = P.partial + static_cast<int64_t>(item) * P.T * 256;
#pragma unroll
  for (int s = 0; s < MAXBLK; ++s) {
    if (s < my_count) {
      double* blk = out + (my_start + s) * 256;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        reinterpret_cast<double2*>(blk + q * 64)[lane] = make_double2(acc[s][q][0], acc[s][q][1]);
    }
  }
}

// One CTA per pair: ordered reduction over the alpha slices, triangular fill,
// PySCF reorder, dm1 transpose.
__global__ void trdm_finalize_kernel(int norb, int T, int nsplit, const double* __restrict__ partial,
                                     double* __restrict__ ovlp, int64_t ovlp_stride, double* __restrict__ dm1,
                                     int64_t dm1_stride, double* __restrict__ dm2, int64_t dm2_stride) {
  extern __shared__ double r1[];  // rdm1_C[(p,q)] = <bra|p^+ q|ket>
  const int pair = blockIdx.x;
  const int n = norb, n2 = n * n;
  const double* base = partial + static_cast<int64_t>(pair) * nsplit * T * 256;
  const int64_t item_stride = static_cast<int64_t>(T) * 256;
  auto fetch = [&](int x, int y) -> double {  // requires blk(x) >= blk(y)
    const int bx = x >> 4, by = y >> 4;
    const int off = (bx * (bx + 1) / 2 + by) * 256 + ((((x >> 3) & 1) << 1) | ((y >> 3) & 1)) * 64 +
                    (x & 7) * 8 + (y & 7);
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += base[k * item_stride + off];
    return s;
  };
  for (int y = threadIdx.x; y <= n2; y += blockDim.x) {
    const double v = fetch(n2, y);
    if (y < n2) r1[y] = v; else ovlp[pair * ovlp_stride] = v;
  }
  __syncthreads();
  double* d1 = dm1 + static_cast<int64_t>(pair) * dm1_stride;
  for (int k = threadIdx.x; k < n2; k += blockDim.x) {
    const int p = k / n, q = k - p * n;
    d1[k] = r1[q * n + p];  // dm1[p,q] = <bra|q^+ p|ket>
  }
  double* d2 = dm2 + static_cast<int64_t>(pair) * dm2_stride;
  const int64_t tot = static_cast<int64_t>(n2) * n2;
  for (int64_t k = threadIdx.x; k < tot; k += blockDim.x) {
    const int x = static_cast<int>(k / n2), y = static_cast<int>(k - static_cast<int64_t>(x) * n2);
    const int p = x / n, q = x - p * n, r = y / n, s = y - r * n;
    double v;
    if ((x >> 4) >= (y >> 4)) {
      v = fetch(x, y);
    } else {
      // E_pq E_rs = E_rs E_pq + d_qr E_ps - d_ps E_rq
      v = fetch(y, x);
      if (q == r) v += r1[p * n + s];
      if (p == s) v -= r1[r * n + q];
    }
    if (q == r) v -= r1[p * n + s];  // reorder: p^+ q r^+ s = d_qr p^+ s + p^+ r^+ s q
    d2[k] = v;
  }
}

// Rows [dm2 (n^4) | dm1 (n^2) | ovlp] of the pairs (a, b), a >= b, as the stack build gathers them from the
// ranks -> the reference's (N, N, ...) arrays: block [a, b] and the same, untransposed, block at [b, a]
// (evcont/FCI_EVCont.py:124-127).  Rows that repeat a pair (slab padding) rewrite the same values.
__global__ void stack_scatter_rows_kernel(int N, int norb, const double* __restrict__ rows, int64_t row_stride,
                                          const int32_t* __restrict__ row_pairs, double* __restrict__ overlap,
                                          double* __restrict__ one_rdm, double* __restrict__ two_rdm) {
  const int r = blockIdx.y, a = row_pairs[2 * r], b = row_pairs[2 * r + 1];
  const int64_t n2 = static_cast<int64_t>(norb) * norb, n4 = n2 * n2, len = n4 + n2 + 1;
  const double* src = rows + static_cast<int64_t>(r) * row_stride;
  const int64_t ab = static_cast<int64_t>(a) * N + b, ba = static_cast<int64_t>(b) * N + a;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < len;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double v = src[k];
    if (k < n4) {
      two_rdm[ab * n4 + k] = v;
      two_rdm[ba * n4 + k] = v;
    } else if (k < n4 + n2) {
      one_rdm[ab * n2 + (k - n4)] = v;
      one_rdm[ba * n2 + (k - n4)] = v;
    } else {
      overlap[ab] = v;
      overlap[ba] = v;
    }
  }
}

struct TrdmPlan {
  int W, nblk, T, nwarps, maxblk, Bt, ntile, nsplit, occupancy;
  size_t smem;
  int pipe;          // producer / consumer kernel (two tile stages, one CTA per SM): 0 off, else the MMA warps / 4
  size_t smem_pipe;
};

int plan_trdm(int norb, int64_t na, int64_t nb, int npairs, int sm_count, TrdmPlan* pl) {
  const int n2 = norb * norb;
  pl->W = (n2 + 1 + 15) / 16 * 16;
  pl->nblk = pl->W / 16;
  pl->T = pl->nblk * (pl->nblk + 1) / 2;
  if (pl->nblk > 11) return -1;
  pl->nwarps = pl->nblk >= 9 ? 16 : 8;
  // 5 macro-block rows (norb = 8): 16 warps with one macro-block each fit 64 registers, so two such CTAs stay
  // resident and the link phases (latency-bound) have twice the warps: 0.623 -> 0.597 ms for 210 pairs.  With two
  // macro-blocks per warp (norb = 9, 10) the 64-register build spills and loses (11.33 against 10.78 ms at norb = 10).
  const bool wide16 = pl->nblk >= 5 && pl->T <= 16;
  if (wide16) pl->nwarps = 16;
  const int per_smsp = (pl->T + 3) / 4;
  const int warps_per_smsp = pl->nwarps / 4;
  pl->maxblk = (per_smsp + warps_per_smsp - 1) / warps_per_smsp;
  // beta tile: Bt % 8 == 4; prefer two CTAs per SM, little padding waste.
  const size_t smem_cap = 227 * 1024;
  const int nbpad = (static_cast<int>(nb) + 1) & ~1;
  const size_t crow_bytes = 2 * static_cast<size_t>(nbpad) * 8;
  const int want_occ = (pl->nwarps == 8 || wide16) ? 2 : 1;
  const size_t budget = smem_cap / want_occ - 1024;  // 1 KB reserved per CTA
  int best = -1;
  double best_cost = 1e300;
  for (int bt = 12; bt <= 100; bt += 8) {
    const size_t bytes = 2 * static_cast<size_t>(pl->W) * bt * 8 + crow_bytes;
    if (bytes > budget) break;
    const int nt = static_cast<int>((nb + bt - 1) / bt);
    // cost model: padded MMA work + a fixed per-tile overhead worth ~8 columns
    const double cost = static_cast<double>(nt) * (bt + 8);
    if (cost < best_cost) { best_cost = cost; best = bt; }
  }
  if (best < 0) {
    // fall back to one CTA per SM
    for (int bt = 12; bt <= 100; bt += 8) {
      const size_t bytes = 2 * static_cast<size_t>(pl->W) * bt * 8 + crow_bytes;
      if (bytes > smem_cap - 1024) break;
      const int nt = static_cast<int>((nb + bt - 1) / bt);
      const double cost = static_cast<double>(nt) * (bt + 8);
      if (cost < best_cost) { best_cost = cost; best = bt; }
    }
  }
  if (best < 0) return -2;
  {
    // development aid: EVC_TRDM_BT forces the tile width (must be 4 mod 8 and fit); the results do not depend on it
    // (the k order is the order of the beta strings whatever the tile boundaries, the padding adds exact zeros)
    static const int force_bt = [] { const char* e = getenv("EVC_TRDM_BT"); return e ? atoi(e) : 0; }();
    if (force_bt >= 12 && force_bt % 8 == 4 &&
        2 * static_cast<size_t>(pl->W) * force_bt * 8 + crow_bytes <= (want_occ == 2 ? budget : smem_cap - 1024))
      best = force_bt;
  }
  pl->Bt = best;
  pl->ntile = static_cast<int>((nb + best - 1) / best);
  pl->smem = 2 * static_cast<size_t>(pl->W) * best * 8 + crow_bytes;
  pl->occupancy = static_cast<int>(smem_cap / (pl->smem + 1024));
  if (pl->occupancy > want_occ) pl->occupancy = want_occ;
  if (pl->occupancy < 1) pl->occupancy = 1;
  // alpha slices: aim for >= 6 waves of CTAs, at least 2 alpha strings per slice
  const int64_t resident = static_cast<int64_t>(sm_count) * pl->occupancy;
  int64_t s = (6 * resident + npairs - 1) / npairs;
  if (s > na / 2) s = na / 2;
  if (s < 1) s = 1;
  if (s > 4096) s = 4096;
  pl->nsplit = static_cast<int>(s);
  // the producer / consumer kernel keeps the tile width and the alpha slices of the fused plan (same bits)
  // EVC_TRDM_PIPE: 0 fused kernel, 1 eight MMA warps, 2 (default) twelve: a warp issues one DMMA per ~17-20 clocks and
  // stops for its fragment loads, so three warps per SM sub-partition keep the pipe fed where two just did
  // (norb = 11: 7.17 -> 7.01 ms, H10 8.33 -> 8.31 ms: there the builders are the slower side)
  static const int use_pipe = [] { const char* e = getenv("EVC_TRDM_PIPE"); return e ? atoi(e) : 2; }();
  pl->smem_pipe = 2 * (pl->smem - crow_bytes) + crow_bytes;
  pl->pipe = (use_pipe && pl->nwarps == 8 && pl->nblk >= 7 && pl->smem_pipe + 1024 <= smem_cap) ? (use_pipe == 2 ? 3 : 2) : 0;
  return 0;
}

void assign_blocks(const TrdmPlan& pl, TrdmParams* P) {
  // spread the T macro-blocks evenly over the 4 SM sub-partitions (warp % 4),
  // then over the warps of each sub-partition; contiguous ranges per warp.
  int per_smsp[4];
  for (int q = 0; q < 4; ++q) per_smsp[q] = pl.T / 4 + (q < pl.T % 4 ? 1 : 0);
  const int wps = pl.nwarps / 4;
  int count[16] = {0};
  for (int q = 0; q < 4; ++q)
    for (int j = 0; j < wps; ++j) count[j * 4 + q] = per_smsp[q] / wps + (j < per_smsp[q] % wps ? 1 : 0);
  int start = 0;
  for (int w = 0; w < 16; ++w) {
    P->blk_start[w] = static_cast<unsigned char>(w < pl.nwarps ? start : 0);
    P->blk_count[w] = static_cast<unsigned char>(w < pl.nwarps ? count[w] : 0);
    if (w < pl.nwarps) start += count[w];
  }
}

template <int NMMA, int MAXBLK>
int launch_pipe(const TrdmParams& P, const TrdmPlan& pl, int nitems, cudaStream_t stream) {
  auto kern = trdm_pipe_kernel<NMMA, MAXBLK>;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(pl.smem_pipe)));
  kern<<<nitems, (NMMA + kPipeBuild) * 32, pl.smem_pipe, stream>>>(P);
  EVC_CHECK_LAUNCH();
  return 0;
}

template <int NWARPS, int MAXBLK, int MINCTA = (NWARPS == 8 ? 2 : 1)>
int launch_fused(const TrdmParams& P, const TrdmPlan& pl, int nitems, cudaStream_t stream) {
  auto kern = trdm_fused_kernel<NWARPS, MAXBLK, MINCTA>;
  EVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(pl.smem)));
  kern<<<nitems, NWARPS * 32, pl.smem, stream>>>(P);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" {

int evc_trans_rdm12_workspace_bytes(int norb, int64_t na, int64_t nb, int npairs, int sm_count,
                                    size_t* bytes) {
  EVC_REQUIRE(bytes != nullptr, "evc_trans_rdm12_workspace_bytes: bytes is NULL");
  EVC_REQUIRE(norb >= 1 && norb <= 13, "evc_trans_rdm12: norb=%d unsupported (1..13)", norb);
  EVC_REQUIRE(npairs >= 1 && na >= 1 && nb >= 1, "evc_trans_rdm12: empty problem");
  TrdmPlan pl;
  EVC_REQUIRE(plan_trdm(norb, na, nb, npairs, sm_count > 0 ? sm_count : 148, &pl) == 0,
              "evc_trans_rdm12: no launch plan for norb=%d nb=%lld", norb, (long long)nb);
  *bytes = static_cast<size_t>(npairs) * pl.nsplit * pl.T * 256 * sizeof(double);
  return 0;
}

int evc_trans_rdm12_batch_strided(evc_ctx* ctx, int norb, int64_t na, int64_t nb, const double* civecs,
                                  int64_t vec_stride, int nvec, const int32_t* pairs, int npairs,
                                  const uint64_t* link_a, int nlink_a, const uint64_t* link_b, int nlink_b,
                                  double* ovlp, int64_t ovlp_stride, double* dm1, int64_t dm1_stride, double* dm2,
                                  int64_t dm2_stride, void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx != nullptr, "evc_trans_rdm12_batch: ctx is NULL");
  EVC_REQUIRE(norb >= 1 && norb <= 13, "evc_trans_rdm12_batch: norb=%d unsupported (1..13)", norb);
  EVC_REQUIRE(civecs && pairs && link_a && link_b && ovlp && dm1 && dm2 && workspace,
              "evc_trans_rdm12_batch: NULL pointer argument");
  EVC_REQUIRE(npairs >= 1 && nvec >= 1 && na >= 1 && nb >= 1, "evc_trans_rdm12_batch: empty problem");
  EVC_REQUIRE(vec_stride >= na * nb, "evc_trans_rdm12_batch: vec_stride %lld < na*nb", (long long)vec_stride);
  EVC_REQUIRE((reinterpret_cast<uintptr_t>(civecs) & 15) == 0 && (vec_stride & 1) == 0,
              "evc_trans_rdm12_batch: civecs must be 16-byte aligned with an even stride");
  TrdmPlan pl;
  // the alpha-slice count fixes the summation order of a pair's result: planned for the pair count of the WHOLE
  // build when the caller computes only a share of it (evc_trans_rdm12_plan_pairs), so that a pair's bits do not
  // depend on how many ranks share the build
  const int plan_pairs = ctx->trdm_plan_pairs > npairs ? ctx->trdm_plan_pairs : npairs;
  EVC_REQUIRE(plan_trdm(norb, na, nb, plan_pairs, ctx->sm_count, &pl) == 0,
              "evc_trans_rdm12_batch: no launch plan for norb=%d nb=%lld", norb, (long long)nb);
  const size_t need = static_cast<size_t>(npairs) * pl.nsplit * pl.T * 256 * sizeof(double);
  EVC_REQUIRE(workspace_bytes >= need, "evc_trans_rdm12_batch: workspace %zu < %zu bytes", workspace_bytes, need);
  EVC_REQUIRE(pl.smem <= ctx->smem_optin, "evc_trans_rdm12_batch: needs %zu B shared memory, device allows %zu",
              pl.smem, ctx->smem_optin);

  TrdmParams P;
  P.norb = norb; P.n2 = norb * norb; P.W = pl.W; P.na = na; P.nb = nb;
  P.civecs = civecs; P.vec_stride = vec_stride; P.pairs = pairs; P.nsplit = pl.nsplit;
  P.link_a = link_a; P.nlink_a = nlink_a; P.link_b = link_b; P.nlink_b = nlink_b;
  P.Bt = pl.Bt; P.ntile = pl.ntile; P.partial = static_cast<double*>(workspace); P.T = pl.T;
  const int nitems = npairs * pl.nsplit;
  int rc = -1;
  if (pl.pipe && pl.smem_pipe <= ctx->smem_optin) {
    TrdmPlan pp = pl;  // macro-blocks dealt to the MMA warps of the producer / consumer kernel
    pp.nwarps = 4 * pl.pipe;
    pp.maxblk = ((pl.T + 3) / 4 + pl.pipe - 1) / pl.pipe;
    assign_blocks(pp, &P);
    if (pl.pipe == 2) {
      switch (pp.maxblk) {
        case 3: rc = launch_pipe<8, 3>(P, pl, nitems, ctx->stream); break;
        case 4: rc = launch_pipe<8, 4>(P, pl, nitems, ctx->stream); break;
        case 5: rc = launch_pipe<8, 5>(P, pl, nitems, ctx->stream); break;
        default: break;
      }
    } else {
      switch (pp.maxblk) {
        case 2: rc = launch_pipe<12, 2>(P, pl, nitems, ctx->stream); break;
        case 3: rc = launch_pipe<12, 3>(P, pl, nitems, ctx->stream); break;
        default: break;
      }
    }
  }
  if (rc == -1) assign_blocks(pl, &P);
  if (rc != -1) {
    // launched above
  } else if (pl.nwarps == 8) {
    switch (pl.maxblk) {
      case 1: rc = launch_fused<8, 1>(P, pl, nitems, ctx->stream); break;
      case 2: rc = launch_fused<8, 2>(P, pl, nitems, ctx->stream); break;
      case 3: rc = launch_fused<8, 3>(P, pl, nitems, ctx->stream); break;
      case 4: rc = launch_fused<8, 4>(P, pl, nitems, ctx->stream); break;
      case 5: rc = launch_fused<8, 5>(P, pl, nitems, ctx->stream); break;
      default: break;
    }
  } else {
    switch (pl.maxblk) {
      case 1: rc = launch_fused<16, 1, 2>(P, pl, nitems, ctx->stream); break;
      case 3: rc = launch_fused<16, 3>(P, pl, nitems, ctx->stream); break;
      case 4: rc = launch_fused<16, 4>(P, pl, nitems, ctx->stream); break;
      case 5: rc = launch_fused<16, 5>(P, pl, nitems, ctx->stream); break;
      default: break;
    }
  }
  EVC_REQUIRE(rc != -1, "evc_trans_rdm12_batch: no kernel instance for nwarps=%d maxblk=%d", pl.nwarps, pl.maxblk);
  if (rc != 0) return rc;
  EVC_REQUIRE(ovlp_stride >= 1 && dm1_stride >= P.n2 && dm2_stride >= static_cast<int64_t>(P.n2) * P.n2,
              "evc_trans_rdm12_batch: output strides too small");
  trdm_finalize_kernel<<<npairs, 256, P.n2 * sizeof(double), ctx->stream>>>(
      norb, pl.T, pl.nsplit, P.partial, ovlp, ovlp_stride, dm1, dm1_stride, dm2, dm2_stride);
  EVC_CHECK_LAUNCH();
  ctx->last_trdm_flops = static_cast<double>(npairs) * static_cast<double>(na) * pl.ntile * pl.Bt *
                         static_cast<double>(pl.T) * 256.0 * 2.0;
  return 0;
}

int evc_trans_rdm12_batch(evc_ctx* ctx, int norb, int64_t na, int64_t nb, const double* civecs,
                          int64_t vec_stride, int nvec, const int32_t* pairs, int npairs,
                          const uint64_t* link_a, int nlink_a, const uint64_t* link_b, int nlink_b,
                          double* ovlp, double* dm1, double* dm2, void* workspace,
                          size_t workspace_bytes) {
  const int64_t n2 = static_cast<int64_t>(norb) * norb;
  return evc_trans_rdm12_batch_strided(ctx, norb, na, nb, civecs, vec_stride, nvec, pairs, npairs, link_a, nlink_a,
                                       link_b, nlink_b, ovlp, 1, dm1, n2, dm2, n2 * n2, workspace, workspace_bytes);
}

double evc_trans_rdm12_last_issued_flops(const evc_ctx* ctx) { return ctx ? ctx->last_trdm_flops : 0.0; }

int evc_trans_rdm12_plan_pairs(evc_ctx* ctx, int total_pairs) {
  EVC_REQUIRE(ctx && total_pairs >= 0, "evc_trans_rdm12_plan_pairs: bad argument");
  ctx->trdm_plan_pairs = total_pairs;
  return 0;
}

int64_t evc_stack_row_len(int norb) {
  const int64_t n2 = static_cast<int64_t>(norb) * norb;
  return (n2 * n2 + n2 + 1 + 1) & ~static_cast<int64_t>(1);
}

int evc_stack_scatter_rows(evc_ctx* ctx, int ntrain, int norb, const double* rows, int64_t row_stride, int nrows,
                           const int32_t* row_pairs, double* overlap, double* one_rdm, double* two_rdm) {
  EVC_REQUIRE(ctx && rows && row_pairs && overlap && one_rdm && two_rdm, "evc_stack_scatter_rows: NULL argument");
  EVC_REQUIRE(ntrain >= 1 && norb >= 1 && row_stride >= evc_stack_row_len(norb) - 1, "evc_stack_scatter_rows: bad sizes");
  if (nrows <= 0) return 0;
  const int64_t n2 = static_cast<int64_t>(norb) * norb, len = n2 * n2 + n2 + 1;
  dim3 grid(static_cast<unsigned>(std::min<int64_t>((len + 255) / 256, 1024)), nrows);
  stack_scatter_rows_kernel<<<grid, 256, 0, ctx->stream>>>(ntrain, norb, rows, row_stride, row_pairs, overlap, one_rdm,
                                                           two_rdm);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
