// Helpers shared by the translation units of the packed prediction step
// (packed.cu: stack packing, per-geometry kernels for 11..13 orbitals, streaming gradient;
//  packed_pipe.cu: persistent warp-specialised K4p / K8a for <= 10 orbitals).
#pragma once
#include "common.cuh"

namespace evcp {

__host__ __device__ inline int tri_idx(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j
__host__ __device__ inline int npair_of(int n) { return n * (n + 1) / 2; }
// row pitch of W (K8a -> K8b): np rounded up to even, so that the accumulator pairs of a DMMA tile are 16-byte
// stores (np = 55 at n = 10: with 8-byte stores every 32-byte sector was written half full and the W store took a
// tenth of K8a's period, profiles/r02_pipe_clock_stamps.txt)
__host__ __device__ inline int w_pitch_of(int n) { return (npair_of(n) + 1) & ~1; }
__host__ __device__ inline int64_t packed_len(int n) {
  const int64_t np = npair_of(n);
  const int64_t l = static_cast<int64_t>(n) * n + np * (np + 1) / 2;
  return (l + 1) & ~static_cast<int64_t>(1);
}

__device__ __forceinline__ void tril_unrank_i(int t, int& a, int& b) {
  int x = static_cast<int>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (x * (x + 1) / 2 > t) --x;
  while ((x + 1) * (x + 2) / 2 <= t) ++x;
  a = x;
  b = t - x * (x + 1) / 2;
}

// ---------------------------------------------------------------------------
// shared-memory geometry of the per-geometry GEMM operands
// ---------------------------------------------------------------------------
struct PGeom {
  int n, np, M8, rows8, K4, pA, pB;
  size_t szA;  // doubles of an A-type image  [rows8][pA]
  size_t szB;  // doubles of a B-type image   [K4][pB]
};

__host__ __device__ inline PGeom pgeom(int n) {
  PGeom g;
  g.n = n;
  g.np = npair_of(n);
  g.M8 = (g.np + 7) / 8;
  g.rows8 = g.M8 * 8;
  g.K4 = (g.np + 3) & ~3;
  // 8-byte shared-memory loads are served one half-warp (lanes 0-15: g = 0..3, tg = 0..3) at a
  // time over 16 double-wide banks.  A-type access (lane (g,tg) reads [row g][col tg], word
  // g*pitch + tg) and B-type access ([row tg][col g], word tg*pitch + g) are both conflict-free
  // when pitch == 4 (mod 8): the four row offsets land on banks {0,4,8,12}.
  int pa = g.K4;
  while ((pa & 7) != 4) ++pa;
  int pb = g.rows8;
  while ((pb & 7) != 4) ++pb;
  g.pA = pa;
  g.pB = pb;
  g.szA = static_cast<size_t>(g.rows8) * pa;
  g.szB = static_cast<size_t>(g.K4) * pb;
  return g;
}

// Packed AO two-electron arrays (include/evcont_b200.h, "packed AO arrays"):
//   erip     [np][pA]        (ab|cd), a >= b, c >= d; row pitch pA = pgeom(n).pA (the A-type
//                            shared-memory pitch, so a geometry's array is ONE bulk copy)
//   eri_ip1p [3][n][n][np]   (d_x m b|c d), c >= d
__host__ __device__ inline int64_t erip_len(int n) {
  const PGeom g = pgeom(n);
  return static_cast<int64_t>(g.np) * g.pA;
}
__host__ __device__ inline int64_t ip1p_len(int n) { return static_cast<int64_t>(3) * n * n * npair_of(n); }

// ---------------------------------------------------------------------------
// asynchronous copies, mbarriers
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit_all() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (bulk copies complete on them)
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// one warp's arrival: every lane's earlier shared-memory accesses are ordered before it
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// 1-D bulk copy global -> shared (TMA unit, SASS UBLKCP); dst/src 16-byte aligned, bytes % 16 == 0;
// completion is signalled on `bar` as `bytes` transaction bytes
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// orders earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// named barrier among `count` threads (count % 32 == 0); id 0 is __syncthreads'
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}

}  // namespace evcp
