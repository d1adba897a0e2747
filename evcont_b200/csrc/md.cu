// Device-resident velocity-Verlet integrator, batched over independent trajectories
// (SURVEY.md section 8 row f3).  The reference drives pyscf.md.NVE from the host
// (evcont/MD_utils.py:60-125: one scanner call, i.e. one get_energy_with_grad, per step);
// here positions, velocities, accelerations and the recorded frames stay in HBM, so a step
// is "positions kernel -> K9 integrals -> K3..K8 prediction -> velocities kernel" with no
// host round trip and can be captured in a CUDA graph.
//
// pyscf.md.integrators.VelocityVerlet, restated:
//   frame 0 = the initial geometry (only the acceleration is computed),
//   x_{k+1} = x_k + dt v_k + dt^2/2 a_k ;  a_{k+1} = -grad(x_{k+1}) / m ;
//   v_{k+1} = v_k + dt/2 (a_k + a_{k+1}) ;  E_kin = 1/2 sum m v^2.
#include "common.cuh"

namespace {

// one thread per coordinate; frame_idx is a device counter so that a captured graph can be
// replayed step after step
__global__ void md_positions_kernel(int64_t ncoord, double dt, const double* __restrict__ v,
                                    const double* __restrict__ a, double* __restrict__ x) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= ncoord) return;
  x[k] = fma(dt, fma(0.5 * dt, a[k], v[k]), x[k]);
}

// a_new = -grad / m; if `first`: a = a_new only (frame 0), else v += dt/2 (a + a_new), a = a_new.
// One warp per trajectory: also reduces the kinetic energy and records the frame.
__global__ void md_velocities_kernel(int nbatch, int natm, double dt, int first, const double* __restrict__ inv_mass,
                                     const double* __restrict__ mass, const double* __restrict__ grad,
                                     const double* __restrict__ x, const double* __restrict__ epot,
                                     double* __restrict__ v, double* __restrict__ a, double* __restrict__ ekin,
                                     int* __restrict__ frame_idx, int max_frames, double* __restrict__ traj,
                                     double* __restrict__ epot_log, double* __restrict__ ekin_log) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= nbatch) return;
  const int nc = natm * 3;
  const int64_t o = static_cast<int64_t>(g) * nc;
  const int f = *frame_idx;
  double ke = 0.0;
  for (int k = lane; k < nc; k += 32) {
    const int at = k / 3;
    const double an = -grad[o + k] * inv_mass[at];
    double vk = v[o + k];
    if (!first) vk = fma(0.5 * dt, a[o + k] + an, vk);
    v[o + k] = vk;
    a[o + k] = an;
    ke = fma(0.5 * mass[at] * vk, vk, ke);
    if (traj && f < max_frames) traj[(static_cast<int64_t>(f) * nbatch + g) * nc + k] = x[o + k];
  }
  for (int s = 16; s > 0; s >>= 1) ke += __shfl_xor_sync(0xffffffffu, ke, s);
  if (lane == 0) {
    ekin[g] = ke;
    if (f < max_frames) {
      if (epot_log) epot_log[static_cast<int64_t>(f) * nbatch + g] = epot[g];
      if (ekin_log) ekin_log[static_cast<int64_t>(f) * nbatch + g] = ke;
    }
  }
}

__global__ void md_advance_frame_kernel(int* frame_idx) { *frame_idx += 1; }

// Berendsen thermostat as pyscf.md.integrators.NVTBerendson applies it before each step:
// v *= clip(sqrt(1 + (T / T_inst - 1) dt / taut), 0.9, 1.1),  T_inst = 2 E_kin / (3 natm k_B)
__global__ void md_berendsen_kernel(int nbatch, int natm, double dt, double taut, double temperature, double kb,
                                    const double* __restrict__ ekin, double* __restrict__ v) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= nbatch) return;
  const int nc = natm * 3;
  const double tinst = 2.0 * ekin[g] / (static_cast<double>(nc) * kb);
  double f = 1.1;  // T_inst == 0: the ratio diverges and the clip applies
  if (tinst > 0.0) {
    const double a = 1.0 + (temperature / tinst - 1.0) * dt / taut;
    f = a > 0.0 ? sqrt(a) : 0.9;
    f = fmin(1.1, fmax(0.9, f));
  }
  for (int k = lane; k < nc; k += 32) v[static_cast<int64_t>(g) * nc + k] *= f;
}

}  // namespace


// Farthest-point selection in Hamiltonian space (evcont/MD_utils.py:363-405, data_addition =
// "farthest_point_ham"): for every trajectory frame g the smallest weighted squared distance
//   d[g][t] = sum_{k < L1} (x[g][k] - y[t][k])^2 + 1/2 sum_{L1 <= k < L} (x[g][k] - y[t][k])^2
// to the T training rows (x, y: [.][L] rows holding h1 (L1 = n^2) followed by h2 (n^4) in the OAO
// basis).  One CTA per frame; differences are formed before squaring (no cancellation), fixed
// reduction order.
__global__ void __launch_bounds__(256)
md_min_sqdist_kernel(int T, int64_t L1, int64_t L, const double* __restrict__ x, const double* __restrict__ y,
                     double* __restrict__ dmin) {
  __shared__ double red[8];
  const int g = blockIdx.x, tid = threadIdx.x;
  const double* xr = x + static_cast<int64_t>(g) * L;
  double best = 0.0;
  for (int t = 0; t < T; ++t) {
    const double* yr = y + static_cast<int64_t>(t) * L;
    double a1 = 0.0, a2 = 0.0;
    for (int64_t k = tid; k < L; k += 256) {
      const double d = xr[k] - yr[k];
      if (k < L1) a1 = fma(d, d, a1); else a2 = fma(d, d, a2);
    }
    double a = a1 + 0.5 * a2;
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = a;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w];
      if (t == 0 || s < best) best = s;
    }
  }
  if (tid == 0) dmin[g] = best;
}

extern "C" {

int evc_md_positions(evc_ctx* ctx, int nbatch, int natm, double dt, const double* v, const double* a, double* x) {
  EVC_REQUIRE(ctx && v && a && x, "evc_md_positions: NULL argument");
  if (nbatch <= 0) return 0;
  const int64_t nc = static_cast<int64_t>(nbatch) * natm * 3;
  md_positions_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, ctx->stream>>>(nc, dt, v, a, x);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_md_berendsen(evc_ctx* ctx, int nbatch, int natm, double dt, double taut, double temperature,
                     const double* ekin, double* v) {
  EVC_REQUIRE(ctx && ekin && v, "evc_md_berendsen: NULL argument");
  EVC_REQUIRE(taut > 0.0 && temperature >= 0.0, "evc_md_berendsen: taut must be positive, T non-negative");
  if (nbatch <= 0) return 0;
  md_berendsen_kernel<<<(nbatch + 3) / 4, 128, 0, ctx->stream>>>(nbatch, natm, dt, taut, temperature,
                                                                  3.166811563e-6 /* Hartree / K */, ekin, v);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_md_velocities(evc_ctx* ctx, int nbatch, int natm, double dt, int first, const double* inv_mass,
                      const double* mass, const double* grad, const double* x, const double* epot, double* v,
                      double* a, double* ekin, int* frame_idx, int max_frames, double* traj, double* epot_log,
                      double* ekin_log) {
  EVC_REQUIRE(ctx && inv_mass && mass && grad && x && epot && v && a && ekin && frame_idx,
              "evc_md_velocities: NULL argument");
  if (nbatch <= 0) return 0;
  md_velocities_kernel<<<(nbatch + 3) / 4, 128, 0, ctx->stream>>>(nbatch, natm, dt, first, inv_mass, mass, grad, x,
                                                                   epot, v, a, ekin, frame_idx, max_frames, traj,
                                                                   epot_log, ekin_log);
  EVC_CHECK_LAUNCH();
  md_advance_frame_kernel<<<1, 1, 0, ctx->stream>>>(frame_idx);
  EVC_CHECK_LAUNCH();
  return 0;
}

int evc_min_sqdist(evc_ctx* ctx, int nframes, int ntrain, int64_t len_one, int64_t len_total, const double* frames,
                   const double* train, double* dmin) {
  EVC_REQUIRE(ctx && frames && train && dmin, "evc_min_sqdist: NULL argument");
  EVC_REQUIRE(ntrain >= 1 && len_one >= 0 && len_total >= len_one, "evc_min_sqdist: bad sizes");
  if (nframes <= 0) return 0;
  md_min_sqdist_kernel<<<nframes, 256, 0, ctx->stream>>>(ntrain, len_one, len_total, frames, train, dmin);
  EVC_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
