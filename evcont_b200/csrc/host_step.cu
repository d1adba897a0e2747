// The prediction step with HOST input/output buffers: the reference-facing form of
// get_energy_with_grad (evcont/ab_initio_gradients_loewdin.py:308-379), where the AO
// arrays of every geometry arrive from the host (libcint output in the reference).
//
// The batch is cut into chunks; chunk c+1 is copied host->device on a copy stream
// while chunk c runs through K3..K8 on the ctx stream and the results of chunk c-1
// return on a third stream (double-buffered staging slots, CUDA events only -- no
// host synchronisation inside).  With pinned host memory the whole call is
// asynchronous with respect to the host; pageable memory works but serialises.
#include "packed.cuh"

namespace {

struct AoSizes {
  size_t ovlp, hcore, eri, ipovlp, hcore_deriv, eri_ip1, e_nuc, grad_nuc;  // doubles per geometry
};

AoSizes ao_sizes(int n, int natm) {
  const size_t n2 = static_cast<size_t>(n) * n, n4 = n2 * n2;
  return AoSizes{n2, n2, n4, 3 * n2, static_cast<size_t>(natm) * 3 * n2, 3 * n4, 1,
                 static_cast<size_t>(natm) * 3};
}

struct Slot {
  double *ovlp, *hcore, *eri, *ipovlp, *hcore_deriv, *eri_ip1, *e_nuc, *grad_nuc, *E, *grad;
};

bool carve_slot(evc_arena& ar, const AoSizes& sz, int chunk, int natm, Slot* s) {
  s->ovlp = ar.take<double>(chunk * sz.ovlp);
  s->hcore = ar.take<double>(chunk * sz.hcore);
  s->eri = ar.take<double>(chunk * sz.eri);
  s->ipovlp = ar.take<double>(chunk * sz.ipovlp);
  s->hcore_deriv = ar.take<double>(chunk * sz.hcore_deriv);
  s->eri_ip1 = ar.take<double>(chunk * sz.eri_ip1);
  s->e_nuc = ar.take<double>(chunk * sz.e_nuc);
  s->grad_nuc = ar.take<double>(chunk * sz.grad_nuc);
  s->E = ar.take<double>(chunk);
  s->grad = ar.take<double>(static_cast<size_t>(chunk) * natm * 3);
  return s->ovlp && s->hcore && s->eri && s->ipovlp && s->hcore_deriv && s->eri_ip1 && s->e_nuc &&
         s->grad_nuc && s->E && s->grad;
}

size_t slot_bytes(const AoSizes& sz, int chunk, int natm) {
  const size_t c = static_cast<size_t>(chunk);
  size_t t = 0;
  for (size_t per : {sz.ovlp, sz.hcore, sz.eri, sz.ipovlp, sz.hcore_deriv, sz.eri_ip1, sz.e_nuc, sz.grad_nuc,
                     static_cast<size_t>(1), static_cast<size_t>(natm) * 3})
    t += evc_align_up(c * per * sizeof(double), 256);
  return t;
}

int ensure_pipeline(evc_ctx* ctx) {
  if (ctx->pipe_ready) return 0;
  EVC_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  EVC_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) {
    EVC_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming));
    EVC_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_compute[k], cudaEventDisableTiming));
    EVC_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_d2h[k], cudaEventDisableTiming));
  }
  EVC_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
  ctx->pipe_ready = 1;
  return 0;
}

}  // namespace

extern "C" {

int evc_energy_with_grad_host_workspace_bytes(int layout, int N, int n, int natm, int chunk, size_t* bytes) {
  EVC_REQUIRE(bytes != nullptr && chunk >= 1, "evc_energy_with_grad_host_workspace_bytes: bad arguments");
  size_t step = 0;
  int rc = evc_energy_with_grad_workspace_bytes(layout, N, n, natm, chunk, &step);
  if (rc) return rc;
  *bytes = evc_align_up(step, 256) + 2 * slot_bytes(ao_sizes(n, natm), chunk, natm) +
           evc_align_up(static_cast<size_t>(natm) * 2 * sizeof(int32_t), 256);
  return 0;
}

}  // extern "C"

namespace {

// step(dev bundle, count, E, grad, step workspace) enqueues the prediction step of
// one chunk on the ctx stream
template <typename Step>
int host_pipeline(evc_ctx* ctx, int n, int natm, int nbatch, const evc_ao_bundle* ao_host, double* E_host,
                  double* grad_host, int chunk, size_t step_b, void* workspace, size_t workspace_bytes,
                  Step step) {
  int rc = ensure_pipeline(ctx);
  if (rc) return rc;
  const AoSizes sz = ao_sizes(n, natm);
  evc_arena ar(workspace, workspace_bytes);
  char* step_ws = ar.take<char>(step_b);
  int32_t* aosl = ar.take<int32_t>(static_cast<size_t>(natm) * 2);
  Slot slot[2];
  const bool ok = step_ws && aosl && carve_slot(ar, sz, chunk, natm, &slot[0]) && carve_slot(ar, sz, chunk, natm, &slot[1]);
  EVC_REQUIRE(ok, "evc_energy_with_grad_host: workspace too small (%zu bytes)", workspace_bytes);

  cudaStream_t cs = ctx->stream, hs = ctx->h2d_stream, ds = ctx->d2h_stream;
  // the copy streams start after everything already queued on the ctx stream
  EVC_CHECK_CUDA(cudaEventRecord(ctx->ev_start, cs));
  EVC_CHECK_CUDA(cudaStreamWaitEvent(hs, ctx->ev_start, 0));
  EVC_CHECK_CUDA(cudaStreamWaitEvent(ds, ctx->ev_start, 0));
  EVC_CHECK_CUDA(cudaMemcpyAsync(aosl, ao_host->aoslices, static_cast<size_t>(natm) * 2 * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, hs));
  const int nchunk = (nbatch + chunk - 1) / chunk;
  for (int c = 0; c < nchunk; ++c) {
    const int k = c & 1;
    const size_t g0 = static_cast<size_t>(c) * chunk;
    const size_t cnt = static_cast<size_t>(c == nchunk - 1 ? nbatch - c * chunk : chunk);
    Slot& s = slot[k];
    // inputs of this slot are free once the compute of chunk c-2 is done
    if (c >= 2) EVC_CHECK_CUDA(cudaStreamWaitEvent(hs, ctx->ev_compute[k], 0));
#define EVC_H2D(field, per)                                                                       \
  EVC_CHECK_CUDA(cudaMemcpyAsync(s.field, ao_host->field + g0 * (per), cnt * (per) * sizeof(double), \
                                 cudaMemcpyHostToDevice, hs))
    EVC_H2D(ovlp, sz.ovlp);
    EVC_H2D(hcore, sz.hcore);
    // two-electron arrays: packed (erip / eri_ip1p, into the same slot buffers) or full tensors
    const size_t ne = static_cast<size_t>(evcp::erip_len(n)), ni = static_cast<size_t>(evcp::ip1p_len(n));
    if (ao_host->erip)
      EVC_CHECK_CUDA(cudaMemcpyAsync(s.eri, ao_host->erip + g0 * ne, cnt * ne * sizeof(double), cudaMemcpyHostToDevice, hs));
    else
      EVC_H2D(eri, sz.eri);
    EVC_H2D(ipovlp, sz.ipovlp);
    EVC_H2D(hcore_deriv, sz.hcore_deriv);
    if (ao_host->eri_ip1p)
      EVC_CHECK_CUDA(cudaMemcpyAsync(s.eri_ip1, ao_host->eri_ip1p + g0 * ni, cnt * ni * sizeof(double), cudaMemcpyHostToDevice, hs));
    else
      EVC_H2D(eri_ip1, sz.eri_ip1);
    if (ao_host->e_nuc) EVC_H2D(e_nuc, sz.e_nuc);
    if (ao_host->grad_nuc) EVC_H2D(grad_nuc, sz.grad_nuc);
#undef EVC_H2D
    EVC_CHECK_CUDA(cudaEventRecord(ctx->ev_h2d[k], hs));
    EVC_CHECK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_h2d[k], 0));
    // outputs of this slot are free once the read-back of chunk c-2 is done
    if (c >= 2) EVC_CHECK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_d2h[k], 0));
    evc_ao_bundle dev;
    dev.ovlp = s.ovlp; dev.hcore = s.hcore; dev.ipovlp = s.ipovlp; dev.hcore_deriv = s.hcore_deriv;
    dev.eri = ao_host->erip ? nullptr : s.eri;
    dev.erip = ao_host->erip ? s.eri : nullptr;
    dev.eri_ip1 = ao_host->eri_ip1p ? nullptr : s.eri_ip1;
    dev.eri_ip1p = ao_host->eri_ip1p ? s.eri_ip1 : nullptr;
    dev.e_nuc = ao_host->e_nuc ? s.e_nuc : nullptr;
    dev.grad_nuc = ao_host->grad_nuc ? s.grad_nuc : nullptr;
    dev.aoslices = aosl;
    if ((rc = step(&dev, static_cast<int>(cnt), s.E, s.grad, step_ws))) return rc;
    EVC_CHECK_CUDA(cudaEventRecord(ctx->ev_compute[k], cs));
    EVC_CHECK_CUDA(cudaStreamWaitEvent(ds, ctx->ev_compute[k], 0));
    EVC_CHECK_CUDA(cudaMemcpyAsync(E_host + g0, s.E, cnt * sizeof(double), cudaMemcpyDeviceToHost, ds));
    EVC_CHECK_CUDA(cudaMemcpyAsync(grad_host + g0 * natm * 3, s.grad, cnt * natm * 3 * sizeof(double),
                                   cudaMemcpyDeviceToHost, ds));
    EVC_CHECK_CUDA(cudaEventRecord(ctx->ev_d2h[k], ds));
  }
  // a caller that synchronises the ctx stream sees the results on the host
  EVC_CHECK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_d2h[0], 0));
  if (nchunk > 1) EVC_CHECK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_d2h[1], 0));
  return 0;
}

bool host_bundle_ok(const evc_ao_bundle* a, bool packed_ok) {
  return a->ovlp && a->hcore && (a->eri || (packed_ok && a->erip)) && a->ipovlp && a->hcore_deriv &&
         (a->eri_ip1 || (packed_ok && a->eri_ip1p)) && a->aoslices;
}

}  // namespace

extern "C" {

int evc_energy_with_grad_host(evc_ctx* ctx, int layout, int N, int n, int natm, const double* one_rdm,
                              const double* two_rdm, const double* Linv, int nbatch,
                              const evc_ao_bundle* ao_host, double* E_host, double* grad_host, int chunk,
                              void* workspace, size_t workspace_bytes) {
  EVC_REQUIRE(ctx && one_rdm && two_rdm && Linv && ao_host && E_host && grad_host && workspace,
              "evc_energy_with_grad_host: NULL argument");
  EVC_REQUIRE(host_bundle_ok(ao_host, false), "evc_energy_with_grad_host: incomplete AO bundle (full tensors needed)");
  EVC_REQUIRE(chunk >= 1, "evc_energy_with_grad_host: chunk must be >= 1");
  if (nbatch <= 0) return 0;
  if (chunk > nbatch) chunk = nbatch;
  size_t step_b = 0;
  int rc = evc_energy_with_grad_workspace_bytes(layout, N, n, natm, chunk, &step_b);
  if (rc) return rc;
  return host_pipeline(ctx, n, natm, nbatch, ao_host, E_host, grad_host, chunk, step_b, workspace, workspace_bytes,
                       [&](const evc_ao_bundle* dev, int cnt, double* E, double* grad, void* ws) {
                         return evc_energy_with_grad(ctx, layout, N, n, natm, one_rdm, two_rdm, Linv, cnt, dev, E,
                                                     grad, nullptr, nullptr, nullptr, ws, step_b);
                       });
}

int evc_energy_with_grad_packed_host_workspace_bytes(int N, int n, int natm, int chunk, size_t* bytes) {
  EVC_REQUIRE(bytes != nullptr && chunk >= 1, "evc_energy_with_grad_packed_host_workspace_bytes: bad arguments");
  size_t step = 0;
  int rc = evc_energy_with_grad_packed_workspace_bytes(N, n, natm, chunk, &step);
  if (rc) return rc;
  *bytes = evc_align_up(step, 256) + 2 * slot_bytes(ao_sizes(n, natm), chunk, natm) +
           evc_align_up(static_cast<size_t>(natm) * 2 * sizeof(int32_t), 256);
  return 0;
}

int evc_energy_with_grad_packed_host(evc_ctx* ctx, int N, int n, int natm, const double* RH, const double* RG,
                                     const double* Linv, int nbatch, const evc_ao_bundle* ao_host,
                                     double* E_host, double* grad_host, int chunk, void* workspace,
                                     size_t workspace_bytes) {
  EVC_REQUIRE(ctx && RH && RG && Linv && ao_host && E_host && grad_host && workspace,
              "evc_energy_with_grad_packed_host: NULL argument");
  EVC_REQUIRE(host_bundle_ok(ao_host, n <= kPackedMaxNorb), "evc_energy_with_grad_packed_host: incomplete AO bundle");
  EVC_REQUIRE(chunk >= 1, "evc_energy_with_grad_packed_host: chunk must be >= 1");
  if (nbatch <= 0) return 0;
  if (chunk > nbatch) chunk = nbatch;
  size_t step_b = 0;
  int rc = evc_energy_with_grad_packed_workspace_bytes(N, n, natm, chunk, &step_b);
  if (rc) return rc;
  return host_pipeline(ctx, n, natm, nbatch, ao_host, E_host, grad_host, chunk, step_b, workspace, workspace_bytes,
                       [&](const evc_ao_bundle* dev, int cnt, double* E, double* grad, void* ws) {
                         return evc_energy_with_grad_packed(ctx, N, n, natm, RH, RG, Linv, cnt, dev, E, grad,
                                                            nullptr, ws, step_b);
                       });
}

}  // extern "C"
