// Layer 2 of the boundary: every compute entry point of include/auxssm.h as an XLA FFI handler, so
// that the reference's Python/JAX host code can call the kernels with jax.ffi.ffi_call (route A).
//
// Built ONLY where jaxlib's headers exist (python -m aux_ssm_samplers_b200.build --xla-ffi: uses
// jax.ffi.include_dir()); JAX is not installable in the build image, where this file is syntax-checked against the API
// mock in tests/xla_ffi_mock and tests/test_xla_ffi.py verifies that every C entry point has a handler.
//
// Conventions of every handler
//  * operands and results are dense row-major XLA buffers; a leading chain axis (what jax.vmap with
//    vmap_method="broadcast_all" produces) is optional on every operand, and axes of extent 1 broadcast (stride 0);
//  * dims that cannot be read off the buffers and all flags are int64 attributes;
//  * scratch comes from ffi::ScratchAllocator (released by XLA when the call returns -- the kernels are stream-ordered
//    and XLA keeps the allocation alive until the stream reaches that point);
//  * a non-zero status becomes ffi::Error(kInvalidArgument / kInternal, auxssm_last_error()).
//  * the dtype is taken from the first operand (F32 / F64 select the _f32 / _f64 entry point).
#include <cstdint>
#include <string>
#include <vector>

#include "auxssm.h"
#include "xla/ffi/api/ffi.h"

#ifndef AUXSSM_FFI_STREAM_T
#include <cuda_runtime_api.h>
#define AUXSSM_FFI_STREAM_T cudaStream_t
#endif

namespace ffi = xla::ffi;

namespace {

using Stream = AUXSSM_FFI_STREAM_T;

ffi::Error Status(int32_t rc) {
  if (rc == AUXSSM_OK) return ffi::Error::Success();
  return ffi::Error(rc == AUXSSM_ERR_LAUNCH ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument,
                    std::string(auxssm_last_error()));
}

struct Call {
  Stream stream;
  ffi::ScratchAllocator* scratch;
  ffi::RemainingArgs* args;
  ffi::RemainingRets* rets;
  ffi::Dictionary* attrs;
  bool f64 = false;
  std::string err;

  ffi::AnyBuffer arg(size_t i) {
    auto b = args->get<ffi::AnyBuffer>(i);
    if (!b.has_value()) { err = "missing operand " + std::to_string(i); return ffi::AnyBuffer(); }
    return b.value();
  }
  void* in(size_t i) { return i < args->size() ? arg(i).untyped_data() : nullptr; }
  // an operand that may be omitted by passing an empty (0-element) buffer: H = I, c = 0, NULL draw arrays ...
  void* opt(size_t i) {
    if (i >= args->size()) return nullptr;
    ffi::AnyBuffer b = arg(i);
    return b.element_count() == 0 ? nullptr : b.untyped_data();
  }
  void* out(size_t i) {
    auto r = rets->get<ffi::AnyBuffer>(i);
    if (!r.has_value()) { err = "missing result " + std::to_string(i); return nullptr; }
    return r.value()->untyped_data();
  }
  int64_t attr(const char* name, int64_t dflt) {
    auto v = attrs->get<int64_t>(name);
    return v.has_value() ? v.value() : dflt;
  }
  double attr_f(const char* name, double dflt) {
    auto v = attrs->get<double>(name);
    return v.has_value() ? v.value() : dflt;
  }
  void* workspace(size_t bytes) {
    if (bytes == 0) bytes = 256;
    auto p = scratch->Allocate(bytes);
    if (!p.has_value()) { err = "scratch allocation of " + std::to_string(bytes) + " bytes failed"; return nullptr; }
    return p.value();
  }
  std::vector<int64_t> dims(size_t i) {
    ffi::AnyBuffer b = arg(i);
    auto d = b.dimensions();
    return std::vector<int64_t>(d.begin(), d.end());
  }
};

// Strides (in elements) of a dense operand [C?, (T)?, (B)?, core...]: axes of extent 1 broadcast.
auxssm_stride StrideOf(const std::vector<int64_t>& dm, int core_rank, bool has_time, bool has_batch) {
  auxssm_stride s = {0, 0, 0};
  int64_t core = 1;
  const int n = (int)dm.size();
  for (int k = n - core_rank; k < n; ++k) core *= dm[k];
  int ax = n - core_rank - 1;
  int64_t run = core;
  if (has_batch && ax >= 0) { s.batch = dm[ax] > 1 ? run : 0; run *= dm[ax]; --ax; }
  if (has_time && ax >= 0) { s.time = dm[ax] > 1 ? run : 0; run *= dm[ax]; --ax; }
  if (ax >= 0) { s.chain = dm[ax] > 1 ? run : 0; }
  return s;
}

// LGSSM descriptor from attributes (T, dx, dy, n_chains, n_batch, parallel, time_chunks, flags) and operand shapes.
// Operand order of every LGSSM handler: ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs (empty buffers = absent).
auxssm_lgssm_desc LgssmDesc(Call& c, size_t first) {
  auxssm_lgssm_desc d = {};
  d.T = (int32_t)c.attr("T", 1); d.dx = (int32_t)c.attr("dx", 1); d.dy = (int32_t)c.attr("dy", d.dx);
  d.n_chains = (int32_t)c.attr("n_chains", 1); d.n_batch = (int32_t)c.attr("n_batch", 1);
  d.parallel = (int32_t)c.attr("parallel", 0); d.time_chunks = (int32_t)c.attr("time_chunks", 0);
  d.flags = (uint32_t)c.attr("flags", 0);
  const bool b = d.n_batch > 1, rdiag = (d.flags & AUXSSM_FLAG_R_DIAGONAL) != 0;
  d.ys = StrideOf(c.dims(first + 0), 1, true, b);
  d.m0 = StrideOf(c.dims(first + 1), 1, false, b);
  d.P0 = StrideOf(c.dims(first + 2), 2, false, b);
  d.Fs = StrideOf(c.dims(first + 3), 2, true, b);
  d.Qs = StrideOf(c.dims(first + 4), 2, true, b);
  d.bs = StrideOf(c.dims(first + 5), 1, true, b);
  d.Hs = StrideOf(c.dims(first + 6), 2, true, b);
  d.Rs = StrideOf(c.dims(first + 7), rdiag ? 1 : 2, true, b);
  d.cs = StrideOf(c.dims(first + 8), 1, true, b);
  const int64_t dx = d.dx, Tn = d.T, B = d.n_batch;
  d.ms = {Tn * B * dx, dx, B * dx};          d.xs = d.ms;  d.eps = d.ms;
  d.Ps = {Tn * B * dx * dx, dx * dx, B * dx * dx};
  return d;
}

auxssm_csmc_desc CsmcDesc(Call& c) {
  auxssm_csmc_desc d = {};
  d.T = (int32_t)c.attr("T", 1); d.N = (int32_t)c.attr("N", 1); d.d = (int32_t)c.attr("d", 1);
  d.n_chains = (int32_t)c.attr("n_chains", 1);
  d.proposal = (int32_t)c.attr("proposal", 0); d.potential = (int32_t)c.attr("potential", 0);
  d.resampling = (int32_t)c.attr("resampling", 0); d.backward = (int32_t)c.attr("backward", 0);
  d.gradient = (int32_t)c.attr("gradient", 0);
  d.ys_chain_stride = c.attr("ys_chain_stride", 0); d.gains_time_stride = c.attr("gains_time_stride", 0);
  d.sig = c.attr_f("sig", 1.0);
  d.rng_seed = (uint64_t)c.attr("rng_seed", 0); d.rng_counter = (uint64_t)c.attr("rng_counter", 0);
  d.rng_counter_ptr = nullptr;
  return d;
}

#define F(p) static_cast<const float*>(p)
#define D(p) static_cast<const double*>(p)
#define FO(p) static_cast<float*>(p)
#define DO(p) static_cast<double*>(p)
#define I32(p) static_cast<int32_t*>(p)
// two-way dispatch on the dtype of the call
#define BOTH(call32, call64) (c.f64 ? (call64) : (call32))

// ---- one function per entry point -----------------------------------------------------------------------------------

// operands: ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs;  results: ms, Ps, ell  (attr ell_only = 1: results: ell)
ffi::Error KalmanFilter(Call& c) {
  auxssm_lgssm_desc d = LgssmDesc(c, 0);
  const size_t es = c.f64 ? 8 : 4;
  const size_t nb = auxssm_kalman_filter_scratch_bytes(&d, (int32_t)es);
  void* ws = c.workspace(nb);
  const bool ell_only = c.attr("ell_only", 0) != 0;
  void* ms = ell_only ? nullptr : c.out(0);
  void* Ps = ell_only ? nullptr : c.out(1);
  void* ell = c.out(ell_only ? 0 : 2);
  if (!c.err.empty()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, c.err);
  return Status(BOTH(
      auxssm_kalman_filter_f32(c.stream, &d, F(c.in(0)), F(c.in(1)), F(c.in(2)), F(c.opt(3)), F(c.opt(4)), F(c.opt(5)),
                               F(c.opt(6)), F(c.in(7)), F(c.opt(8)), FO(ms), FO(Ps), FO(ell), ws, nb),
      auxssm_kalman_filter_f64(c.stream, &d, D(c.in(0)), D(c.in(1)), D(c.in(2)), D(c.opt(3)), D(c.opt(4)), D(c.opt(5)),
                               D(c.opt(6)), D(c.in(7)), D(c.opt(8)), DO(ms), DO(Ps), DO(ell), ws, nb)));
}

// operands: ms, Ps, Fs, Qs, bs, eps;  results: xs
ffi::Error KalmanSample(Call& c) {
  auxssm_lgssm_desc d = {};
  d.T = (int32_t)c.attr("T", 1); d.dx = d.dy = (int32_t)c.attr("dx", 1);
  d.n_chains = (int32_t)c.attr("n_chains", 1); d.n_batch = (int32_t)c.attr("n_batch", 1);
  d.parallel = (int32_t)c.attr("parallel", 0); d.time_chunks = (int32_t)c.attr("time_chunks", 0);
  d.flags = (uint32_t)c.attr("flags", 0) | AUXSSM_FLAG_H_IDENTITY;
  const bool b = d.n_batch > 1;
  d.ms = StrideOf(c.dims(0), 1, true, b); d.Ps = StrideOf(c.dims(1), 2, true, b);
  d.Fs = StrideOf(c.dims(2), 2, true, b); d.Qs = StrideOf(c.dims(3), 2, true, b); d.bs = StrideOf(c.dims(4), 1, true, b);
  d.eps = StrideOf(c.dims(5), 1, true, b);
  const int64_t dx = d.dx, Tn = d.T, B = d.n_batch;
  d.xs = {Tn * B * dx, dx, B * dx};
  const size_t nb = auxssm_kalman_sample_scratch_bytes(&d, c.f64 ? 8 : 4);
  void* ws = c.workspace(nb);
  void* xs = c.out(0);
  if (!c.err.empty()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, c.err);
  return Status(BOTH(auxssm_kalman_sample_f32(c.stream, &d, F(c.in(0)), F(c.in(1)), F(c.opt(2)), F(c.opt(3)), F(c.opt(4)),
                                              F(c.in(5)), FO(xs), ws, nb),
                     auxssm_kalman_sample_f64(c.stream, &d, D(c.in(0)), D(c.in(1)), D(c.opt(2)), D(c.opt(3)), D(c.opt(4)),
                                              D(c.in(5)), DO(xs), ws, nb)));
}

// operands: ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs, xs, ell;  results: out[n_chains];  attr which = AUXSSM_LOGPDF_*
ffi::Error LgssmLogpdf(Call& c) {
  auxssm_lgssm_desc d = LgssmDesc(c, 0);
  d.xs = StrideOf(c.dims(9), 1, true, d.n_batch > 1);
  const uint32_t which = (uint32_t)c.attr("which", AUXSSM_LOGPDF_POSTERIOR);
  const size_t nb = auxssm_lgssm_logpdf_scratch_bytes(&d, c.f64 ? 8 : 4);
  void* ws = c.workspace(nb);
  void* out = c.out(0);
  if (!c.err.empty()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, c.err);
  return Status(BOTH(
      auxssm_lgssm_logpdf_f32(c.stream, &d, which, F(c.opt(0)), F(c.in(9)), F(c.opt(10)), F(c.in(1)), F(c.in(2)), F(c.opt(3)),
                              F(c.opt(4)), F(c.opt(5)), F(c.opt(6)), F(c.opt(7)), F(c.opt(8)), FO(out), ws, nb),
      auxssm_lgssm_logpdf_f64(c.stream, &d, which, D(c.opt(0)), D(c.in(9)), D(c.opt(10)), D(c.in(1)), D(c.in(2)), D(c.opt(3)),
                              D(c.opt(4)), D(c.opt(5)), D(c.opt(6)), D(c.opt(7)), D(c.opt(8)), DO(out), ws, nb)));
}

// operands: x, normals, delta;  results: u
ffi::Error AuxiliaryDraw(Call& c) {
  const int64_t C = c.attr("n_chains", 1), n = c.attr("n_per_chain", 1);
  const int32_t ds = (int32_t)c.attr("delta_stride", 0);
  void* u = c.out(0);
  return Status(BOTH(auxssm_auxiliary_draw_f32(c.stream, C, n, F(c.in(0)), F(c.in(1)), F(c.in(2)), ds, FO(u)),
                     auxssm_auxiliary_draw_f64(c.stream, C, n, D(c.in(0)), D(c.in(1)), D(c.in(2)), ds, DO(u))));
}

// operands: delta, lp_prop, lp_rev, lt_prop, lt_rev, u, x_prop, uniforms, x (aliased to result 0);  results: x, accepted, alpha
ffi::Error KalmanMh(Call& c) {
  const int64_t C = c.attr("n_chains", 1), n = c.attr("n_per_chain", 1);
  const int32_t ds = (int32_t)c.attr("delta_stride", 0);
  void* x = c.out(0);  // input_output_aliases={8: 0} on the JAX side: x is updated in place
  return Status(BOTH(
      auxssm_kalman_mh_f32(c.stream, C, n, F(c.in(0)), ds, F(c.in(1)), F(c.in(2)), F(c.in(3)), F(c.in(4)), F(c.in(5)),
                           F(c.in(6)), F(c.in(7)), FO(x), I32(c.out(1)), FO(c.out(2))),
      auxssm_kalman_mh_f64(c.stream, C, n, D(c.in(0)), ds, D(c.in(1)), D(c.in(2)), D(c.in(3)), D(c.in(4)), D(c.in(5)),
                           D(c.in(6)), D(c.in(7)), DO(x), I32(c.out(1)), DO(c.out(2)))));
}

// operands: x, m0, P0, Fs, Qs, bs, ys, normals_aux, normals_eps, uniforms, delta;  results: x_out, accepted, alpha
ffi::Error KalmanSweep(Call& c) {
  auxssm_sweep_desc d = {};
  d.lg.T = (int32_t)c.attr("T", 2); d.lg.dx = d.lg.dy = (int32_t)c.attr("dx", 1);
  d.lg.n_chains = (int32_t)c.attr("n_chains", 1); d.lg.n_batch = 1;
  d.lg.m0 = StrideOf(c.dims(1), 1, false, false); d.lg.P0 = StrideOf(c.dims(2), 2, false, false);
  d.lg.Fs = StrideOf(c.dims(3), 2, true, false); d.lg.Qs = StrideOf(c.dims(4), 2, true, false);
  d.lg.bs = StrideOf(c.dims(5), 1, true, false);
  d.order = (int32_t)c.attr("order", 1); d.potential = (int32_t)c.attr("potential", 0);
  d.delta_stride = (int32_t)c.attr("delta_stride", 0); d.sig = c.attr_f("sig", 1.0);
  d.ys_chain_stride = c.attr("ys_chain_stride", 0);
  const size_t nb = auxssm_kalman_sweep_scratch_bytes(&d, c.f64 ? 8 : 4);
  void* ws = c.workspace(nb);
  if (!c.err.empty()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, c.err);
  return Status(BOTH(
      auxssm_kalman_sweep_f32(c.stream, &d, F(c.in(0)), F(c.in(1)), F(c.in(2)), F(c.in(3)), F(c.in(4)), F(c.in(5)), F(c.opt(6)),
                              F(c.in(7)), F(c.in(8)), F(c.in(9)), F(c.in(10)), FO(c.out(0)), I32(c.out(1)), FO(c.out(2)), ws, nb),
      auxssm_kalman_sweep_f64(c.stream, &d, D(c.in(0)), D(c.in(1)), D(c.in(2)), D(c.in(3)), D(c.in(4)), D(c.in(5)), D(c.opt(6)),
                              D(c.in(7)), D(c.in(8)), D(c.in(9)), D(c.in(10)), DO(c.out(0)), I32(c.out(1)), DO(c.out(2)), ws, nb)));
}

// operands: ys, x, u, delta;  results: aux_ys, Rs
ffi::Error SvObsFactory(Call& c) {
  const int32_t order = (int32_t)c.attr("order", 1), d = (int32_t)c.attr("d", 1), dense = (int32_t)c.attr("dense_R", 0);
  const int64_t C = c.attr("n_chains", 1), Tn = c.attr("T", 1), ycs = c.attr("ys_chain_stride", 0);
  const int32_t ds = (int32_t)c.attr("delta_stride", 0);
  return Status(BOTH(auxssm_sv_obs_factory_f32(c.stream, order, C, Tn, d, F(c.in(0)), ycs, F(c.in(1)), F(c.in(2)), F(c.in(3)), ds,
                                               FO(c.out(0)), FO(c.out(1)), dense),
                     auxssm_sv_obs_factory_f64(c.stream, order, C, Tn, d, D(c.in(0)), ycs, D(c.in(1)), D(c.in(2)), D(c.in(3)), ds,
                                               DO(c.out(0)), DO(c.out(1)), dense)));
}

// operands: ys, x (, out_in aliased to result 0 when accumulate = 1);  results: out
ffi::Error SvLogPotential(Call& c) {
  const int64_t C = c.attr("n_chains", 1), n = c.attr("n_per_chain", 1), ycs = c.attr("ys_chain_stride", 0);
  const int32_t acc = (int32_t)c.attr("accumulate", 0);
  return Status(BOTH(auxssm_sv_log_potential_f32(c.stream, C, n, F(c.in(0)), ycs, F(c.in(1)), FO(c.out(0)), acc),
                     auxssm_sv_log_potential_f64(c.stream, C, n, D(c.in(0)), ycs, D(c.in(1)), DO(c.out(0)), acc)));
}

// operands: x_star, m0, chol_P0, F, b, chol_Q, ys, u, sqrt_half_delta, grad, K, chol_Lam, normals, u_resample, u_kill,
// u_backward (empty buffers = absent; all four draw arrays empty = in-kernel Philox);  results: x_out, ancestors
ffi::Error CsmcSweep(Call& c) {
  auxssm_csmc_desc d = CsmcDesc(c);
  const size_t nb = auxssm_csmc_scratch_bytes(&d, c.f64 ? 8 : 4);
  void* ws = c.workspace(nb);
  if (!c.err.empty()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, c.err);
  return Status(BOTH(
      auxssm_csmc_sweep_f32(c.stream, &d, F(c.in(0)), F(c.in(1)), F(c.in(2)), F(c.in(3)), F(c.in(4)), F(c.in(5)), F(c.opt(6)),
                            F(c.opt(7)), F(c.opt(8)), F(c.opt(9)), F(c.opt(10)), F(c.opt(11)), F(c.opt(12)), F(c.opt(13)),
                            F(c.opt(14)), F(c.opt(15)), FO(c.out(0)), I32(c.out(1)), nullptr, nullptr, nullptr, ws, nb),
      auxssm_csmc_sweep_f64(c.stream, &d, D(c.in(0)), D(c.in(1)), D(c.in(2)), D(c.in(3)), D(c.in(4)), D(c.in(5)), D(c.opt(6)),
                            D(c.opt(7)), D(c.opt(8)), D(c.opt(9)), D(c.opt(10)), D(c.opt(11)), D(c.opt(12)), D(c.opt(13)),
                            D(c.opt(14)), D(c.opt(15)), DO(c.out(0)), I32(c.out(1)), nullptr, nullptr, nullptr, ws, nb)));
}

// operands: x_star, m0, chol_P0, F, b, chol_Q, ys, u, sqrt_half_delta, grad, normals, uniforms (empty = absent; both draw
// arrays empty = in-kernel Philox);  results: x_out, ancestors
ffi::Error PitCsmcSweep(Call& c) {
  auxssm_csmc_desc d = CsmcDesc(c);
  const size_t nb = auxssm_pit_csmc_scratch_bytes(&d, c.f64 ? 8 : 4);
  void* ws = c.workspace(nb);
  if (!c.err.empty()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, c.err);
  return Status(BOTH(
      auxssm_pit_csmc_sweep_f32(c.stream, &d, F(c.in(0)), F(c.in(1)), F(c.in(2)), F(c.in(3)), F(c.in(4)), F(c.in(5)), F(c.opt(6)),
                                F(c.in(7)), F(c.in(8)), F(c.opt(9)), F(c.opt(10)), F(c.opt(11)), FO(c.out(0)), I32(c.out(1)), ws, nb),
      auxssm_pit_csmc_sweep_f64(c.stream, &d, D(c.in(0)), D(c.in(1)), D(c.in(2)), D(c.in(3)), D(c.in(4)), D(c.in(5)), D(c.opt(6)),
                                D(c.in(7)), D(c.in(8)), D(c.opt(9)), D(c.opt(10)), D(c.opt(11)), DO(c.out(0)), I32(c.out(1)), ws, nb)));
}

// operands: u, m0, chol_P0, F, b, chol_Q, ys;  results: grad
ffi::Error CsmcGradLogTarget(Call& c) {
  auxssm_csmc_desc d = CsmcDesc(c);
  return Status(BOTH(auxssm_csmc_grad_log_target_f32(c.stream, &d, F(c.in(0)), F(c.in(1)), F(c.in(2)), F(c.in(3)), F(c.in(4)),
                                                     F(c.in(5)), F(c.opt(6)), FO(c.out(0))),
                     auxssm_csmc_grad_log_target_f64(c.stream, &d, D(c.in(0)), D(c.in(1)), D(c.in(2)), D(c.in(3)), D(c.in(4)),
                                                     D(c.in(5)), D(c.opt(6)), DO(c.out(0)))));
}

// operands: P0, Q, sqrt_half_delta;  results: K, chol_Lam
ffi::Error GuidedGains(Call& c) {
  const int32_t Tn = (int32_t)c.attr("T", 1), d = (int32_t)c.attr("d", 1);
  return Status(BOTH(auxssm_guided_gains_f32(c.stream, Tn, d, F(c.in(0)), F(c.in(1)), F(c.in(2)), FO(c.out(0)), FO(c.out(1))),
                     auxssm_guided_gains_f64(c.stream, Tn, d, D(c.in(0)), D(c.in(1)), D(c.in(2)), DO(c.out(0)), DO(c.out(1)))));
}

// operands: x_star [n,dx], P_star [n,dx,dx], xi [dx,n_pts];  results: points [n,n_pts,dx], chol [n,dx,dx]
ffi::Error SigmaPoints(Call& c) {
  const int64_t n = c.attr("n", 1);
  const int32_t dx = (int32_t)c.attr("dx", 1), np = (int32_t)c.attr("n_pts", 1);
  return Status(BOTH(auxssm_sigma_points_f32(c.stream, n, dx, np, F(c.in(0)), F(c.in(1)), F(c.in(2)), FO(c.out(0)), FO(c.out(1))),
                     auxssm_sigma_points_f64(c.stream, n, dx, np, D(c.in(0)), D(c.in(1)), D(c.in(2)), DO(c.out(0)), DO(c.out(1)))));
}

// operands: w, points, x_star, chol, f_pts, v_pts;  results: F, Q, b
ffi::Error SigmaFit(Call& c) {
  const int64_t n = c.attr("n", 1), vis = c.attr("v_item_stride", 0), vps = c.attr("v_point_stride", 0);
  const int32_t dx = (int32_t)c.attr("dx", 1), dy = (int32_t)c.attr("dy", 1), np = (int32_t)c.attr("n_pts", 1);
  return Status(BOTH(auxssm_sigma_fit_f32(c.stream, n, dx, dy, np, F(c.in(0)), F(c.in(1)), F(c.in(2)), F(c.in(3)), F(c.in(4)), F(c.in(5)),
                                          vis, vps, FO(c.out(0)), FO(c.out(1)), FO(c.out(2))),
                     auxssm_sigma_fit_f64(c.stream, n, dx, dy, np, D(c.in(0)), D(c.in(1)), D(c.in(2)), D(c.in(3)), D(c.in(4)), D(c.in(5)),
                                          vis, vps, DO(c.out(0)), DO(c.out(1)), DO(c.out(2)))));
}

// rows = product of the leading axes of operand 0, N = its last axis
void RowsOf(Call& c, size_t i, int64_t& rows, int32_t& n) {
  std::vector<int64_t> dm = c.dims(i);
  n = dm.empty() ? 1 : (int32_t)dm.back();
  rows = 1;
  for (size_t k = 0; k + 1 < dm.size(); ++k) rows *= dm[k];
}
ffi::Error Normalize(Call& c) {  // operands: log_w;  results: w
  int64_t rows; int32_t n; RowsOf(c, 0, rows, n);
  return Status(BOTH(auxssm_normalize_f32(c.stream, rows, n, F(c.in(0)), FO(c.out(0))),
                     auxssm_normalize_f64(c.stream, rows, n, D(c.in(0)), DO(c.out(0)))));
}
ffi::Error ResampleMultinomial(Call& c) {  // operands: w, u;  results: idx
  int64_t rows; int32_t n; RowsOf(c, 0, rows, n);
  return Status(BOTH(auxssm_resample_multinomial_f32(c.stream, rows, n, F(c.in(0)), F(c.in(1)), I32(c.out(0))),
                     auxssm_resample_multinomial_f64(c.stream, rows, n, D(c.in(0)), D(c.in(1)), I32(c.out(0)))));
}
ffi::Error ResampleKilling(Call& c) {  // operands: w, u_kill, u;  results: idx
  int64_t rows; int32_t n; RowsOf(c, 0, rows, n);
  return Status(BOTH(auxssm_resample_killing_f32(c.stream, rows, n, F(c.in(0)), F(c.in(1)), F(c.in(2)), I32(c.out(0))),
                     auxssm_resample_killing_f64(c.stream, rows, n, D(c.in(0)), D(c.in(1)), D(c.in(2)), I32(c.out(0)))));
}
ffi::Error ResampleSystematic(Call& c) {  // operands: w [.., M], u3 [.., 3];  results: idx [.., N];  attr N
  int64_t rows; int32_t m; RowsOf(c, 0, rows, m);
  const int32_t n = (int32_t)c.attr("N", m);
  return Status(BOTH(auxssm_resample_systematic_f32(c.stream, rows, m, n, F(c.in(0)), F(c.in(1)), I32(c.out(0))),
                     auxssm_resample_systematic_f64(c.stream, rows, m, n, D(c.in(0)), D(c.in(1)), I32(c.out(0)))));
}
ffi::Error Categorical(Call& c) {  // operands: w, u;  results: idx
  int64_t rows; int32_t n; RowsOf(c, 0, rows, n);
  return Status(BOTH(auxssm_categorical_f32(c.stream, rows, n, F(c.in(0)), F(c.in(1)), I32(c.out(0))),
                     auxssm_categorical_f64(c.stream, rows, n, D(c.in(0)), D(c.in(1)), I32(c.out(0)))));
}
ffi::Error Cholesky(Call& c) {  // operands: A [.., d, d];  results: L
  std::vector<int64_t> dm = c.dims(0);
  const int32_t d = dm.empty() ? 1 : (int32_t)dm.back();
  int64_t n = 1;
  for (size_t k = 0; k + 2 < dm.size(); ++k) n *= dm[k];
  return Status(BOTH(auxssm_cholesky_f32(c.stream, n, d, F(c.in(0)), FO(c.out(0))),
                     auxssm_cholesky_f64(c.stream, n, d, D(c.in(0)), DO(c.out(0)))));
}
ffi::Error DeltaAdaptation(Call& c) {  // operands: delta, acc_rate;  results: delta_out;  attrs: target, rate, min, max
  const int64_t n = c.arg(0).element_count();
  const int32_t as = c.arg(1).element_count() > 1 ? 1 : 0;
  const double tg = c.attr_f("target_rate", 0.5), rt = c.attr_f("adaptation_rate", 0.05);
  const double lo = c.attr_f("min_delta", 1e-20), hi = c.attr_f("max_delta", 1e20);
  return Status(BOTH(auxssm_delta_adaptation_f32(c.stream, n, F(c.in(0)), F(c.in(1)), as, tg, rt, lo, hi, FO(c.out(0))),
                     auxssm_delta_adaptation_f64(c.stream, n, D(c.in(0)), D(c.in(1)), as, tg, rt, lo, hi, DO(c.out(0)))));
}
ffi::Error McmcStats(Call& c) {  // operands: x_old, x_new, updated (S32), stats (aliased to result 0);  results: stats
  const int32_t C = (int32_t)c.attr("n_chains", 1), per = (int32_t)c.attr("per_step", 1);
  const int64_t n = c.attr("n_per_chain", 1);
  const double beta = c.attr_f("beta", 0.9);
  return Status(BOTH(auxssm_mcmc_stats_f32(c.stream, C, n, per, F(c.in(0)), F(c.in(1)), I32(c.in(2)), beta, FO(c.out(0))),
                     auxssm_mcmc_stats_f64(c.stream, C, n, per, D(c.in(0)), D(c.in(1)), I32(c.in(2)), beta, DO(c.out(0)))));
}
ffi::Error Autocov(Call& c) {  // operands: x [C, n];  results: mean [C] (F64), acov [C, n] (F64)
  std::vector<int64_t> dm = c.dims(0);
  const int32_t C = (int32_t)dm[0];
  const int64_t n = dm[1];
  return Status(BOTH(auxssm_autocov_f32(c.stream, C, n, F(c.in(0)), DO(c.out(0)), DO(c.out(1))),
                     auxssm_autocov_f64(c.stream, C, n, D(c.in(0)), DO(c.out(0)), DO(c.out(1)))));
}
// operands: like (any buffer whose dtype selects f32 / f64);  results: out;  attrs: seed, counter.
// XLA programs are functional, so the draw-call counter is an attribute rather than device state: it is folded into the
// key (seed + golden-ratio * counter), giving an independent Philox stream per call.
ffi::Error PhiloxFill(Call& c, bool normal) {
  auto r = c.rets->get<ffi::AnyBuffer>(0);
  if (!r.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "missing result");
  const int64_t n = (int64_t)r.value()->element_count();
  const uint64_t key = (uint64_t)c.attr("seed", 0) + 0x9E3779B97F4A7C15ull * (uint64_t)c.attr("counter", 0);
  void* out = r.value()->untyped_data();
  if (normal)
    return Status(BOTH(auxssm_philox_normal_f32(c.stream, key, nullptr, n, FO(out)),
                       auxssm_philox_normal_f64(c.stream, key, nullptr, n, DO(out))));
  return Status(BOTH(auxssm_philox_uniform_f32(c.stream, key, nullptr, n, FO(out)),
                     auxssm_philox_uniform_f64(c.stream, key, nullptr, n, DO(out))));
}
ffi::Error PhiloxUniform(Call& c) { return PhiloxFill(c, false); }
ffi::Error PhiloxNormal(Call& c) { return PhiloxFill(c, true); }

template <ffi::Error (*Fn)(Call&)>
ffi::Error Handler(Stream stream, ffi::ScratchAllocator scratch, ffi::RemainingArgs args, ffi::RemainingRets rets,
                   ffi::Dictionary attrs) {
  Call c{stream, &scratch, &args, &rets, &attrs};
  if (args.size() > 0) {
    auto b = args.get<ffi::AnyBuffer>(0);
    if (b.has_value()) c.f64 = b.value().element_type() == ffi::DataType::F64;
  }
  ffi::Error e = Fn(c);
  if (!c.err.empty()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, c.err);
  return e;
}

}  // namespace

// One exported handler symbol per entry point family: register with
//   jax.ffi.register_ffi_target("auxssm_<name>", jax.ffi.pycapsule(lib.auxssm_ffi_<name>), platform="CUDA")
#define AUXSSM_FFI(name, fn)                                                                                          \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(auxssm_ffi_##name, Handler<fn>,                                                       \
                                ffi::Ffi::Bind()                                                                      \
                                    .Ctx<ffi::PlatformStream<Stream>>()                                               \
                                    .Ctx<ffi::ScratchAllocator>()                                                     \
                                    .RemainingArgs()                                                                  \
                                    .RemainingRets()                                                                  \
                                    .Attrs())

AUXSSM_FFI(kalman_filter, KalmanFilter);
AUXSSM_FFI(kalman_sample, KalmanSample);
AUXSSM_FFI(lgssm_logpdf, LgssmLogpdf);
AUXSSM_FFI(auxiliary_draw, AuxiliaryDraw);
AUXSSM_FFI(kalman_mh, KalmanMh);
AUXSSM_FFI(kalman_sweep, KalmanSweep);
AUXSSM_FFI(sv_obs_factory, SvObsFactory);
AUXSSM_FFI(sv_log_potential, SvLogPotential);
AUXSSM_FFI(csmc_sweep, CsmcSweep);
AUXSSM_FFI(pit_csmc_sweep, PitCsmcSweep);
AUXSSM_FFI(csmc_grad_log_target, CsmcGradLogTarget);
AUXSSM_FFI(guided_gains, GuidedGains);
AUXSSM_FFI(normalize, Normalize);
AUXSSM_FFI(resample_multinomial, ResampleMultinomial);
AUXSSM_FFI(resample_killing, ResampleKilling);
AUXSSM_FFI(resample_systematic, ResampleSystematic);
AUXSSM_FFI(categorical, Categorical);
AUXSSM_FFI(cholesky, Cholesky);
AUXSSM_FFI(delta_adaptation, DeltaAdaptation);
AUXSSM_FFI(mcmc_stats, McmcStats);
AUXSSM_FFI(autocov, Autocov);
AUXSSM_FFI(sigma_points, SigmaPoints);
AUXSSM_FFI(sigma_fit, SigmaFit);
AUXSSM_FFI(philox_uniform, PhiloxUniform);
AUXSSM_FFI(philox_normal, PhiloxNormal);
