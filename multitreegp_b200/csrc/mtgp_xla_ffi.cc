// mtgp_xla_ffi.cc -- XLA-FFI (jax.ffi) adapter over the C ABI of include/mtgp.h.
//
// jax / jaxlib (and with them xla/ffi/api/ffi.h) are not installable in this image.  multitreegp_b200/build.py
// builds this file into libmtgp_b200_ffi.so when `jax.ffi.include_dir()` resolves; __graft_entry__.build()
// additionally builds it against tests/fake_xla_ffi (a stand-in for the header) into
// tests/hostlib/libmtgp_ffi_fake.so, so that the adapter COMPILES and its handler bodies are EXECUTED on the
// GPU by tests/test_ffi_adapter.py (through the mtgp_*_ffi_test entry points at the end of this file) with
// buffers shaped as XLA delivers them.  It is a thin adapter: every line of work happens behind
// mtgp_population_fitness / mtgp_population_fitness_pointwise.
//
// Boundary: the drop-in evaluator's __call__ issues
//     jax.ffi.ffi_call("mtgp_rollout", ShapeDtypeStruct((), f32), vmap_method="expand_dims")(
//         coefficients, nodes, x0, ts, targets, obs_keys, params, ys_true, **attrs)
// Under the strategy's vmap over the population (genetic_programming.py:252) and shard_map over
// devices (:259-262) the handler runs ONCE PER DEVICE SHARD with
//     coefficients f32[P/n, T, N, 1], nodes f32[P/n, T, N, 3]      (batched operands)
//     data operands with a leading dimension of 1                   (un-batched, expand_dims)
// and returns fitness f32[P/n]; XLA gathers the shards.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "../../include/mtgp.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error RolloutImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                              ffi::Buffer<ffi::F32> coefficients, ffi::Buffer<ffi::F32> nodes,
                              ffi::Buffer<ffi::F32> x0, ffi::Buffer<ffi::F32> ts, ffi::Buffer<ffi::F32> targets,
                              ffi::Buffer<ffi::U32> obs_keys, ffi::Buffer<ffi::F32> params,
                              ffi::Buffer<ffi::F32> ys_true, ffi::ResultBuffer<ffi::F32> fitness,
                              ffi::Span<const int32_t> opcodes, ffi::Span<const int32_t> args, int32_t kind,
                              int32_t env_id, int32_t n_obs, int32_t state_size, int32_t max_steps, float dt0,
                              float max_fitness, float obs_noise, int32_t threefry_partitionable, int32_t solver,
                              int32_t controller, float rtol, float atol, float dtmin, float pcoeff, float icoeff,
                              float dcoeff, float pid_safety, float pid_factormin, float pid_factormax) {
  auto nd = nodes.dimensions();
  if (nd.size() < 4 || nd[nd.size() - 1] != 3) return ffi::Error::InvalidArgument("nodes must be [..., T, N, 3]");
  const int N = (int)nd[nd.size() - 2], T = (int)nd[nd.size() - 3];
  long long P = 1;
  for (size_t i = 0; i + 3 < nd.size(); i++) P *= nd[i];
  auto xd = x0.dimensions();  // [1?, B, D]
  const int D = (int)xd[xd.size() - 1], B = (int)xd[xd.size() - 2];
  const int K = (int)ts.dimensions().back();

  mtgp_node_table table;
  std::memset(&table, 0, sizeof(table));
  if (opcodes.size() != args.size() || opcodes.size() > MTGP_MAX_NODE_IDS)
    return ffi::Error::InvalidArgument("opcodes/args attributes must have the same length <= 64");
  table.n_ids = (int32_t)opcodes.size();
  for (size_t i = 0; i < opcodes.size(); i++) {
    table.opcode[i] = (uint8_t)opcodes[i];
    table.arg[i] = args[i];
  }
  mtgp_rollout_desc d;
  std::memset(&d, 0, sizeof(d));
  d.kind = kind; d.env_id = env_id; d.B = B; d.K = K; d.D = D; d.n_obs = n_obs;
  d.n_targets = targets.element_count() ? (int)targets.dimensions().back() : 0;   // f32[B, 0] is legal (Acrobot)
  d.n_params = params.element_count() ? (int)params.dimensions().back() : 0;
  d.state_size = state_size; d.max_steps = max_steps; d.dt0 = dt0; d.max_fitness = max_fitness;
  d.obs_noise = obs_noise; d.threefry_partitionable = threefry_partitionable;
  d.solver = solver; d.controller = controller; d.rtol = rtol; d.atol = atol; d.dtmin = dtmin;
  d.pcoeff = pcoeff; d.icoeff = icoeff; d.dcoeff = dcoeff;  // all 0: diffrax defaults (0, 1, 0)
  d.pid_safety = pid_safety; d.pid_factormin = pid_factormin; d.pid_factormax = pid_factormax;  // 0: 0.9, 0.2, 10
  d.x0 = x0.typed_data(); d.ts = ts.typed_data();
  d.targets = d.n_targets ? targets.typed_data() : nullptr;
  d.obs_keys = obs_keys.element_count() ? obs_keys.typed_data() : nullptr;
  d.params = d.n_params ? params.typed_data() : nullptr;
  d.ys_true = ys_true.element_count() ? ys_true.typed_data() : nullptr;

  const size_t ws_bytes = mtgp_workspace_bytes((int)P, T, N, &d);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error::Internal("mtgp: scratch allocation failed");
  const int rc = mtgp_population_fitness(nodes.typed_data(), 3, coefficients.typed_data(), 1, (int)P, T, N, &table,
                                         &d, 0.0f, fitness->typed_data(), nullptr, *ws, ws_bytes, stream);
  if (rc == MTGP_ERR_INVALID_ARGUMENT || rc == MTGP_ERR_UNSUPPORTED) return ffi::Error::InvalidArgument(mtgp_last_error());
  if (rc != MTGP_OK) return ffi::Error::Internal(mtgp_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mtgp_rollout_ffi, RolloutImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::Buffer<ffi::F32>>()   // coefficients [P, T, N, 1]
        .Arg<ffi::Buffer<ffi::F32>>()   // nodes        [P, T, N, 3]
        .Arg<ffi::Buffer<ffi::F32>>()   // x0
        .Arg<ffi::Buffer<ffi::F32>>()   // ts
        .Arg<ffi::Buffer<ffi::F32>>()   // targets (may be zero-size)
        .Arg<ffi::Buffer<ffi::U32>>()   // obs_noise_keys
        .Arg<ffi::Buffer<ffi::F32>>()   // params  [B, n_params]
        .Arg<ffi::Buffer<ffi::F32>>()   // ys_true (SR; zero-size otherwise)
        .Ret<ffi::Buffer<ffi::F32>>()   // fitness [P]
        .Attr<ffi::Span<const int32_t>>("opcodes")
        .Attr<ffi::Span<const int32_t>>("args")
        .Attr<int32_t>("kind")
        .Attr<int32_t>("env_id")
        .Attr<int32_t>("n_obs")
        .Attr<int32_t>("state_size")
        .Attr<int32_t>("max_steps")
        .Attr<float>("dt0")
        .Attr<float>("max_fitness")
        .Attr<float>("obs_noise")
        .Attr<int32_t>("threefry_partitionable")
        .Attr<int32_t>("solver")
        .Attr<int32_t>("controller")
        .Attr<float>("rtol")
        .Attr<float>("atol")
        .Attr<float>("dtmin")
        .Attr<float>("pcoeff")
        .Attr<float>("icoeff")
        .Attr<float>("dcoeff")
        .Attr<float>("pid_safety")
        .Attr<float>("pid_factormin")
        .Attr<float>("pid_factormax"));

// Pointwise evaluator (pointwise_evaluate.Evaluator): data = (X [n, n_vars], Y [n, T]), un-batched operands
// arrive with a leading dimension of 1.
static ffi::Error PointwiseImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> coefficients,
                                ffi::Buffer<ffi::F32> nodes, ffi::Buffer<ffi::F32> X, ffi::Buffer<ffi::F32> Y,
                                ffi::ResultBuffer<ffi::F32> fitness, ffi::Span<const int32_t> opcodes,
                                ffi::Span<const int32_t> args, float max_fitness) {
  auto nd = nodes.dimensions();
  if (nd.size() < 4 || nd[nd.size() - 1] != 3) return ffi::Error::InvalidArgument("nodes must be [..., T, N, 3]");
  const int N = (int)nd[nd.size() - 2], T = (int)nd[nd.size() - 3];
  long long P = 1;
  for (size_t i = 0; i + 3 < nd.size(); i++) P *= nd[i];
  auto xd = X.dimensions();
  if (xd.size() < 2) return ffi::Error::InvalidArgument("X must be [..., n_points, n_vars]");
  const int n_vars = (int)xd[xd.size() - 1], n_points = (int)xd[xd.size() - 2];
  if (Y.element_count() != (size_t)n_points * T) return ffi::Error::InvalidArgument("Y must be [..., n_points, T]");
  mtgp_node_table table;
  std::memset(&table, 0, sizeof(table));
  if (opcodes.size() != args.size() || opcodes.size() > MTGP_MAX_NODE_IDS)
    return ffi::Error::InvalidArgument("opcodes/args attributes must have the same length <= 64");
  table.n_ids = (int32_t)opcodes.size();
  for (size_t i = 0; i < opcodes.size(); i++) {
    table.opcode[i] = (uint8_t)opcodes[i];
    table.arg[i] = args[i];
  }
  const size_t ws_bytes = mtgp_workspace_bytes_pointwise((int)P, T, N, n_points, n_vars);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error::Internal("mtgp: scratch allocation failed");
  const int rc = mtgp_population_fitness_pointwise(nodes.typed_data(), 3, coefficients.typed_data(), 1, (int)P, T, N,
                                                   &table, X.typed_data(), Y.typed_data(), n_points, n_vars, max_fitness,
                                                   0.0f, fitness->typed_data(), *ws, ws_bytes, stream);
  if (rc == MTGP_ERR_INVALID_ARGUMENT || rc == MTGP_ERR_UNSUPPORTED) return ffi::Error::InvalidArgument(mtgp_last_error());
  if (rc != MTGP_OK) return ffi::Error::Internal(mtgp_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mtgp_pointwise_ffi, PointwiseImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::Buffer<ffi::F32>>()   // coefficients [P, T, N, 1]
        .Arg<ffi::Buffer<ffi::F32>>()   // nodes        [P, T, N, 3]
        .Arg<ffi::Buffer<ffi::F32>>()   // X [1, n_points, n_vars]
        .Arg<ffi::Buffer<ffi::F32>>()   // Y [1, n_points, T]
        .Ret<ffi::Buffer<ffi::F32>>()   // fitness [P]
        .Attr<ffi::Span<const int32_t>>("opcodes")
        .Attr<ffi::Span<const int32_t>>("args")
        .Attr<float>("max_fitness"));

#ifdef MTGP_FAKE_XLA_FFI
// ---- test entry points (stand-in header only): build the Buffers XLA would pass and run the handler bodies ----
static std::vector<int64_t> dims_of(const int64_t* d, int n) { return std::vector<int64_t>(d, d + n); }
struct mtgp_ffi_test_buffer { void* data; const int64_t* dims; int32_t rank; };

extern "C" __attribute__((visibility("default")))
int mtgp_rollout_ffi_test(void* stream, const mtgp_ffi_test_buffer* bufs /* 9: the 8 operands + the result */,
                          const int32_t* opcodes, const int32_t* args, int32_t n_ids, const int32_t* iattrs /* 8 */,
                          const float* fattrs /* 12 */, char* err, int32_t err_cap) {
  auto F = [&](int i) { return ffi::Buffer<ffi::F32>((float*)bufs[i].data, dims_of(bufs[i].dims, bufs[i].rank)); };
  ffi::Buffer<ffi::U32> keys((uint32_t*)bufs[5].data, dims_of(bufs[5].dims, bufs[5].rank));
  ffi::Buffer<ffi::F32> out = F(8);
  ffi::Error e = RolloutImpl((cudaStream_t)stream, ffi::ScratchAllocator(), F(0), F(1), F(2), F(3), F(4), keys, F(6), F(7),
                             ffi::ResultBuffer<ffi::F32>(&out), ffi::Span<const int32_t>(opcodes, (size_t)n_ids),
                             ffi::Span<const int32_t>(args, (size_t)n_ids), iattrs[0], iattrs[1], iattrs[2], iattrs[3],
                             iattrs[4], fattrs[0], fattrs[1], fattrs[2], iattrs[5], iattrs[6], iattrs[7], fattrs[3],
                             fattrs[4], fattrs[5], fattrs[6], fattrs[7], fattrs[8], fattrs[9], fattrs[10], fattrs[11]);
  cudaStreamSynchronize((cudaStream_t)stream);  // the scratch is freed when this returns
  if (e.failure() && err && err_cap > 0) { std::strncpy(err, e.message().c_str(), err_cap - 1); err[err_cap - 1] = 0; }
  return e.code();
}

extern "C" __attribute__((visibility("default")))
int mtgp_pointwise_ffi_test(void* stream, const mtgp_ffi_test_buffer* bufs /* 5: 4 operands + the result */,
                            const int32_t* opcodes, const int32_t* args, int32_t n_ids, float max_fitness, char* err,
                            int32_t err_cap) {
  auto F = [&](int i) { return ffi::Buffer<ffi::F32>((float*)bufs[i].data, dims_of(bufs[i].dims, bufs[i].rank)); };
  ffi::Buffer<ffi::F32> out = F(4);
  ffi::Error e = PointwiseImpl((cudaStream_t)stream, ffi::ScratchAllocator(), F(0), F(1), F(2), F(3),
                               ffi::ResultBuffer<ffi::F32>(&out), ffi::Span<const int32_t>(opcodes, (size_t)n_ids),
                               ffi::Span<const int32_t>(args, (size_t)n_ids), max_fitness);
  cudaStreamSynchronize((cudaStream_t)stream);
  if (e.failure() && err && err_cap > 0) { std::strncpy(err, e.message().c_str(), err_cap - 1); err[err_cap - 1] = 0; }
  return e.code();
}
#endif  // MTGP_FAKE_XLA_FFI
#endif  // xla/ffi/api/ffi.h
