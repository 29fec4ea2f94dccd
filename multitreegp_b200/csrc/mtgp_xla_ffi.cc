// mtgp_xla_ffi.cc -- XLA-FFI (jax.ffi) adapter over the C ABI of include/mtgp.h.
//
// NOT COMPILED IN THIS IMAGE: jax / jaxlib (and with them xla/ffi/api/ffi.h) are not installed and
// there is no network.  multitreegp_b200/build.py builds this file into libmtgp_b200_ffi.so only
// when `jax.ffi.include_dir()` resolves.  It is deliberately a thin adapter: every line of work
// happens behind mtgp_population_fitness, which the GPU parity tests exercise through ctypes.
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
                              int32_t controller, float rtol, float atol, float dtmin) {
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
  d.solver = solver; d.controller = controller; d.rtol = rtol; d.atol = atol; d.dtmin = dtmin;  // PID gains: diffrax defaults
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
        .Attr<float>("dtmin"));
#endif  // xla/ffi/api/ffi.h
