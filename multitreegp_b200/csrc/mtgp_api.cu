// mtgp_api.cu -- the extern "C" surface of libmtgp_b200.so (declared in include/mtgp.h).
//
// Validates arguments, carves the caller's workspace, and enqueues encode -> lower -> evaluate
// on the caller's stream.  No entry point synchronises except the *_host ones, none allocates
// device memory except the *_host ones (a per-device scratch cache, grown on demand).
#include <stdarg.h>
#include <string.h>

#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "mtgp_internal.cuh"

int mtgp_rollout_slot_map(const mtgp_rollout_desc& d, int T, int* n_data, int* n_latent);

// ---- errors --------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void mtgp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int mtgp_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    mtgp_set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return MTGP_ERR_CUDA;
  }
  return MTGP_OK;
}

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {
  unsigned char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<unsigned char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return r;
  }
};

struct Workspace {
  mtgp_bytecode bc;
  uint2* prog;
  uint32_t* meta;
  int32_t* order;
  float* Xt;
  float* Yt;
  void* noise;
  size_t bytes;
};

// layout shared by mtgp_workspace_bytes and every entry point
Workspace carve(void* ws, int P, int T, int N, int n_points_padded, int n_vars, size_t noise_bytes = 0) {
  Carver c(ws);
  Workspace w{};
  const size_t rows = (size_t)P * T * N, trees = (size_t)P * T;
  w.bc.op = c.take<uint8_t>(rows);
  w.bc.a = c.take<int16_t>(rows);
  w.bc.b = c.take<int16_t>(rows);
  w.bc.coef = c.take<float>(rows);
  w.bc.live = c.take<int32_t>(trees);
  w.bc.flags = c.take<uint8_t>(trees);
  w.prog = c.take<uint2>(trees * (size_t)mtgp_lstride(N));
  w.meta = c.take<uint32_t>(trees);
  w.order = c.take<int32_t>((size_t)P + 2);  // + the two slot counters of the persistent rollout launches
  w.Xt = c.take<float>((size_t)n_points_padded * (n_vars > 0 ? n_vars : 0));
  w.Yt = c.take<float>((size_t)n_points_padded * T);
  w.noise = noise_bytes ? c.take<unsigned char>(noise_bytes) : nullptr;
  w.bytes = align_up(c.off, 256);
  return w;
}

int validate_table(const mtgp_node_table* t) {
  if (!t || t->n_ids < 2 || t->n_ids > MTGP_MAX_NODE_IDS) {
    mtgp_set_error("node table: n_ids must be in [2, %d]", MTGP_MAX_NODE_IDS);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (t->opcode[0] != MTGP_OP_EMPTY || t->opcode[1] != MTGP_OP_COEF) {
    mtgp_set_error("node table: id 0 must be the empty node and id 1 the coefficient leaf (genetic_programming.py:135)");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  for (int i = 2; i < t->n_ids; i++) {
    if (t->opcode[i] < MTGP_OP_VAR || t->opcode[i] >= MTGP_OP_COUNT_) {
      mtgp_set_error("node table: id %d has unknown opcode %d", i, (int)t->opcode[i]);
      return MTGP_ERR_INVALID_ARGUMENT;
    }
  }
  return MTGP_OK;
}

int validate_shape(int P, int T, int N) {
  if (P < 0 || T < 1 || N < 1) {
    mtgp_set_error("population shape must satisfy P >= 0, T >= 1, N >= 1 (got P=%d T=%d N=%d)", P, T, N);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (N > MTGP_MAX_NODES) {
    mtgp_set_error("max_nodes=%d exceeds MTGP_MAX_NODES=%d", N, MTGP_MAX_NODES);
    return MTGP_ERR_UNSUPPORTED;
  }
  return MTGP_OK;
}

int check_vars(const mtgp_node_table* t, int n_data) {
  for (int i = 2; i < t->n_ids; i++) {
    if (t->opcode[i] == MTGP_OP_VAR && (t->arg[i] < 0 || t->arg[i] >= n_data)) {
      mtgp_set_error("variable node id %d reads data[%d] but the data vector has %d entries", i, t->arg[i], n_data);
      return MTGP_ERR_INVALID_ARGUMENT;
    }
  }
  return MTGP_OK;
}

int validate_rollout(const mtgp_rollout_desc* d, int T, int N, const mtgp_node_table* table, int* tmp_base) {
  if (!d) { mtgp_set_error("rollout descriptor is NULL"); return MTGP_ERR_INVALID_ARGUMENT; }
  if (T > 32) {  // per-warp table of program offsets in the rollout kernels (kMaxTrees)
    mtgp_set_error("rollout kernels support at most 32 trees per individual (got %d)", T);
    return MTGP_ERR_UNSUPPORTED;
  }
  if (d->B < 1 || d->K < 2) {
    mtgp_set_error("need B >= 1 trajectories and K >= 2 time points (got B=%d K=%d)", d->B, d->K);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (!d->x0 || !d->ts) { mtgp_set_error("x0 / ts must not be NULL"); return MTGP_ERR_INVALID_ARGUMENT; }
  if (!(d->dt0 > 0.0f) || d->max_steps < 1) {
    mtgp_set_error("dt0 must be > 0 and max_steps >= 1");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if ((d->solver != MTGP_SOLVER_EULER && d->solver != MTGP_SOLVER_DOPRI5) ||
      (d->controller != MTGP_CONTROLLER_CONSTANT && d->controller != MTGP_CONTROLLER_PID)) {
    mtgp_set_error("unknown solver %d / step-size controller %d", d->solver, d->controller);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (d->controller == MTGP_CONTROLLER_PID) {
    if (d->solver != MTGP_SOLVER_DOPRI5) {  // diffrax raises too: Euler has no embedded error estimate
      mtgp_set_error("PIDController needs a solver with an error estimate (Dopri5), not Euler");
      return MTGP_ERR_INVALID_ARGUMENT;
    }
    if (!(d->rtol > 0.0f) && !(d->atol > 0.0f)) {
      mtgp_set_error("PIDController needs rtol > 0 or atol > 0");
      return MTGP_ERR_INVALID_ARGUMENT;
    }
  }
  int envD = 0, envP = 0;
  switch (d->env_id) {
    case MTGP_ENV_NONE: break;
    case MTGP_ENV_ACROBOT: envD = 4; envP = 4; break;
    case MTGP_ENV_HARMONIC: envD = 2; envP = 2; break;
    case MTGP_ENV_REACTOR: envD = 3; envP = 8; break;
    default: mtgp_set_error("unknown environment id %d", d->env_id); return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (d->kind == MTGP_EVAL_SR) {
    if (d->env_id != MTGP_ENV_NONE || d->D != T || !d->ys_true) {
      mtgp_set_error("SR evaluator: env must be NONE, state dimension D (%d) must equal the number of trees (%d), ys_true required", d->D, T);
      return MTGP_ERR_INVALID_ARGUMENT;
    }
  } else if (d->kind == MTGP_EVAL_STATIC || d->kind == MTGP_EVAL_DYNAMIC) {
    if (d->env_id == MTGP_ENV_NONE || d->D != envD || d->n_params != envP || !d->params) {
      mtgp_set_error("control evaluator: environment %d needs D=%d and %d parameters per trajectory (got D=%d, n_params=%d)",
                     d->env_id, envD, envP, d->D, d->n_params);
      return MTGP_ERR_INVALID_ARGUMENT;
    }
    if (d->n_obs < 1 || d->n_obs > d->D) {
      mtgp_set_error("n_obs must be in [1, D]");
      return MTGP_ERR_INVALID_ARGUMENT;
    }
    const int want_targets = d->env_id == MTGP_ENV_ACROBOT ? 0 : 1;
    if (d->n_targets != want_targets || (want_targets && !d->targets)) {
      mtgp_set_error("environment %d expects %d target(s)", d->env_id, want_targets);
      return MTGP_ERR_INVALID_ARGUMENT;
    }
    if (d->obs_noise != 0.0f && !d->obs_keys) {
      mtgp_set_error("obs_noise != 0 needs obs_keys");
      return MTGP_ERR_INVALID_ARGUMENT;
    }
    const int S = d->kind == MTGP_EVAL_DYNAMIC ? d->state_size : 0;
    if (d->kind == MTGP_EVAL_DYNAMIC && S < 1) {
      mtgp_set_error("dynamic evaluator needs state_size >= 1");
      return MTGP_ERR_INVALID_ARGUMENT;
    }
    if (T - S != 1) {  // jnp.squeeze(u) must be a scalar for these environments (acrobot.py:52, reactor.py:80)
      mtgp_set_error("the environment takes n_control = 1 control: expected %d trees, got %d", S + 1, T);
      return MTGP_ERR_INVALID_ARGUMENT;
    }
  } else {
    mtgp_set_error("unknown evaluator kind %d", d->kind);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  int n_data, n_latent;
  *tmp_base = mtgp_rollout_slot_map(*d, T, &n_data, &n_latent);
  if (*tmp_base + mtgp_tmp_bound(N) > MTGP_MAX_SLOTS) {
    mtgp_set_error("max_nodes=%d needs %d value-file slots, more than %d", N, *tmp_base + mtgp_tmp_bound(N), MTGP_MAX_SLOTS);
    return MTGP_ERR_UNSUPPORTED;
  }
  return check_vars(table, n_data);
}

}  // namespace

// ---- exported ------------------------------------------------------------------------------------
extern "C" {

int mtgp_abi_version(void) { return MTGP_ABI_VERSION; }
const char* mtgp_last_error(void) { return g_err; }

static int padded_points(int n_points) { return (int)align_up((size_t)(n_points > 0 ? n_points : 0), 1024); }

size_t mtgp_workspace_bytes(int P, int T, int N, const mtgp_rollout_desc* rollout) {
  return carve(nullptr, P, T, N, 0, 0, rollout ? mtgp_noise_workspace_bytes(*rollout) : 0).bytes;
}

size_t mtgp_workspace_bytes_pointwise(int P, int T, int N, int n_points, int n_vars) {
  return carve(nullptr, P, T, N, padded_points(n_points), n_vars).bytes;
}

size_t mtgp_bytecode_bytes(int P, int T, int N) {
  // a caller that wants its own mtgp_bytecode arrays: op(1) + a(2) + b(2) + coef(4) per row, live(4) + flags(1) per tree
  return (size_t)P * T * N * 9 + (size_t)P * T * 5;
}

int mtgp_encode(const float* nodes, int nodes_stride, const float* coef, int coef_stride, int P, int T, int N,
                const mtgp_node_table* table, mtgp_bytecode out, void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if ((rc = validate_table(table))) return rc;
  if (!nodes || !coef || nodes_stride < 3 || coef_stride < 1) {
    mtgp_set_error("mtgp_encode: nodes/coef must not be NULL, strides must be >= 3 / >= 1");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (!out.op || !out.a || !out.b || !out.coef || !out.live || !out.flags) {
    mtgp_set_error("mtgp_encode: every array of the output bytecode must be allocated");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  return mtgp_launch_encode(nodes, nodes_stride, coef, coef_stride, P, T, N, table, out, (cudaStream_t)stream);
}

int mtgp_decode(mtgp_bytecode in, int P, int T, int N, float* trees_out, void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if (!trees_out || !in.op || !in.a || !in.b || !in.coef) {
    mtgp_set_error("mtgp_decode: NULL buffer");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  return mtgp_launch_decode(in, P, T, N, trees_out, (cudaStream_t)stream);
}

// Lowered programs of a bytecode, for tests and tools: prog_out [P*T, lstride] uint2, meta_out [P*T].
int mtgp_lower(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table, int tmp_base,
               void* prog_out, uint32_t* meta_out, void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if ((rc = validate_table(table))) return rc;
  return mtgp_launch_lower(bc, P, T, N, table, tmp_base, mtgp_lstride(N), static_cast<uint2*>(prog_out), meta_out,
                           (cudaStream_t)stream);
}
int mtgp_program_stride(int N) { return mtgp_lstride(N); }

int mtgp_rollout(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table,
                 const mtgp_rollout_desc* desc, float* fitness, const mtgp_trajectory_out* traj,
                 unsigned long long* steps_done, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if ((rc = validate_table(table))) return rc;
  int tmp_base = 0;
  if ((rc = validate_rollout(desc, T, N, table, &tmp_base))) return rc;
  if (!fitness) { mtgp_set_error("fitness is NULL"); return MTGP_ERR_INVALID_ARGUMENT; }
  Workspace w = carve(workspace, P, T, N, 0, 0, mtgp_noise_workspace_bytes(*desc));
  if (!workspace || workspace_bytes < w.bytes) {
    mtgp_set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return MTGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = mtgp_launch_lower(bc, P, T, N, table, tmp_base, mtgp_lstride(N), w.prog, w.meta, s))) return rc;
  mtgp_programs progs{w.prog, w.meta, mtgp_lstride(N), tmp_base, mtgp_tmp_bound(N), w.order};
  return mtgp_launch_rollout(progs, nullptr, P, T, N, *desc, 0.0f, fitness, traj, steps_done, w.noise, s);
}

int mtgp_population_fitness_ex(const float* nodes, int nodes_stride, const float* coef, int coef_stride, int P,
                               int T, int N, const mtgp_node_table* table, const mtgp_rollout_desc* desc,
                               float size_parsimony, float* fitness, const mtgp_trajectory_out* traj,
                               unsigned long long* steps_done, void* workspace, size_t workspace_bytes,
                               void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if ((rc = validate_table(table))) return rc;
  int tmp_base = 0;
  if ((rc = validate_rollout(desc, T, N, table, &tmp_base))) return rc;
  if (!fitness || !nodes || !coef) { mtgp_set_error("NULL population / fitness buffer"); return MTGP_ERR_INVALID_ARGUMENT; }
  Workspace w = carve(workspace, P, T, N, 0, 0, mtgp_noise_workspace_bytes(*desc));
  if (!workspace || workspace_bytes < w.bytes) {
    mtgp_set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return MTGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = mtgp_launch_encode(nodes, nodes_stride, coef, coef_stride, P, T, N, table, w.bc, s))) return rc;
  if ((rc = mtgp_launch_lower(w.bc, P, T, N, table, tmp_base, mtgp_lstride(N), w.prog, w.meta, s))) return rc;
  mtgp_programs progs{w.prog, w.meta, mtgp_lstride(N), tmp_base, mtgp_tmp_bound(N), w.order};
  return mtgp_launch_rollout(progs, w.bc.live, P, T, N, *desc, size_parsimony, fitness, traj, steps_done, w.noise, s);
}

int mtgp_population_fitness(const float* nodes, int nodes_stride, const float* coef, int coef_stride, int P, int T,
                            int N, const mtgp_node_table* table, const mtgp_rollout_desc* desc,
                            float size_parsimony, float* fitness, unsigned long long* steps_done, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return mtgp_population_fitness_ex(nodes, nodes_stride, coef, coef_stride, P, T, N, table, desc, size_parsimony,
                                    fitness, nullptr, steps_done, workspace, workspace_bytes, stream);
}

int mtgp_eval_pointwise(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table, const float* X,
                        const float* Y, int n_points, int n_vars, float max_fitness, float* fitness, float* pred,
                        void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if ((rc = validate_table(table))) return rc;
  if (!X || !Y || !fitness || n_points < 1 || n_vars < 1) {
    mtgp_set_error("pointwise: X, Y, fitness must not be NULL; n_points, n_vars >= 1");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if ((rc = check_vars(table, n_vars))) return rc;
  const int tmp_base = n_vars + T;
  if (tmp_base + mtgp_tmp_bound(N) > MTGP_MAX_SLOTS) {
    mtgp_set_error("max_nodes=%d needs too many value-file slots", N);
    return MTGP_ERR_UNSUPPORTED;
  }
  const int npad = padded_points(n_points);
  Workspace w = carve(workspace, P, T, N, npad, n_vars);
  if (!workspace || workspace_bytes < w.bytes) {
    mtgp_set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return MTGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = mtgp_launch_transpose_points(X, Y, n_points, npad, n_vars, T, w.Xt, w.Yt, s))) return rc;
  if ((rc = mtgp_launch_lower(bc, P, T, N, table, tmp_base, mtgp_lstride(N), w.prog, w.meta, s))) return rc;
  mtgp_programs progs{w.prog, w.meta, mtgp_lstride(N), tmp_base, mtgp_tmp_bound(N), w.order};
  return mtgp_launch_pointwise(progs, nullptr, P, T, N, w.Xt, w.Yt, n_points, npad, n_vars, max_fitness, 0.0f,
                               fitness, pred, s);
}

int mtgp_population_fitness_pointwise(const float* nodes, int nodes_stride, const float* coef, int coef_stride,
                                      int P, int T, int N, const mtgp_node_table* table, const float* X,
                                      const float* Y, int n_points, int n_vars, float max_fitness,
                                      float size_parsimony, float* fitness, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if ((rc = validate_table(table))) return rc;
  if (!X || !Y || !fitness || !nodes || !coef || n_points < 1 || n_vars < 1) {
    mtgp_set_error("pointwise: NULL buffer or empty data set");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if ((rc = check_vars(table, n_vars))) return rc;
  const int tmp_base = n_vars + T;
  if (tmp_base + mtgp_tmp_bound(N) > MTGP_MAX_SLOTS) {
    mtgp_set_error("max_nodes=%d needs too many value-file slots", N);
    return MTGP_ERR_UNSUPPORTED;
  }
  const int npad = padded_points(n_points);
  Workspace w = carve(workspace, P, T, N, npad, n_vars);
  if (!workspace || workspace_bytes < w.bytes) {
    mtgp_set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return MTGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = mtgp_launch_encode(nodes, nodes_stride, coef, coef_stride, P, T, N, table, w.bc, s))) return rc;
  if ((rc = mtgp_launch_transpose_points(X, Y, n_points, npad, n_vars, T, w.Xt, w.Yt, s))) return rc;
  if ((rc = mtgp_launch_lower(w.bc, P, T, N, table, tmp_base, mtgp_lstride(N), w.prog, w.meta, s))) return rc;
  mtgp_programs progs{w.prog, w.meta, mtgp_lstride(N), tmp_base, mtgp_tmp_bound(N), w.order};
  return mtgp_launch_pointwise(progs, w.bc.live, P, T, N, w.Xt, w.Yt, n_points, npad, n_vars, max_fitness,
                               size_parsimony, fitness, nullptr, s);
}

// ---- host-buffer entry points: H2D copies + evaluation + D2H, one call ---------------------------
namespace {
// Scratch of the *_host entry points: one grow-only buffer per device, owned for the WHOLE call by the caller
// that holds the device's mutex -- two host threads evaluating on the same device are serialised (the calls
// synchronise their stream before returning, so the buffer is idle whenever the mutex is free and may be
// replaced by a larger one).  mtgp_release_scratch() frees the buffers.
struct DeviceScratch {
  std::mutex mu;
  void* ptr = nullptr;
  size_t bytes = 0;
  // the host entry points overlap the population upload with the evaluation (see mtgp_population_fitness_host):
  // one copy stream, two compute streams, events; created on first use, owned with the scratch
  cudaStream_t copy_s = nullptr, comp_s[2] = {nullptr, nullptr};
  cudaEvent_t ev[12] = {};
};
std::mutex g_scratch_table_mu;
std::map<int, std::unique_ptr<DeviceScratch>> g_scratch;

DeviceScratch* scratch_of_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(g_scratch_table_mu);
  auto& slot = g_scratch[dev];
  if (!slot) slot.reset(new DeviceScratch());
  return slot.get();
}

// call with s.mu held
int scratch_reserve(DeviceScratch& s, size_t bytes, void** out) {
  if (s.bytes < bytes) {
    if (s.ptr) MTGP_CUDA_TRY(cudaFree(s.ptr));  // idle: every host call synchronises before it releases the mutex
    s.ptr = nullptr;
    s.bytes = 0;
    MTGP_CUDA_TRY(cudaMalloc(&s.ptr, bytes));
    s.bytes = bytes;
  }
  *out = s.ptr;
  return MTGP_OK;
}

// ts must be finite and non-decreasing (diffrax requires it of SaveAt(ts)); host copy at hand in the *_host path
int validate_ts_host(const float* ts, int K) {
  if (K < 1 || !ts) { mtgp_set_error("ts: need at least one save point"); return MTGP_ERR_INVALID_ARGUMENT; }
  for (int k = 0; k < K; k++) {
    if (!(ts[k] - ts[k] == 0.0f) || (k > 0 && ts[k] < ts[k - 1])) {
      mtgp_set_error("ts[%d] = %g: save times must be finite and non-decreasing", k, (double)ts[k]);
      return MTGP_ERR_INVALID_ARGUMENT;
    }
  }
  return MTGP_OK;
}
}  // namespace

// call with s.mu held: the streams / events of the overlapped uploads
static int scratch_streams(DeviceScratch& sc) {
  if (sc.copy_s) return MTGP_OK;
  MTGP_CUDA_TRY(cudaStreamCreateWithFlags(&sc.copy_s, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) MTGP_CUDA_TRY(cudaStreamCreateWithFlags(&sc.comp_s[i], cudaStreamNonBlocking));
  for (int i = 0; i < 12; i++) MTGP_CUDA_TRY(cudaEventCreateWithFlags(&sc.ev[i], cudaEventDisableTiming));
  return MTGP_OK;
}

int mtgp_release_scratch(void) {
  std::lock_guard<std::mutex> lk(g_scratch_table_mu);
  for (auto& kv : g_scratch) {
    std::lock_guard<std::mutex> lk2(kv.second->mu);
    if (kv.second->ptr) cudaFree(kv.second->ptr);
    kv.second->ptr = nullptr;
    kv.second->bytes = 0;
    if (kv.second->copy_s) {
      cudaStreamDestroy(kv.second->copy_s);
      for (int i = 0; i < 2; i++) cudaStreamDestroy(kv.second->comp_s[i]);
      for (int i = 0; i < 12; i++) cudaEventDestroy(kv.second->ev[i]);
      kv.second->copy_s = nullptr;
    }
  }
  return MTGP_OK;
}

int mtgp_population_fitness_host(const float* population_host, int P, int T, int N, const mtgp_node_table* table,
                                 const mtgp_rollout_desc* dh, float size_parsimony, float* fitness_host,
                                 void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if (!dh || !population_host || !fitness_host) { mtgp_set_error("NULL host buffer"); return MTGP_ERR_INVALID_ARGUMENT; }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t pop_bytes = (size_t)P * T * N * 4 * sizeof(float);
  // populations worth overlapping (>= 32 MB of rows) are evaluated as a prologue slice + the rest (see below)
#ifndef MTGP_HOST_SPLIT_MIN_MB
#define MTGP_HOST_SPLIT_MIN_MB 32  // (8 measured on a 24 MB population: 12.8 vs 12.2 ms -- the second launch costs more than the upload it hides)
#endif
  const bool split = pop_bytes >= ((size_t)MTGP_HOST_SPLIT_MIN_MB << 20) && P >= 64;
  const size_t PA = split ? (size_t)((P / 8 + 7) & ~7) : 0;
  const size_t ws_a = split ? mtgp_workspace_bytes((int)PA, T, N, dh) : 0;
  const size_t ws_b = split ? mtgp_workspace_bytes((int)(P - PA), T, N, dh) : 0;
  const size_t ws_bytes = split ? align_up(ws_a, 256) + ws_b : mtgp_workspace_bytes(P, T, N, dh);
  const size_t B = dh->B, K = dh->K, D = dh->D;
  if ((rc = validate_ts_host(dh->ts, dh->K))) return rc;
  DeviceScratch* sc = scratch_of_current_device();
  if (!sc) { mtgp_set_error("cudaGetDevice failed"); return MTGP_ERR_CUDA; }
  std::lock_guard<std::mutex> scratch_lock(sc->mu);  // the scratch is this call's until it has synchronised
  if (split && (rc = scratch_streams(*sc))) return rc;
  // layout of the scratch: population | fitness | data arrays | workspace
  size_t off_pop = 0, off_fit, off_x0, off_ts, off_tg, off_keys, off_par, off_ys, off_ws, total;
  {
    size_t o = align_up(pop_bytes, 256);
    off_fit = o; o = align_up(o + (size_t)P * 4, 256);
    off_x0 = o; o = align_up(o + B * D * 4, 256);
    off_ts = o; o = align_up(o + K * 4, 256);
    off_tg = o; o = align_up(o + B * (size_t)dh->n_targets * 4, 256);
    off_keys = o; o = align_up(o + B * 8, 256);
    off_par = o; o = align_up(o + B * (size_t)dh->n_params * 4, 256);
    off_ys = o; o = align_up(o + (dh->ys_true ? B * K * D * 4 : 0), 256);
    off_ws = o; total = o + ws_bytes;
  }
  void* base_v = nullptr;
  if ((rc = scratch_reserve(*sc, total, &base_v))) return rc;
  unsigned char* base = static_cast<unsigned char*>(base_v);
  mtgp_rollout_desc dd = *dh;
  dd.params_host = dh->params;  // the host copy is at hand: trajectory-uniform parameters are detected
#define H2D(field, off, bytes)                                                                          \
  if (dh->field && (bytes) > 0) {                                                                       \
    MTGP_CUDA_TRY(cudaMemcpyAsync(base + off, dh->field, (bytes), cudaMemcpyHostToDevice, s));          \
    dd.field = reinterpret_cast<decltype(dd.field)>(base + off);                                        \
  } else {                                                                                              \
    dd.field = nullptr;                                                                                 \
  }
  H2D(x0, off_x0, B * D * 4)
  H2D(ts, off_ts, K * 4)
  H2D(targets, off_tg, B * (size_t)dh->n_targets * 4)
  H2D(obs_keys, off_keys, B * 8)
  H2D(params, off_par, B * (size_t)dh->n_params * 4)
  H2D(ys_true, off_ys, B * K * D * 4)
#undef H2D
  const float* pop_d = reinterpret_cast<const float*>(base + off_pop);
  float* fit_d = reinterpret_cast<float*>(base + off_fit);
  if (!split) {
    MTGP_CUDA_TRY(cudaMemcpyAsync(base + off_pop, population_host, pop_bytes, cudaMemcpyHostToDevice, s));
    rc = mtgp_population_fitness(pop_d, 4, pop_d + 3, 4, P, T, N, table, &dd, size_parsimony, fit_d, nullptr,
                                 base + off_ws, ws_bytes, stream);
    if (rc) return rc;
  } else {
    // Upload overlapped with evaluation: a small PROLOGUE slice of the population (1/8) is uploaded and starts
    // evaluating while the rest is still crossing PCIe; the two evaluations run on their own streams, so the warps
    // of the second fill the SM slots the first one frees (both kernels are persistent).  Individuals are
    // independent (genetic_programming.py:252 vmaps over them), so slicing changes no result.
    cudaStream_t cs = sc->copy_s, k0 = sc->comp_s[0], k1 = sc->comp_s[1];
    cudaEvent_t ev_data = sc->ev[0], ev_a = sc->ev[1], ev_b = sc->ev[2], done_a = sc->ev[3], done_b = sc->ev[4];
    const size_t row_bytes = (size_t)T * N * 4 * sizeof(float);
    MTGP_CUDA_TRY(cudaEventRecord(ev_data, s));  // the data arrays (and whatever the caller queued before) are on s
    MTGP_CUDA_TRY(cudaStreamWaitEvent(cs, ev_data, 0));
    MTGP_CUDA_TRY(cudaMemcpyAsync(base + off_pop, population_host, PA * row_bytes, cudaMemcpyHostToDevice, cs));
    MTGP_CUDA_TRY(cudaEventRecord(ev_a, cs));
    MTGP_CUDA_TRY(cudaMemcpyAsync(base + off_pop + PA * row_bytes, reinterpret_cast<const unsigned char*>(population_host) + PA * row_bytes,
                                  (size_t)(P - PA) * row_bytes, cudaMemcpyHostToDevice, cs));
    MTGP_CUDA_TRY(cudaEventRecord(ev_b, cs));
    MTGP_CUDA_TRY(cudaStreamWaitEvent(k0, ev_a, 0));
    rc = mtgp_population_fitness(pop_d, 4, pop_d + 3, 4, (int)PA, T, N, table, &dd, size_parsimony, fit_d, nullptr,
                                 base + off_ws, ws_a, k0);
    if (rc) { cudaStreamSynchronize(k0); return rc; }
    MTGP_CUDA_TRY(cudaEventRecord(done_a, k0));
    MTGP_CUDA_TRY(cudaStreamWaitEvent(k1, ev_b, 0));
    const float* pop_b = pop_d + PA * (size_t)T * N * 4;
    rc = mtgp_population_fitness(pop_b, 4, pop_b + 3, 4, (int)(P - PA), T, N, table, &dd, size_parsimony, fit_d + PA, nullptr,
                                 base + off_ws + align_up(ws_a, 256), ws_b, k1);
    if (rc) { cudaStreamSynchronize(k0); cudaStreamSynchronize(k1); return rc; }
    MTGP_CUDA_TRY(cudaEventRecord(done_b, k1));
    MTGP_CUDA_TRY(cudaStreamWaitEvent(s, done_a, 0));
    MTGP_CUDA_TRY(cudaStreamWaitEvent(s, done_b, 0));
  }
  MTGP_CUDA_TRY(cudaMemcpyAsync(fitness_host, fit_d, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
  MTGP_CUDA_TRY(cudaStreamSynchronize(s));
  return MTGP_OK;
}

int mtgp_population_fitness_pointwise_host(const float* population_host, int P, int T, int N,
                                           const mtgp_node_table* table, const float* X_host, const float* Y_host,
                                           int n_points, int n_vars, float max_fitness, float size_parsimony,
                                           float* fitness_host, void* stream) {
  int rc = validate_shape(P, T, N);
  if (rc) return rc;
  if (P == 0) return MTGP_OK;  // an empty population is a legal no-op (buffers may be NULL)
  if (!population_host || !fitness_host || !X_host || !Y_host || n_points < 1 || n_vars < 1) {
    mtgp_set_error("NULL host buffer or empty data set");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t pop_bytes = (size_t)P * T * N * 4 * sizeof(float);
  // populations of 32 MB and more: eight slices, uploaded back to back on a copy stream and evaluated on two compute
  // streams as they land (the kernel takes longer than the upload of a slice, so only the first upload is exposed)
  const bool split = pop_bytes >= ((size_t)32 << 20) && P >= 64;
#ifndef MTGP_PW_UPLOAD_SLICES
#define MTGP_PW_UPLOAD_SLICES 8  // (<= 8: one event per slice; C4 end to end: 2 slices 21.5 ms, 4: 19.9 ms, 8: 18.9 ms)
#endif
  const int n_chunks = split ? MTGP_PW_UPLOAD_SLICES : 1;
  const int Pc = split ? (((P + n_chunks - 1) / n_chunks + 7) & ~7) : P;
  const size_t ws_one = mtgp_workspace_bytes_pointwise(Pc, T, N, n_points, n_vars);
  const size_t ws_bytes = split ? 2 * align_up(ws_one, 256) : ws_one;
  size_t o = align_up(pop_bytes, 256);
  const size_t off_fit = o; o = align_up(o + (size_t)P * 4, 256);
  const size_t off_x = o; o = align_up(o + (size_t)n_points * n_vars * 4, 256);
  const size_t off_y = o; o = align_up(o + (size_t)n_points * T * 4, 256);
  const size_t off_ws = o;
  DeviceScratch* sc = scratch_of_current_device();
  if (!sc) { mtgp_set_error("cudaGetDevice failed"); return MTGP_ERR_CUDA; }
  std::lock_guard<std::mutex> scratch_lock(sc->mu);
  if (split && (rc = scratch_streams(*sc))) return rc;
  void* base_v = nullptr;
  if ((rc = scratch_reserve(*sc, o + ws_bytes, &base_v))) return rc;
  unsigned char* base = static_cast<unsigned char*>(base_v);
  MTGP_CUDA_TRY(cudaMemcpyAsync(base + off_x, X_host, (size_t)n_points * n_vars * 4, cudaMemcpyHostToDevice, s));
  MTGP_CUDA_TRY(cudaMemcpyAsync(base + off_y, Y_host, (size_t)n_points * T * 4, cudaMemcpyHostToDevice, s));
  const float* pop_d = reinterpret_cast<const float*>(base);
  float* fit_d = reinterpret_cast<float*>(base + off_fit);
  const float* X_d = reinterpret_cast<const float*>(base + off_x);
  const float* Y_d = reinterpret_cast<const float*>(base + off_y);
  if (!split) {
    MTGP_CUDA_TRY(cudaMemcpyAsync(base, population_host, pop_bytes, cudaMemcpyHostToDevice, s));
    rc = mtgp_population_fitness_pointwise(pop_d, 4, pop_d + 3, 4, P, T, N, table, X_d, Y_d, n_points, n_vars, max_fitness,
                                           size_parsimony, fit_d, base + off_ws, ws_bytes, stream);
    if (rc) return rc;
  } else {
    const size_t row_bytes = (size_t)T * N * 4 * sizeof(float);
    cudaStream_t cs = sc->copy_s;
    MTGP_CUDA_TRY(cudaEventRecord(sc->ev[0], s));  // X, Y (and whatever the caller queued before) are on s
    MTGP_CUDA_TRY(cudaStreamWaitEvent(cs, sc->ev[0], 0));
    for (int c = 0; c < n_chunks; c++) {
      const int p0 = c * Pc, pn = (P - p0 < Pc) ? P - p0 : Pc;
      if (pn <= 0) break;
      MTGP_CUDA_TRY(cudaMemcpyAsync(base + (size_t)p0 * row_bytes,
                                    reinterpret_cast<const unsigned char*>(population_host) + (size_t)p0 * row_bytes,
                                    (size_t)pn * row_bytes, cudaMemcpyHostToDevice, cs));
      MTGP_CUDA_TRY(cudaEventRecord(sc->ev[1 + c], cs));
    }
    for (int c = 0; c < n_chunks; c++) {
      const int p0 = c * Pc, pn = (P - p0 < Pc) ? P - p0 : Pc;
      if (pn <= 0) break;
      cudaStream_t k = sc->comp_s[c & 1];  // slices c and c + 2 share a stream and a workspace half, in order
      MTGP_CUDA_TRY(cudaStreamWaitEvent(k, sc->ev[1 + c], 0));
      const float* pc = pop_d + (size_t)p0 * T * N * 4;
      rc = mtgp_population_fitness_pointwise(pc, 4, pc + 3, 4, pn, T, N, table, X_d, Y_d, n_points, n_vars, max_fitness,
                                             size_parsimony, fit_d + p0, base + off_ws + (size_t)(c & 1) * align_up(ws_one, 256),
                                             ws_one, k);
      if (rc) { cudaStreamSynchronize(sc->comp_s[0]); cudaStreamSynchronize(sc->comp_s[1]); return rc; }
    }
    for (int i = 0; i < 2; i++) {
      MTGP_CUDA_TRY(cudaEventRecord(sc->ev[10 + i], sc->comp_s[i]));
      MTGP_CUDA_TRY(cudaStreamWaitEvent(s, sc->ev[10 + i], 0));
    }
  }
  MTGP_CUDA_TRY(cudaMemcpyAsync(fitness_host, fit_d, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
  MTGP_CUDA_TRY(cudaStreamSynchronize(s));
  return MTGP_OK;
}

}  // extern "C"
