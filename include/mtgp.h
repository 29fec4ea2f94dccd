/*
 * mtgp.h -- C ABI of libmtgp_b200.so: the B200-native replacement for MultiTreeGP's population
 * fitness-evaluation hot path.
 *
 * The reference has no FFI of its own (it is 100 % Python/JAX); its only plug point is the
 * `fitness_function` callable handed to GeneticProgramming (genetic_programming.py:63,249) and
 * invoked under vmap o shard_map o jit (genetic_programming.py:252,259-269,415).  The entry
 * points below are what an XLA-FFI custom call (jax.ffi) or a ctypes stub for that callable
 * binds; each one cites the reference code it replaces.  See INTEGRATION.md for the binding.
 *
 * Conventions
 *   - every pointer marked [device] is a CUDA device pointer valid on the current device,
 *     every pointer marked [host] is ordinary (preferably pinned) host memory;
 *   - all buffers are owned by the caller and only borrowed for the duration of the call
 *     (device work is enqueued on `stream` and may still be running when the call returns;
 *     the *_host entry points synchronise `stream` before returning);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 on success, a negative mtgp_status otherwise; mtgp_last_error() gives
 *     the message for the calling thread.  Numerical failure (NaN/inf trajectories) is NOT
 *     an error: it is the max_fitness penalty, exactly as in the reference
 *     (feedforward_evaluate.py:48-51, dynamic_evaluate.py:49-52, SR_evaluator.py:42-45);
 *   - nothing here ever falls back to a CPU implementation.
 */
#ifndef MTGP_H_
#define MTGP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTGP_ABI_VERSION 3

#if defined(__GNUC__)
#define MTGP_API __attribute__((visibility("default")))
#else
#define MTGP_API
#endif

typedef enum mtgp_status {
  MTGP_OK = 0,
  MTGP_ERR_INVALID_ARGUMENT = -1, /* shapes / sizes / NULL pointers / unknown opcode or env */
  MTGP_ERR_UNSUPPORTED = -2,      /* a valid request this build cannot run (e.g. N too large) */
  MTGP_ERR_WORKSPACE = -3,        /* workspace too small; see mtgp_workspace_bytes */
  MTGP_ERR_CUDA = -4              /* a CUDA runtime call failed; message has the detail */
} mtgp_status;

/* ---- node library ---------------------------------------------------------------------------
 * The reference assigns node ids in GeneticProgramming.__init__ (genetic_programming.py:132-199):
 * id 0 = empty row (evaluates to 0.0), id 1 = coefficient leaf, then the operators in
 * operator_list order (de-duplicated by display string), then the variables in variable_list
 * order (de-duplicated; the j-th new variable reads data[j], lambda_leaf at :30-31,:177).
 * There the node functions are opaque Python lambdas; here every node id is mapped to one
 * opcode of a closed instruction set.
 */
typedef enum mtgp_opcode {
  MTGP_OP_EMPTY = 0,  /* 0.0                                   (node id 0) */
  MTGP_OP_COEF = 1,   /* the row's coefficient (column 3)       (node id 1) */
  MTGP_OP_VAR = 2,    /* data[arg]                              lambda_leaf */
  MTGP_OP_ADD = 3,    /* x + y    x = value of row a, y = value of row b    */
  MTGP_OP_SUB = 4,    /* x - y */
  MTGP_OP_MUL = 5,    /* x * y */
  MTGP_OP_DIV = 6,    /* x / y    (unprotected: /0 -> inf/nan, as jnp.divide) */
  MTGP_OP_MAX = 7,    /* jnp.maximum (NaN if either is NaN) */
  MTGP_OP_MIN = 8,    /* jnp.minimum */
  MTGP_OP_SIN = 9,    /* unary ops use x only */
  MTGP_OP_COS = 10,
  MTGP_OP_EXP = 11,
  MTGP_OP_LOG = 12,
  MTGP_OP_SQRT = 13,
  MTGP_OP_TANH = 14,
  MTGP_OP_NEG = 15,
  MTGP_OP_ABS = 16,
  MTGP_OP_SQUARE = 17,
  MTGP_OP_COUNT_ = 18
} mtgp_opcode;

#define MTGP_MAX_NODE_IDS 64   /* F = 2 + n_operators + n_variables */
#define MTGP_MAX_NODES 512     /* max_nodes (rows per tree) */

typedef struct mtgp_node_table {
  int32_t n_ids;                          /* F */
  uint8_t opcode[MTGP_MAX_NODE_IDS];      /* mtgp_opcode of node id f */
  int32_t arg[MTGP_MAX_NODE_IDS];         /* MTGP_OP_VAR: index into the data vector */
} mtgp_node_table;

/* ---- level-1 bytecode: the population as structure-of-arrays -------------------------------
 * One program per tree per individual; row order is kept, which for every tree produced by the
 * reference's operators (initialization.py:82-98, mutation.py, crossover.py) is postfix order
 * (right subtree, left subtree, operator) in a right-aligned block -- SURVEY.md App. A.
 * Lossless: mtgp_decode reproduces the float32 rows bit for bit unless MTGP_BC_LOSSY is set.
 */
#define MTGP_BC_CANONICAL 1u /* live rows form a right-aligned postfix block: a = i-1, b = i-1-size(left) */
#define MTGP_BC_LOSSY 2u     /* some id/index field was non-integral or out of range (it is evaluated
                                with the reference's astype(int)/clamp semantics, but cannot be decoded) */

typedef struct mtgp_bytecode {
  uint8_t* op;    /* [device] [P,T,N] node id of each row, clamped to [0,F-1] like lax.switch */
  int16_t* a;     /* [device] [P,T,N] row index of the first operand  (column 1, -1 = none) */
  int16_t* b;     /* [device] [P,T,N] row index of the second operand (column 2, -1 = none) */
  float* coef;    /* [device] [P,T,N] column 3, verbatim bits */
  int32_t* live;  /* [device] [P,T]   rows with node id != 0 (the parsimony count, genetic_programming.py:424) */
  uint8_t* flags; /* [device] [P,T]   MTGP_BC_* */
} mtgp_bytecode;

/* ---- environments (reference: MultiTreeGP/environments/control_environments/) ------------- */
typedef enum mtgp_env_id {
  MTGP_ENV_NONE = 0,      /* SR evaluator: the trees ARE the vector field (SR_evaluator.py:85-88) */
  MTGP_ENV_ACROBOT = 1,   /* acrobot.py:7-87,    params per trajectory: l1, l2, m1, m2 */
  MTGP_ENV_HARMONIC = 2,  /* harmonic_oscillator.py:8-84,  params: omega, zeta */
  MTGP_ENV_REACTOR = 3    /* reactor.py:7-99,    params: Vol, Cp, dHr, UA, q, Tf, Tcf, Volc */
} mtgp_env_id;

typedef enum mtgp_solver { MTGP_SOLVER_EULER = 0, MTGP_SOLVER_DOPRI5 = 1 } mtgp_solver;
typedef enum mtgp_controller { MTGP_CONTROLLER_CONSTANT = 0, MTGP_CONTROLLER_PID = 1 } mtgp_controller;

typedef enum mtgp_evaluator_kind {
  MTGP_EVAL_SR = 0,       /* evaluators/SR_evaluator.py        max_fitness 1e5 */
  MTGP_EVAL_STATIC = 1,   /* evaluators/feedforward_evaluate.py max_fitness 1e4 */
  MTGP_EVAL_DYNAMIC = 2   /* evaluators/dynamic_evaluate.py     max_fitness 1e4 */
} mtgp_evaluator_kind;

/* Everything one fitness evaluation of a population needs besides the trees.  `data` in the
 * reference is the tuple (x0, ts, targets, process_noise_keys, obs_noise_keys, params) for the
 * control evaluators (feedforward_evaluate.py:62, dynamic_evaluate.py:63) and
 * (x0, ts, ys, noise_keys) for SR (SR_evaluator.py:55); process-noise keys are dead inputs
 * (only ODETerm(_drift) is solved) and are not passed. */
typedef struct mtgp_rollout_desc {
  int32_t kind;         /* mtgp_evaluator_kind */
  int32_t env_id;       /* mtgp_env_id (NONE for SR) */
  int32_t B;            /* trajectories (initial conditions) per individual */
  int32_t K;            /* len(ts): save points */
  int32_t D;            /* env state dimension (SR: = T) */
  int32_t n_obs;        /* observed components, C = eye(D)[:n_obs] */
  int32_t n_targets;    /* 0 (Acrobot) or 1 */
  int32_t n_params;     /* per-trajectory env parameters */
  int32_t state_size;   /* DYNAMIC: number of latent-state trees (the rest are readout trees) */
  int32_t max_steps;    /* diffrax max_steps (default 16**4) */
  float dt0;            /* fixed Euler step (diffrax.ConstantStepSize) */
  float max_fitness;    /* penalty / clip bound */
  float obs_noise;      /* W = obs_noise * I  (0 => noise skipped, bit-identical) */
  int32_t threefry_partitionable; /* jax_threefry_partitionable (default True since JAX 0.5) */
  /* solver / step-size controller.  All zero = the evaluators' class defaults, diffrax.Euler() +
   * ConstantStepSize() (SR_evaluator.py:21, feedforward_evaluate.py:11, dynamic_evaluate.py:11); the
   * shipped notebooks pass Dopri5() + PIDController(atol, rtol, dtmin=0.001) (the notebooks under examples/). */
  int32_t solver;        /* mtgp_solver */
  int32_t controller;    /* mtgp_controller (PID needs an embedded error estimate: Dopri5 only) */
  float rtol, atol;      /* PIDController tolerances */
  float dtmin;           /* PIDController dtmin with force_dtmin=True; <= 0: none */
  float pid_safety;      /* 0 -> diffrax default 0.9 */
  float pid_factormin;   /* 0 -> 0.2 */
  float pid_factormax;   /* 0 -> 10.0 */
  float pcoeff, icoeff, dcoeff; /* all 0 -> diffrax defaults (0, 1, 0): an I-controller */
  uint32_t flags;        /* MTGP_DESC_* */
  const float* x0;        /* [B, D] */
  const float* ts;        /* [K]    */
  const float* targets;   /* [B, n_targets] or NULL */
  const uint32_t* obs_keys; /* [B, 2] threefry keys (NULL allowed when obs_noise == 0) */
  const float* params;    /* [B, n_params] or NULL */
  const float* ys_true;   /* SR only: [B, K, D] recorded trajectories */
  /* [host] optional copy of `params` (same layout), or NULL.  The reference's own parameter samplers return
   * the same values for every trajectory (acrobot.py:24-27: ones); when the library can see that on the
   * host it runs kernels that take the environment constants from the constant bank instead of registers.
   * Results are bit-identical either way. */
  const float* params_host;
} mtgp_rollout_desc;

/* mtgp_rollout_desc.flags */
#define MTGP_DESC_NOISE_INLINE 1u /* generate the observation noise inside the rollout kernel instead of reading
                                     the per-evaluation table (same bits; exercised by the tests) */

/* Optional per-trajectory outputs of evaluate_candidate (feedforward_evaluate.py:53-62 returns
 * xs, ys, us, fitness; dynamic_evaluate.py:54-63 adds activities; SR_evaluator.py:47-55 returns
 * pred_ys).  Any pointer may be NULL. */
typedef struct mtgp_trajectory_out {
  float* xs;          /* [P, B, K, D]          saved env states (SR: predictions) */
  float* ys;          /* [P, B, K, n_obs]      observations at the save points */
  float* us;          /* [P, B, K, n_control]  controls at the save points */
  float* activities;  /* [P, B, K, state_size] latent states (DYNAMIC) */
  float* traj_fitness;/* [P, B]                per-trajectory fitness before penalty/mean/clip */
} mtgp_trajectory_out;

/* ---- library ------------------------------------------------------------------------------ */
MTGP_API int mtgp_abi_version(void);
MTGP_API const char* mtgp_last_error(void);

/* Bytes of device workspace the evaluation entry points need for a population of P individuals
 * with T trees of N rows (lowered programs, step schedule, counters).  rollout may be NULL. */
MTGP_API size_t mtgp_workspace_bytes(int P, int T, int N, const mtgp_rollout_desc* rollout);

MTGP_API size_t mtgp_workspace_bytes_pointwise(int P, int T, int N, int n_points, int n_vars);

/* Bytes a caller needs for its own mtgp_bytecode arrays (9 per row + 5 per tree). */
MTGP_API size_t mtgp_bytecode_bytes(int P, int T, int N);

/* K0 -- encoder.  Replaces nothing in the reference (there the f32[P,T,N,4] arrays are consumed
 * directly by body_fun, genetic_programming.py:356-376); it is step (1) of the north star.
 * `nodes` holds columns 0..2 and `coef` column 3 of every row; with the reference's combined
 * array pass nodes = pop, coef = pop + 3, both strides 4; with the split arrays shard_eval
 * produces (genetic_programming.py:261) pass strides 3 and 1.  [device] pointers. */
MTGP_API int mtgp_encode(const float* nodes, int nodes_stride, const float* coef, int coef_stride,
                int P, int T, int N, const mtgp_node_table* table, mtgp_bytecode out, void* stream);

/* Inverse of mtgp_encode: rebuilds f32[P,T,N,4] rows from the bytecode. [device] */
MTGP_API int mtgp_decode(mtgp_bytecode in, int P, int T, int N, float* trees_out, void* stream);

/* Lowering pass, exposed for tests and tools: level-1 bytecode -> executable accumulator-machine
 * programs (multitreegp_b200/csrc/mtgp_program.cuh).  prog_out [device] is [P*T, stride] 8-byte
 * instructions with stride = mtgp_program_stride(N); meta_out [device] uint32[P*T] holds
 * n_instructions (bits 0-9) | n_threaded_instructions << 10 (bits 10-21, the rollout kernels'
 * expanded form incl. header and END) | n_temporaries << 22.  tmp_base = first value-file slot free
 * for temporaries. */
MTGP_API int mtgp_lower(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table, int tmp_base,
                        void* prog_out, uint32_t* meta_out, void* stream);
MTGP_API int mtgp_program_stride(int N);

/* K1 -- pointwise tree evaluation (north-star item (2); oracle = vmap(vmap_foriloop) over data
 * rows, genetic_programming.py:390-401, with the SR penalty/mean/clip, SR_evaluator.py:42-45).
 *   X [device] f32[n_points, n_vars] row-major, Y [device] f32[n_points, T]
 *   fitness [device] f32[P]:  mean_n sum_t (tree_t(X[n]) - Y[n,t])^2, non-finite -> max_fitness, clip.
 *   pred (optional, may be NULL) [device] f32[P, n_points, T]. */
MTGP_API int mtgp_eval_pointwise(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table,
                        const float* X, const float* Y, int n_points, int n_vars,
                        float max_fitness, float* fitness, float* pred,
                        void* workspace, size_t workspace_bytes, void* stream);

/* K2 -- fused fixed-step rollout: tree interpreter + environment + Euler/SaveAt/Event + cost +
 * penalty/mean/clip.  Replaces Evaluator.__call__ vmapped over the population for the three
 * evaluators (SR_evaluator.py:30-45, feedforward_evaluate.py:36-51, dynamic_evaluate.py:37-52)
 * including diffrax.diffeqsolve(Euler, ConstantStepSize, SaveAt(ts), Event) inside them.
 * All pointers inside `desc` are [device].  fitness [device] f32[P].
 * `steps_done` (optional [device] uint64[1]) is incremented by the number of Euler steps
 * actually integrated over all trajectories. */
MTGP_API int mtgp_rollout(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table,
                 const mtgp_rollout_desc* desc, float* fitness, const mtgp_trajectory_out* traj,
                 unsigned long long* steps_done,
                 void* workspace, size_t workspace_bytes, void* stream);

/* The call an evaluator shard makes: encode + rollout + parsimony term
 *   fitness[p] += size_parsimony * sum_t live[p,t]      (genetic_programming.py:424)
 * on device buffers (what the XLA-FFI handler forwards to), and the same with HOST buffers
 * (population, data and fitness in host memory; H2D/D2H copies are part of the call) -- the
 * latter is what bench.py times as `e2e`.  The host form orders itself after the work already queued on `stream`,
 * synchronises before it returns, and may use library-owned streams in between: populations of 32 MB and more are
 * uploaded as a prologue slice + the rest and evaluated on two internal streams while the upload is still running
 * (individuals are independent, the results are the bits of the single-launch path); pinned host memory is what
 * makes that overlap real.  The scratch and the streams are per device (mtgp_release_scratch frees them). */
MTGP_API int mtgp_population_fitness(const float* nodes, int nodes_stride, const float* coef, int coef_stride,
                            int P, int T, int N, const mtgp_node_table* table,
                            const mtgp_rollout_desc* desc, float size_parsimony, float* fitness,
                            unsigned long long* steps_done,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Same as mtgp_population_fitness with the optional evaluate_candidate outputs. */
MTGP_API int mtgp_population_fitness_ex(const float* nodes, int nodes_stride, const float* coef, int coef_stride,
                                        int P, int T, int N, const mtgp_node_table* table,
                                        const mtgp_rollout_desc* desc, float size_parsimony, float* fitness,
                                        const mtgp_trajectory_out* traj, unsigned long long* steps_done,
                                        void* workspace, size_t workspace_bytes, void* stream);

MTGP_API int mtgp_population_fitness_host(const float* population_host /* [P,T,N,4] */, int P, int T, int N,
                                 const mtgp_node_table* table,
                                 const mtgp_rollout_desc* desc_host /* pointers inside are [host] */,
                                 float size_parsimony, float* fitness_host, void* stream);

/* Pointwise counterpart of the two calls above. */
MTGP_API int mtgp_population_fitness_pointwise(const float* nodes, int nodes_stride, const float* coef,
                                      int coef_stride, int P, int T, int N,
                                      const mtgp_node_table* table, const float* X, const float* Y,
                                      int n_points, int n_vars, float max_fitness,
                                      float size_parsimony, float* fitness,
                                      void* workspace, size_t workspace_bytes, void* stream);

MTGP_API int mtgp_population_fitness_pointwise_host(const float* population_host, int P, int T, int N,
                                           const mtgp_node_table* table, const float* X_host,
                                           const float* Y_host, int n_points, int n_vars,
                                           float max_fitness, float size_parsimony,
                                           float* fitness_host, void* stream);

/* Frees the device scratch buffers the *_host entry points keep per device (they are grow-only and reused
 * from call to call).  The *_host entry points are serialised per device by the library; the pointer-based
 * entry points above own no state and are re-entrant. */
MTGP_API int mtgp_release_scratch(void);

/* What the strategy does with the gathered fitness vector, without a host round trip:
 *   best_index = argmin(fitness) over the flattened population, best_fitness = fitness[best_index],
 *   best_solution = population[best_index]                   (genetic_programming.py:426-431)
 *   elite[s, :] = argsort(fitness[s, :])[:elite_size] per sub-population s, indices local to s
 *                                                             (reproduction.py:82)
 * with numpy / jax.numpy ordering: argmin returns the first index of the minimum and treats NaN as
 * the minimum; argsort is stable, -0 == +0, NaN last.  fitness [device] [num_populations *
 * population_size]; every output is optional (NULL) and [device]: best_index int32[1],
 * best_fitness f32[1], elite int32[num_populations, elite_size]; population [device]
 * [P, row_floats] and best_solution [device] [row_floats] go together (row_floats = T*N*4). */
MTGP_API int mtgp_select(const float* fitness, int num_populations, int population_size, int elite_size,
                         int32_t* best_index, float* best_fitness, int32_t* elite,
                         const float* population, int row_floats, float* best_solution, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTGP_H_ */
