/*
 * mtgp_oracle.c -- CPU restatement of MultiTreeGP's population fitness-evaluation path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product (libmtgp_b200.so)
 * never links, imports or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference cannot be imported in this image (jax, diffrax, equinox are not
 * installed and there is no network) and it ships no tests, golden vectors or fixtures
 * (SURVEY.md section 4 / 8c).  Every function below restates reference source line by line and
 * cites it; semantics that live in third-party packages absent from /root/reference (diffrax's
 * fixed-step loop, jax.random's threefry/normal, XLA's gather clamping) are restated from their
 * published algorithms and marked THIRD-PARTY.  What pins this file instead: hand-computed
 * interpreter cases, Random123 threefry known-answer vectors, closed-form Euler recurrences and
 * a float64 cross-check -- see tests/test_oracle_*.py and tests/golden/.
 *
 * Arithmetic: float32 throughout (the reference never enables x64), compiled with
 * -ffp-contract=off so that no a*b+c is fused behind our back.  Elementary functions come from
 * include/mtgp_math.h (the numerical contract, bit-identical on CPU and GPU) or, with
 * oracle_set_math(1), from the host libm (independent cross-check).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/mtgp.h"
#include "../include/mtgp_math.h"

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* elementary functions: contract (0) or libm (1)                                              */
/* ------------------------------------------------------------------------------------------ */
static int g_math_mode = 0;
ORACLE_API void oracle_set_math(int mode) { g_math_mode = mode; }
ORACLE_API int oracle_get_math(void) { return g_math_mode; }
/* Reference-faithful WORK: under vmap, lax.switch over a batched index becomes a select over ALL
   node functions (genetic_programming.py:372), so XLA computes every branch of every row.  With mode 1 the
   interpreter does the same work -- every node function of the table on every row, the unselected results
   discarded -- and returns the same values.  Timing only (bench.py's cpu_baseline); default 0. */
static int g_all_branches = 0;
ORACLE_API void oracle_set_all_branches(int mode) { g_all_branches = mode; }

static inline float m_sin(float x) { return g_math_mode ? sinf(x) : mtgp_sinf(x); }
static inline float m_cos(float x) { return g_math_mode ? cosf(x) : mtgp_cosf(x); }
static inline float m_exp(float x) { return g_math_mode ? expf(x) : mtgp_expf(x); }
static inline float m_log(float x) { return g_math_mode ? logf(x) : mtgp_logf(x); }
static inline float m_tanh(float x) { return g_math_mode ? tanhf(x) : mtgp_tanhf(x); }
/* x ** c for the PID step-size factor: exp(c * log(x)) in contract mode, libm powf otherwise */
static inline float m_pow(float x, float c) { return g_math_mode ? powf(x, c) : mtgp_expf(c * mtgp_logf(x)); }
static inline float m_log1p(float x) { return g_math_mode ? log1pf(x) : mtgp_log1pf(x); }
static inline float m_remainder_py(float x, float y) {
  if (!g_math_mode) return mtgp_remainderf_py(x, y);
  float r = fmodf(x, y);
  if (r != 0.0f && ((r < 0.0f) != (y < 0.0f))) r = r + y;
  return r;
}
/* Giles single-precision erfinv, as in XLA's ErfInv32 (THIRD-PARTY) */
static float m_erfinv(float x) {
  if (!g_math_mode) return mtgp_erfinvf(x);
  float w = -log1pf(-x * x), p;
  if (fabsf(x) == 1.0f) return x * INFINITY;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w;
    p = -3.5233877e-06f + p * w;
    p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w;
    p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w;
    p = 0.00134934322f + p * w;
    p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w;
    p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  return p * x;
}

/* exported so tests can measure the contract functions against float64 */
ORACLE_API void oracle_math_apply(int fn, const float* x, const float* y, float* out, int n) {
  for (int i = 0; i < n; i++) {
    switch (fn) {
      case 0: out[i] = m_sin(x[i]); break;
      case 1: out[i] = m_cos(x[i]); break;
      case 2: out[i] = m_exp(x[i]); break;
      case 3: out[i] = m_log(x[i]); break;
      case 4: out[i] = m_tanh(x[i]); break;
      case 5: out[i] = m_log1p(x[i]); break;
      case 6: out[i] = m_erfinv(x[i]); break;
      case 7: out[i] = m_remainder_py(x[i], y[i]); break;
      default: out[i] = NAN;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* jax.random restatement (THIRD-PARTY: jax/_src/prng.py, jax/_src/random.py)                  */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3") */
ORACLE_API void oracle_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1,
                                    uint32_t* out) {
  static const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int g = 0; g < 5; g++) {
    const int* R = (g & 1) ? R1 : R0;
    for (int r = 0; r < 4; r++) {
      x0 += x1;
      x1 = rotl32(x1, R[r]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  out[0] = x0;
  out[1] = x1;
}

/* jax.random.fold_in(key, data): threefry_2x32(key, [0, uint32(data)]) */
ORACLE_API void oracle_fold_in(const uint32_t* key, uint32_t data, uint32_t* out) {
  oracle_threefry2x32(key[0], key[1], 0u, data, out);
}

/* 32 random bits for shape (n,):  partitionable -> counters (0, i), bits = o0 ^ o1;
   original -> counters iota(n) padded to even, split in halves, outputs concatenated. */
static void random_bits(const uint32_t* key, int n, int partitionable, uint32_t* bits) {
  uint32_t o[2];
  if (partitionable) {
    for (int i = 0; i < n; i++) {
      oracle_threefry2x32(key[0], key[1], 0u, (uint32_t)i, o);
      bits[i] = o[0] ^ o[1];
    }
  } else {
    int half = (n + 1) / 2;
    for (int i = 0; i < half; i++) {
      uint32_t c1 = (uint32_t)(half + i);
      if (half + i >= n) c1 = 0u; /* the zero pad */
      oracle_threefry2x32(key[0], key[1], (uint32_t)i, c1, o);
      bits[i] = o[0];
      if (half + i < n) bits[half + i] = o[1];
    }
  }
}

/* jax.random.normal(key, (n,), float32): sqrt(2) * erf_inv(uniform(lo, 1)), lo = nextafter(-1, 0) */
ORACLE_API void oracle_normal(const uint32_t* key, int n, int partitionable, float* out) {
  uint32_t bits[16];
  if (n > 16) n = 16;
  random_bits(key, n, partitionable, bits);
  const float lo = -0.99999994f; /* np.nextafter(np.float32(-1), np.float32(0)) */
  const float scale = 1.0f - lo; /* maxval - minval, rounds to 2.0f */
  for (int i = 0; i < n; i++) {
    float f = mtgp_u2f((bits[i] >> 9) | 0x3F800000u) - 1.0f;
    float u = f * scale + lo;
    u = (u > lo) ? u : lo; /* lax.max(minval, u); u is never NaN */
    out[i] = 1.41421356f * m_erfinv(u);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* tree interpreter: genetic_programming.py:356-388 (body_fun / foriloop), literal             */
/* ------------------------------------------------------------------------------------------ */
static inline int f2i_trunc(float v) {
  /* .astype(int) on float32: truncation toward zero; NaN/out-of-range are undefined in XLA,
     pick 0 / saturation (never produced by the reference's own operators) */
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 2147483647;
  if (v <= -2147483648.0f) return (-2147483647 - 1);
  return (int)v;
}
static inline int wrap_clamp(int i, int n) {
  /* NumPy-style negative wrap, then XLA gather clamping (THIRD-PARTY) */
  if (i < 0) i += n;
  if (i < 0) i = 0;
  if (i > n - 1) i = n - 1;
  return i;
}
static float apply_node(int opcode, int arg, float x, float y, const float* data) {
  switch (opcode) {
    case MTGP_OP_EMPTY: return 0.0f;  /* node_function_list[0]: lambda x, y, _data: 0.0  (:135) */
    case MTGP_OP_COEF: return 0.0f;   /* node_function_list[1] is also 0.0; the select at :372 picks the coefficient */
    case MTGP_OP_VAR: return data[arg]; /* lambda_leaf (:30-31) */
    case MTGP_OP_ADD: return x + y;
    case MTGP_OP_SUB: return x - y;
    case MTGP_OP_MUL: return x * y;
    case MTGP_OP_DIV: return x / y;
    case MTGP_OP_MAX: return mtgp_maximumf(x, y) /* jnp.maximum */;
    case MTGP_OP_MIN: return mtgp_minimumf(x, y) /* jnp.minimum */;
    case MTGP_OP_SIN: return m_sin(x);
    case MTGP_OP_COS: return m_cos(x);
    case MTGP_OP_EXP: return m_exp(x);
    case MTGP_OP_LOG: return m_log(x);
    case MTGP_OP_SQRT: return sqrtf(x);
    case MTGP_OP_TANH: return m_tanh(x);
    case MTGP_OP_NEG: return -x;
    case MTGP_OP_ABS: return fabsf(x);
    case MTGP_OP_SQUARE: return x * x;
    default: return 0.0f;
  }
}

/* One tree: fori_loop(0, max_nodes) over ALL rows (:387), result tree[-1,-1] (:388). */
ORACLE_API float oracle_tree_eval(const float* tree /* [N,4] */, int N, const mtgp_node_table* tab,
                                  const float* data, float* val /* scratch [N] */) {
  for (int i = 0; i < N; i++) val[i] = tree[4 * i + 3];
  for (int i = 0; i < N; i++) {
    float f_idx = tree[4 * i + 0], a_idx = tree[4 * i + 1], b_idx = tree[4 * i + 2];
    float coefficient = val[i];                          /* tree[i] row read (:368) */
    float x = val[wrap_clamp(f2i_trunc(a_idx), N)];      /* tree[a_idx.astype(int), 3] (:370) */
    float y = val[wrap_clamp(f2i_trunc(b_idx), N)];      /* (:371) */
    int f = f2i_trunc(f_idx);
    if (f < 0) f = 0;                                    /* lax.switch clamps its index */
    if (f > tab->n_ids - 1) f = tab->n_ids - 1;
    float v = apply_node(tab->opcode[f], tab->arg[f], x, y, data);
    if (g_all_branches) { /* the select over every branch: same value, all the work */
      for (int g = 0; g < tab->n_ids; g++) {
        volatile float w = apply_node(tab->opcode[g], tab->arg[g], x, y, data);
        if (g == f) v = w;
      }
    }
    val[i] = (f_idx == 1.0f) ? coefficient : v;          /* select(f_idx == 1, coefficient, switch(...)) (:372) */
  }
  return val[N - 1];
}

/* vmap_foriloop (:390-401): all trees of a candidate on the same data vector */
static void eval_trees(const float* cand, int t_begin, int t_end, int N, const mtgp_node_table* tab,
                       const float* data, float* out, float* scratch) {
  for (int t = t_begin; t < t_end; t++)
    out[t - t_begin] = oracle_tree_eval(cand + (size_t)t * N * 4, N, tab, data, scratch);
}

/* ------------------------------------------------------------------------------------------ */
/* environments: MultiTreeGP/environments/control_environments/                                */
/* ------------------------------------------------------------------------------------------ */
#define MAXD 8
#define MAXVARS 32

typedef struct env_t {
  int id, D, n_obs, n_targets, partitionable;
  float obs_noise;
  const float* p; /* per-trajectory parameters */
  uint32_t key[2];
  /* Acrobot derived (acrobot.py:34-43) */
  float l1, l2, m1, m2, lc1, lc2, moi1, moi2, g;
  /* Reactor derived (reactor.py:22-35) */
  float Vol, Cp, dHr, UA, q, Tf, Tcf, Volc;
  /* HarmonicOscillator (harmonic_oscillator.py:58-63) */
  float omega, zeta;
} env_t;

static void env_init(env_t* e, const mtgp_rollout_desc* d, int b) {
  e->id = d->env_id;
  e->D = d->D;
  e->n_obs = d->n_obs;
  e->n_targets = d->n_targets;
  e->partitionable = d->threefry_partitionable;
  e->obs_noise = d->obs_noise;
  e->p = d->params ? d->params + (size_t)b * d->n_params : NULL;
  e->key[0] = d->obs_keys ? d->obs_keys[2 * b] : 0u;
  e->key[1] = d->obs_keys ? d->obs_keys[2 * b + 1] : 0u;
  if (e->id == MTGP_ENV_ACROBOT) { /* initialize_parameters, acrobot.py:34-43 */
    e->l1 = e->p[0]; e->l2 = e->p[1]; e->m1 = e->p[2]; e->m2 = e->p[3];
    e->lc1 = 0.5f * e->l1;
    e->lc2 = 0.5f * e->l2;
    e->moi1 = e->moi2 = 1.0f;
    e->g = 9.81f;
  } else if (e->id == MTGP_ENV_HARMONIC) { /* harmonic_oscillator.py:58-61 */
    e->omega = e->p[0]; e->zeta = e->p[1];
  } else if (e->id == MTGP_ENV_REACTOR) { /* reactor.py:22-35 */
    e->Vol = e->p[0]; e->Cp = e->p[1]; e->dHr = e->p[2]; e->UA = e->p[3];
    e->q = e->p[4]; e->Tf = e->p[5]; e->Tcf = e->p[6]; e->Volc = e->p[7];
  }
}

/* EnvironmentBase.f_obs (control_environment_base.py:43-48): C@x + normal(fold_in(key, bits(t)))@W,
   C = eye(D)[:n_obs], W = obs_noise*eye(n_obs) (reactor.py:43 scales it by [15,15,0.1]).
   Acrobot.f_obs (acrobot.py:29-32) wraps the first two observations to (-pi, pi].
   The matrix products are evaluated literally (including the 0*x terms: 0*inf = NaN). */
static void env_f_obs(const env_t* e, float t, const float* x, float* out) {
  int n_obs = e->n_obs, D = e->D;
  float noise[MAXD];
  for (int i = 0; i < n_obs; i++) noise[i] = 0.0f;
  if (e->obs_noise != 0.0f) {
    uint32_t nk[2];
    float nrm[MAXD];
    oracle_fold_in(e->key, mtgp_f2u(t), nk); /* force_bitcast_convert_type(t) (:16-20,45) */
    oracle_normal(nk, n_obs, e->partitionable, nrm);
    static const float reactor_scale[3] = {15.0f, 15.0f, 0.1f};
    for (int i = 0; i < n_obs; i++) {
      float acc = 0.0f;
      for (int j = 0; j < n_obs; j++) {
        float w = (i == j) ? e->obs_noise : 0.0f;
        if (e->id == MTGP_ENV_REACTOR) w = w * reactor_scale[i];
        float term = nrm[j] * w;
        acc = (j == 0) ? term : acc + term;
      }
      noise[i] = acc;
    }
  }
  for (int i = 0; i < n_obs; i++) {
    float acc = 0.0f;
    for (int j = 0; j < D; j++) {
      float term = ((i == j) ? 1.0f : 0.0f) * x[j];
      acc = (j == 0) ? term : acc + term;
    }
    /* with obs_noise == 0 the noise term is exactly +-0 and x + 0 == x bitwise for x != 0 */
    out[i] = (e->obs_noise != 0.0f) ? acc + noise[i] : acc;
  }
  if (e->id == MTGP_ENV_ACROBOT) {
    const float pi = 3.14159274f, two_pi = 6.28318548f; /* f32(jnp.pi), f32(2*jnp.pi) */
    for (int i = 0; i < 2 && i < n_obs; i++) out[i] = m_remainder_py(out[i] + pi, two_pi) - pi;
  }
}

static inline float clipf(float v, float lo, float hi) {
  /* jnp.clip = minimum(maximum(v, lo), hi); NaN propagates */
  if (v != v) return v;
  v = v > lo ? v : lo;
  return v < hi ? v : hi;
}

static void env_drift(const env_t* e, const float* s, float u_in, float* ds) {
  if (e->id == MTGP_ENV_ACROBOT) { /* acrobot.py:51-72, operation order as written */
    float control = clipf(u_in, -1.0f, 1.0f);
    float theta1 = s[0], theta2 = s[1], theta1_dot = s[2], theta2_dot = s[3];
    float l1 = e->l1, m1 = e->m1, m2 = e->m2, lc1 = e->lc1, lc2 = e->lc2, g = e->g;
    const float half_pi = 1.57079637f; /* f32(jnp.pi/2) */
    float c2 = m_cos(theta2), s2 = m_sin(theta2);
    float d1 = m1 * (lc1 * lc1) + m2 * (((l1 * l1) + (lc2 * lc2)) + ((2.0f * l1) * lc2) * c2) + e->moi1 + e->moi2;
    float d2 = m2 * ((lc2 * lc2) + (l1 * lc2) * c2) + e->moi2;
    float phi2 = ((m2 * lc2) * g) * m_cos((theta1 + theta2) - half_pi);
    float phi1 = ((((-m2) * l1) * lc2) * (theta2_dot * theta2_dot)) * s2
               - (((((2.0f * m2) * l1) * lc2) * theta1_dot) * theta2_dot) * m_sin(theta1)
               + (((m1 * lc1) + (m2 * l1)) * g) * m_cos(theta1 - half_pi)
               + phi2;
    float theta2_acc = (((control + (d2 / d1) * phi1) - (((m2 * l1) * lc2) * (theta1_dot * theta1_dot)) * s2) - phi2)
                     / (((m2 * (lc2 * lc2)) + e->moi2) - (d2 * d2) / d1);
    float theta1_acc = (-((d2 * theta2_acc) + phi1)) / d1;
    ds[0] = theta1_dot; ds[1] = theta2_dot; ds[2] = theta1_acc; ds[3] = theta2_acc;
  } else if (e->id == MTGP_ENV_HARMONIC) { /* A@state + b@args, harmonic_oscillator.py:61-63,70-71 */
    float A10 = -e->omega, A11 = -e->zeta;
    ds[0] = ((0.0f * s[0]) + (1.0f * s[1])) + (0.0f * u_in);
    ds[1] = ((A10 * s[0]) + (A11 * s[1])) + (1.0f * u_in);
  } else if (e->id == MTGP_ENV_REACTOR) { /* reactor.py:78-87 */
    float Tc = s[0], T = s[1], c = s[2];
    float control = clipf(u_in, 0.0f, 300.0f);
    /* k(T) = k0*exp(-Ea/R/T): -72750/8.314 is folded in Python double, then cast (reactor.py:24-26,37) */
    const float mEaR = (float)(-72750.0 / 8.314);
    const float k0 = 7.2e10f, Cf = 1.0f;
    float kT = k0 * m_exp(mEaR / T);
    float dc = ((e->q / e->Vol) * (Cf - c)) - (kT * c);
    float dT = (((e->q / e->Vol) * (e->Tf - T)) + ((((-e->dHr) / e->Cp) * kT) * c))
             + (((e->UA / e->Vol) / e->Cp) * (Tc - T));
    float dTc = ((control / e->Volc) * (e->Tcf - Tc)) + (((e->UA / e->Volc) / e->Cp) * (T - Tc));
    ds[0] = dTc; ds[1] = dT; ds[2] = dc;
  }
}

/* cond_fn_nan: acrobot.py:86-87, harmonic_oscillator.py:83-84, reactor.py:98-99, SR_evaluator.py:93-94 */
static float env_cond(int env_id, const float* y, int n) {
  int bad = 0;
  for (int i = 0; i < n; i++) bad |= !mtgp_isfinitef(y[i]);
  if (env_id == MTGP_ENV_ACROBOT) {
    const float lim1 = 25.1327419f, lim2 = 56.5486679f; /* f32(8*pi), f32(18*pi) */
    bad |= (fabsf(y[2]) > lim1) | (fabsf(y[3]) > lim2);
  }
  return bad ? -1.0f : 1.0f;
}

/* env.fitness_function on the saved trajectory (xs [K,D], us [K]) */
static float env_fitness(const env_t* e, const float* xs, const float* us, const float* target,
                         const float* ts, int K) {
  int D = e->D;
  if (e->id == MTGP_ENV_ACROBOT) { /* acrobot.py:77-84 */
    int first_success = 0, found = 0;
    for (int k = 0; k < K && !found; k++) {
      float th1 = xs[k * D + 0], th2 = xs[k * D + 1];
      if (((-m_cos(th1)) - m_cos(th1 + th2)) > 1.5f) { first_success = k; found = 1; } /* argmax of bools */
    }
    float dts = ts[1] - ts[0];
    float sum = 0.0f;
    for (int k = 0; k < K; k++) {
      float control_cost = (us[k] * 0.01f) * us[k]; /* _u @ R @ _u */
      float c = ((ts[k] / dts) > (float)first_success) ? 0.0f : control_cost;
      sum = sum + c;
    }
    return (float)(first_success + (first_success == 0 ? K : 0)) + sum;
  } else if (e->id == MTGP_ENV_HARMONIC) { /* harmonic_oscillator.py:76-81, Q = diag(q,0), R = r, q = r = 0.5 */
    float xd0 = target[0], xd1 = 0.0f;
    /* u_d = -pinv(b) @ A @ x_d with pinv([[0],[1]]) = [[0,1]] (THIRD-PARTY: assumed exact) */
    float r0 = ((-0.0f) * 0.0f) + ((-1.0f) * (-e->omega));
    float r1 = ((-0.0f) * 1.0f) + ((-1.0f) * (-e->zeta));
    float u_d = (r0 * xd0) + (r1 * xd1);
    float sum = 0.0f;
    for (int k = 0; k < K; k++) {
      float v0 = xs[k * D + 0] - xd0, v1 = xs[k * D + 1] - xd1;
      float w0 = (v0 * 0.5f) + (v1 * 0.0f), w1 = (v0 * 0.0f) + (v1 * 0.0f); /* v.T @ Q */
      float quad = (w0 * v0) + (w1 * v1);
      float du = us[k] - u_d;
      float cost = quad + ((du * 0.5f) * du);
      sum = sum + cost;
    }
    return sum;
  } else { /* reactor.py:92-96, Q = diag(0,0.01,0), r = 1e-4 */
    float xd[3] = {0.0f, target[0], 0.0f};
    static const float Qd[3] = {0.0f, 0.01f, 0.0f};
    float sum = 0.0f;
    for (int k = 0; k < K; k++) {
      float v[3], w[3];
      for (int j = 0; j < 3; j++) v[j] = xs[k * D + j] - xd[j];
      for (int j = 0; j < 3; j++) {
        float acc = 0.0f;
        for (int i = 0; i < 3; i++) {
          float term = v[i] * ((i == j) ? Qd[i] : 0.0f);
          acc = (i == 0) ? term : acc + term;
        }
        w[j] = acc;
      }
      float quad = ((w[0] * v[0]) + (w[1] * v[1])) + (w[2] * v[2]);
      float cost = quad + ((us[k] * 0.0001f) * us[k]);
      sum = sum + cost;
    }
    return sum;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* diffrax.diffeqsolve(Euler, ConstantStepSize, SaveAt(ts), Event(cond_fn), throw=False)       */
/* THIRD-PARTY (diffrax/_integrate.py, _solver/euler.py, _step_size_controller/constant.py,    */
/* _local_interpolation.py) -- SURVEY.md App. C                                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct rollout_ctx {
  const mtgp_rollout_desc* d;
  const mtgp_node_table* tab;
  const float* cand; /* [T,N,4] */
  int T, N;
  env_t env;
  const float* target;
  float* scratch; /* [N] */
} rollout_ctx;

/* the vector field of the three evaluators */
static void vector_field(rollout_ctx* c, float t, const float* y, float* dy) {
  const mtgp_rollout_desc* d = c->d;
  float data[MAXVARS], out[MAXD];
  if (d->kind == MTGP_EVAL_SR) { /* SR_evaluator.py:85-88 */
    eval_trees(c->cand, 0, c->T, c->N, c->tab, y, dy, c->scratch);
    return;
  }
  int D = d->D, n_obs = d->n_obs, nt = d->n_targets;
  float obs[MAXD];
  env_f_obs(&c->env, t, y, obs);
  if (d->kind == MTGP_EVAL_STATIC) { /* feedforward_evaluate.py:104-110 */
    for (int i = 0; i < n_obs; i++) data[i] = obs[i];
    for (int i = 0; i < nt; i++) data[n_obs + i] = c->target[i];
    eval_trees(c->cand, 0, c->T, c->N, c->tab, data, out, c->scratch);
    env_drift(&c->env, y, out[0], dy); /* jnp.squeeze(u), n_control = 1 */
  } else { /* dynamic_evaluate.py:107-118 */
    int S = d->state_size, nc = c->T - S;
    const float* a = y + D;
    int n = 0;
    for (int i = 0; i < n_obs; i++) data[n++] = 0.0f;
    for (int i = 0; i < S; i++) data[n++] = a[i];
    for (int i = 0; i < nc; i++) data[n++] = 0.0f;
    for (int i = 0; i < nt; i++) data[n++] = c->target[i];
    float u[MAXD];
    eval_trees(c->cand, S, c->T, c->N, c->tab, data, u, c->scratch); /* readout (:113) */
    env_drift(&c->env, y, u[0], dy);                                 /* (:115) */
    n = 0;
    for (int i = 0; i < n_obs; i++) data[n++] = obs[i];
    for (int i = 0; i < S; i++) data[n++] = a[i];
    for (int i = 0; i < nc; i++) data[n++] = u[i];
    for (int i = 0; i < nt; i++) data[n++] = c->target[i];
    eval_trees(c->cand, 0, S, c->N, c->tab, data, dy + D, c->scratch); /* state equation (:116) */
  }
}

/* diffrax _clip_to_end for float32 times */
static inline float clip_to_end(float tprev, float tnext, float t1) {
  (void)tprev;
  return (tnext > t1 - 1e-6f) ? t1 : tnext; /* keep_step is always True for ConstantStepSize */
}

/* Step schedule shared by every trajectory (times do not depend on the state).  Exported so
   the tests can compare the device-side schedule kernel against it. */
ORACLE_API int oracle_schedule(const float* ts, int K, float dt0, int max_steps, float* tprev_out,
                               float* tnext_out, int cap) {
  float t0 = ts[0], t1 = ts[K - 1];
  float tprev = t0, tnext = t0 + dt0;
  tnext = tnext < t1 ? tnext : t1;
  int n = 0;
  while (tprev < t1 && n < max_steps) {
    if (n < cap) { tprev_out[n] = tprev; tnext_out[n] = tnext; }
    float new_tprev = tnext < t1 ? tnext : t1;
    float new_tnext = clip_to_end(new_tprev, tnext + dt0, t1);
    tprev = new_tprev;
    tnext = new_tnext;
    n++;
  }
  return n;
}

/* Dormand-Prince 5(4) tableau (diffrax/_solver/dopri5.py, THIRD-PARTY): a_lower, b_sol (= last row,
   FSAL), b_error = b_sol - b_hat, c, and the mid-point weights of its 4th-order dense output. */
static const float DP_A[6][6] = {
    {(float)(1.0 / 5), 0, 0, 0, 0, 0},
    {(float)(3.0 / 40), (float)(9.0 / 40), 0, 0, 0, 0},
    {(float)(44.0 / 45), (float)(-56.0 / 15), (float)(32.0 / 9), 0, 0, 0},
    {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), 0, 0},
    {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176), (float)(-5103.0 / 18656), 0},
    {(float)(35.0 / 384), 0.0f, (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784), (float)(11.0 / 84)}};
static const float DP_C[6] = {(float)(1.0 / 5), (float)(3.0 / 10), (float)(4.0 / 5), (float)(8.0 / 9), 1.0f, 1.0f};
static const float DP_BERR[7] = {(float)(35.0 / 384 - 1951.0 / 21600), 0.0f, (float)(500.0 / 1113 - 22642.0 / 50085),
                                 (float)(125.0 / 192 - 451.0 / 720), (float)(-2187.0 / 6784 - -12231.0 / 42400),
                                 (float)(11.0 / 84 - 649.0 / 6300), (float)(-1.0 / 60)};
static const float DP_CMID[7] = {(float)(6025192743.0 / 30085553152.0 / 2), 0.0f,
                                 (float)(51252292925.0 / 65400821598.0 / 2), (float)(-2691868925.0 / 45128329728.0 / 2),
                                 (float)(187940372067.0 / 1594534317056.0 / 2), (float)(-1776094331.0 / 19743644256.0 / 2),
                                 (float)(11237099.0 / 235043384.0 / 2)};

/* Solves one trajectory with diffrax.diffeqsolve semantics; ys_out [K, Dtot] pre-filled with inf.
   solver: Euler or Dopri5; controller: ConstantStepSize or PIDController.  Returns the number of
   ACCEPTED steps (rejected steps count towards max_steps like diffrax's num_steps). */
static int diffeqsolve(rollout_ctx* c, const float* y0, int Dtot, float* ys_out) {
  const mtgp_rollout_desc* d = c->d;
  const float* ts = d->ts;
  const int K = d->K;
  const int dopri = d->solver == MTGP_SOLVER_DOPRI5, pid = d->controller == MTGP_CONTROLLER_PID;
  const float safety = d->pid_safety > 0.0f ? d->pid_safety : 0.9f;
  const float factormin = d->pid_factormin > 0.0f ? d->pid_factormin : 0.2f;
  const float factormax = d->pid_factormax > 0.0f ? d->pid_factormax : 10.0f;
  const int default_pid = (d->pcoeff == 0.0f && d->icoeff == 0.0f && d->dcoeff == 0.0f);
  const float pc = default_pid ? 0.0f : d->pcoeff, ic = default_pid ? 1.0f : d->icoeff, dc = default_pid ? 0.0f : d->dcoeff;
  const float error_order = 5.0f; /* Dopri5.order */
  const float coeff1 = (ic + pc + dc) / error_order, coeff2 = -(pc + 2.0f * dc) / error_order, coeff3 = dc / error_order;
  float t0 = ts[0], t1 = ts[K - 1];
  float y[2 * MAXD], y1[2 * MAXD], yi[2 * MAXD], yerr[2 * MAXD], ymid[2 * MAXD];
  float f[7][2 * MAXD], ks[7][2 * MAXD];
  for (int i = 0; i < Dtot; i++) y[i] = y0[i];
  for (int i = 0; i < K * Dtot; i++) ys_out[i] = INFINITY;
  float tprev = t0;
  float tnext = t0 + d->dt0; /* controller.init: t0 + dt0 */
  tnext = tnext < t1 ? tnext : t1;
  float event_old = env_cond(d->env_id, y, Dtot); /* Event: cond_fn evaluated at (t0, y0) */
  int save_idx = 0, num_steps = 0, accepted = 0, fsal_valid = 0, at_dtmin = 0;
  float prev_inv_err = 1.0f, prev_prev_inv_err = 1.0f;
  while (tprev < t1 && num_steps < d->max_steps) {
    const float dt = tnext - tprev; /* terms.contr(t0, t1) */
    int keep = 1;
    float new_tprev, new_tnext;
    if (!dopri) {
      vector_field(c, tprev, y, f[0]);                                  /* Euler: f(t0, y0) */
      for (int i = 0; i < Dtot; i++) { ks[0][i] = dt * f[0][i]; y1[i] = y[i] + ks[0][i]; }
    } else {
      if (!fsal_valid) vector_field(c, tprev, y, f[0]);                  /* FSAL: k1 = last stage of the previous accepted step */
      for (int i = 0; i < Dtot; i++) ks[0][i] = dt * f[0][i];
      for (int st = 1; st < 7; st++) {
        for (int i = 0; i < Dtot; i++) {
          float acc = 0.0f;
          for (int j = 0; j < st; j++) {
            float term = DP_A[st - 1][j] * ks[j][i];
            acc = (j == 0) ? term : acc + term;
          }
          yi[i] = y[i] + acc;
        }
        if (st == 6) for (int i = 0; i < Dtot; i++) y1[i] = yi[i];       /* b_sol = a_lower[-1] */
        const float tst = (DP_C[st - 1] == 1.0f) ? tnext : tprev + DP_C[st - 1] * dt;
        vector_field(c, tst, yi, f[st]);
        for (int i = 0; i < Dtot; i++) ks[st][i] = dt * f[st][i];
      }
      for (int i = 0; i < Dtot; i++) {
        float acc = 0.0f;
        for (int j = 0; j < 7; j++) {
          float term = DP_BERR[j] * ks[j][i];
          acc = (j == 0) ? term : acc + term;
        }
        yerr[i] = acc;
      }
    }
    if (!pid) { /* ConstantStepSize.adapt_step_size: (True, t1, t1 + dt0) */
      new_tprev = tnext;
      new_tnext = tnext + d->dt0;
    } else {    /* PIDController.adapt_step_size */
      float sumsq = 0.0f;
      for (int i = 0; i < Dtot; i++) {
        float ay0 = fabsf(y[i]), ay1 = fabsf(y1[i]);
        float ym = (ay0 != ay0 || ay1 != ay1) ? NAN : (ay0 > ay1 ? ay0 : ay1);   /* jnp.maximum */
        float sc = yerr[i] / (d->atol + ym * d->rtol);
        sumsq = sumsq + sc * sc;
      }
      const float scaled_error = sqrtf(sumsq / (float)Dtot);             /* rms_norm */
      keep = (scaled_error < 1.0f) || at_dtmin;
      float inv_err = (scaled_error == 0.0f) ? INFINITY : 1.0f / scaled_error;
      float factor1 = (coeff1 == 0.0f) ? 1.0f : m_pow(inv_err, coeff1);
      float factor2 = (coeff2 == 0.0f) ? 1.0f : m_pow(prev_inv_err, coeff2);
      float factor3 = (coeff3 == 0.0f) ? 1.0f : m_pow(prev_prev_inv_err, coeff3);
      float fmin = keep ? 1.0f : factormin;
      float factor = ((safety * factor1) * factor2) * factor3;
      factor = clipf(factor, fmin, factormax);
      float ndt = dt * factor;
      if (d->dtmin > 0.0f) {                                              /* force_dtmin=True */
        at_dtmin = (ndt <= d->dtmin);
        ndt = (ndt != ndt) ? ndt : (ndt > d->dtmin ? ndt : d->dtmin);
      }
      new_tprev = keep ? tnext : tprev;
      new_tnext = new_tprev + ndt;
      if (keep) { prev_prev_inv_err = prev_inv_err; prev_inv_err = inv_err; }
      if (scaled_error != scaled_error) { /* NaN error estimate: no further step can be accepted; the
                                             solve ends here and the remaining save points keep inf */
        break;
      }
    }
    new_tprev = new_tprev < t1 ? new_tprev : t1;
    {
      const int clip = new_tnext > t1 - 1e-6f;                            /* _clip_to_end */
      const float tclip = keep ? t1 : new_tprev + 0.5f * (t1 - new_tprev);
      new_tnext = clip ? tclip : new_tnext;
    }
    num_steps++;
    if (keep) {
      if (dopri) { /* 4th-order dense output through (y0, f0, ymid, y1, f1) */
        for (int i = 0; i < Dtot; i++) {
          float acc = 0.0f;
          for (int j = 0; j < 7; j++) {
            float term = DP_CMID[j] * ks[j][i];
            acc = (j == 0) ? term : acc + term;
          }
          ymid[i] = y[i] + acc;
        }
      }
      while (save_idx < K && ts[save_idx] <= tnext) {                      /* SaveAt(ts) */
        float tk = ts[save_idx];
        float th = (tprev == tnext) ? 0.0f : (tk - tprev) / (tnext - tprev); /* linear_rescale */
        for (int i = 0; i < Dtot; i++) {
          if (!dopri) {
            ys_out[save_idx * Dtot + i] = y[i] + th * (y1[i] - y[i]);
          } else {
            float f0 = ks[0][i], f1 = ks[6][i];
            float a = ((2.0f * (f1 - f0)) - (8.0f * (y1[i] + y[i]))) + (16.0f * ymid[i]);
            float bq = ((((5.0f * f0) - (3.0f * f1)) + (18.0f * y[i])) + (14.0f * y1[i])) - (32.0f * ymid[i]);
            float cq = ((((f1 - (4.0f * f0)) - (11.0f * y[i])) - (5.0f * y1[i]))) + (16.0f * ymid[i]);
            ys_out[save_idx * Dtot + i] = ((((a * th + bq) * th + cq) * th + f0) * th) + y[i];
          }
        }
        save_idx++;
      }
      float event_new = env_cond(d->env_id, y1, Dtot);                     /* cond_fn(tnext, y1) */
      int fired = (event_old > 0.0f) != (event_new > 0.0f);
      event_old = event_new;
      for (int i = 0; i < Dtot; i++) { y[i] = y1[i]; f[0][i] = f[6][i]; }
      fsal_valid = dopri;
      accepted++;
      tprev = new_tprev;
      tnext = new_tnext;
      if (fired) break;
    } else {
      tprev = new_tprev; /* unchanged */
      tnext = new_tnext;
      /* f[0] stays valid: same (t0, y0) */
    }
  }
  return accepted;
}

/* Evaluator.evaluate_control_loop / evaluate_time_series for trajectory b of candidate `cand`.
   Optional outputs may be NULL.  Returns the per-trajectory fitness. */
static float eval_trajectory(const float* cand, int T, int N, const mtgp_node_table* tab,
                             const mtgp_rollout_desc* d, int b, float* xs_o, float* ys_o,
                             float* us_o, float* act_o, unsigned long long* steps_acc,
                             float* work /* K*(2*MAXD) + K*MAXD*2 + N */) {
  int K = d->K, D = d->D;
  rollout_ctx c;
  c.d = d; c.tab = tab; c.cand = cand; c.T = T; c.N = N;
  c.target = d->targets ? d->targets + (size_t)b * d->n_targets : NULL;
  float* sol = work;                 /* [K, Dtot] */
  float* xs = sol + K * 2 * MAXD;    /* [K, D] */
  float* us = xs + K * MAXD;         /* [K] (n_control = 1) */
  c.scratch = us + K * MAXD;
  float y0[2 * MAXD];
  int Dtot;
  if (d->kind == MTGP_EVAL_SR) {
    Dtot = D;
    for (int i = 0; i < D; i++) y0[i] = d->x0[(size_t)b * D + i];
    c.env.id = MTGP_ENV_NONE;
    int steps = diffeqsolve(&c, y0, Dtot, sol);
    if (steps_acc) *steps_acc += (unsigned long long)steps;
    /* fitness_function: mean(sum(square(pred - true), axis=-1)) (SR_evaluator.py:24) */
    const float* ytrue = d->ys_true + (size_t)b * K * D;
    float total = 0.0f;
    for (int k = 0; k < K; k++) {
      float row = 0.0f;
      for (int j = 0; j < D; j++) {
        float e = sol[k * D + j] - ytrue[k * D + j];
        row = row + e * e;
      }
      total = total + row;
    }
    if (xs_o) memcpy(xs_o, sol, sizeof(float) * K * D);
    return total / (float)K;
  }
  env_init(&c.env, d, b);
  int S = (d->kind == MTGP_EVAL_DYNAMIC) ? d->state_size : 0;
  Dtot = D + S;
  for (int i = 0; i < D; i++) y0[i] = d->x0[(size_t)b * D + i];
  for (int i = 0; i < S; i++) y0[D + i] = 0.0f; /* _x0 = concat([x0, zeros(state_size)]) (dynamic_evaluate.py:89) */
  int steps = diffeqsolve(&c, y0, Dtot, sol);
  if (steps_acc) *steps_acc += (unsigned long long)steps;
  /* post-pass at the save points (feedforward_evaluate.py:95-99, dynamic_evaluate.py:98-103) */
  int n_obs = d->n_obs, nt = d->n_targets, nc = T - S;
  for (int k = 0; k < K; k++) {
    float obs[MAXD], data[MAXVARS], u[MAXD];
    for (int j = 0; j < D; j++) xs[k * D + j] = sol[k * Dtot + j];
    env_f_obs(&c.env, d->ts[k], xs + k * D, obs);
    int n = 0;
    for (int i = 0; i < n_obs; i++) data[n++] = obs[i];
    if (d->kind == MTGP_EVAL_DYNAMIC) {
      for (int i = 0; i < S; i++) data[n++] = sol[k * Dtot + D + i];
      for (int i = 0; i < nc; i++) data[n++] = 0.0f;
    }
    for (int i = 0; i < nt; i++) data[n++] = c.target[i];
    eval_trees(cand, S, T, N, tab, data, u, c.scratch);
    us[k] = u[0];
    if (ys_o) for (int i = 0; i < n_obs; i++) ys_o[k * n_obs + i] = obs[i];
    if (us_o) for (int i = 0; i < nc; i++) us_o[k * nc + i] = u[i];
    if (act_o) for (int i = 0; i < S; i++) act_o[k * S + i] = sol[k * Dtot + D + i];
  }
  if (xs_o) memcpy(xs_o, xs, sizeof(float) * K * D);
  return env_fitness(&c.env, xs, us, c.target, d->ts, K);
}

/* Evaluator.__call__ for every individual of a population (the vmap at genetic_programming.py:252):
   per-trajectory fitness -> where(nan|inf, max_fitness) -> mean over B -> clip(0, max_fitness). */
ORACLE_API int oracle_rollout_population(const float* pop /* [P,T,N,4] */, int P, int T, int N,
                                         const mtgp_node_table* tab, const mtgp_rollout_desc* d,
                                         float* fitness /* [P] */, float* traj_fitness /* [P,B] or NULL */,
                                         float* xs_o, float* ys_o, float* us_o, float* act_o,
                                         unsigned long long* steps_done, int n_threads) {
  int B = d->B, K = d->K, D = d->D;
  int S = (d->kind == MTGP_EVAL_DYNAMIC) ? d->state_size : 0;
  int nc = (d->kind == MTGP_EVAL_SR) ? 0 : T - S;
  if (D > MAXD || S > MAXD || d->n_obs > MAXD || nc > MAXD) return -1;
  if (d->n_obs + S + nc + d->n_targets > MAXVARS) return -1;
  unsigned long long steps_total = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : steps_total)
  {
    float* work = (float*)malloc(sizeof(float) * ((size_t)K * 4 * MAXD + N + 16));
#pragma omp for schedule(dynamic, 4)
    for (int p = 0; p < P; p++) {
      const float* cand = pop + (size_t)p * T * N * 4;
      float sum = 0.0f;
      for (int b = 0; b < B; b++) {
        size_t pb = (size_t)p * B + b;
        float f = eval_trajectory(cand, T, N, tab, d, b,
                                  xs_o ? xs_o + pb * K * D : NULL,
                                  ys_o ? ys_o + pb * K * d->n_obs : NULL,
                                  us_o ? us_o + pb * K * nc : NULL,
                                  act_o ? act_o + pb * K * S : NULL, &steps_total, work);
        if (traj_fitness) traj_fitness[pb] = f;
        if (!mtgp_isfinitef(f)) f = d->max_fitness; /* where(isinf + isnan, max_fitness, f) */
        sum = sum + f;
      }
      float mean = sum / (float)B;
      fitness[p] = clipf(mean, 0.0f, d->max_fitness);
    }
    free(work);
  }
  if (steps_done) *steps_done = steps_total;
  return 0;
}

/* Pointwise regression (not in the reference, defined with its primitives):
   pred = vmap(vmap_foriloop, in_axes=[None,0])(candidate, X); fitness = mean(sum(square(pred-Y),-1))
   then the SR penalty/clip.  The sum over points is accumulated in float64. */
ORACLE_API int oracle_pointwise_population(const float* pop, int P, int T, int N,
                                           const mtgp_node_table* tab, const float* X, const float* Y,
                                           int n_points, int n_vars, float max_fitness, float* fitness,
                                           float* pred /* [P,n_points,T] or NULL */, int n_threads) {
  (void)n_vars;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    float* scratch = (float*)malloc(sizeof(float) * (N + 16));
#pragma omp for schedule(dynamic, 4)
    for (int p = 0; p < P; p++) {
      const float* cand = pop + (size_t)p * T * N * 4;
      double total = 0.0;
      for (int n = 0; n < n_points; n++) {
        float row = 0.0f;
        for (int t = 0; t < T; t++) {
          float v = oracle_tree_eval(cand + (size_t)t * N * 4, N, tab, X + (size_t)n * n_vars, scratch);
          if (pred) pred[((size_t)p * n_points + n) * T + t] = v;
          float e = v - Y[(size_t)n * T + t];
          row = row + e * e;
        }
        total += (double)row;
      }
      float f = (float)(total / (double)n_points);
      if (!mtgp_isfinitef(f)) f = max_fitness;
      fitness[p] = clipf(f, 0.0f, max_fitness);
    }
    free(scratch);
  }
  return 0;
}

/* parsimony term of evaluate_population (genetic_programming.py:424): count of rows with id != 0 */
ORACLE_API void oracle_live_count(const float* pop, int P, int T, int N, int32_t* live /* [P] */) {
  for (int p = 0; p < P; p++) {
    int c = 0;
    for (int i = 0; i < T * N; i++) c += pop[((size_t)p * T * N + i) * 4] != 0.0f;
    live[p] = c;
  }
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Test entry points: the environment functions alone.  tests/test_reference_shim_golden.py     */
/* compares them with outputs of the reference's OWN environment classes, recorded by           */
/* tools/record_reference_shim_golden.py (the reference imported from /root/reference over a    */
/* numpy stand-in for jax).  d / b: descriptor and trajectory whose parameters apply.            */
/* ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_env_drift(const mtgp_rollout_desc* d, int b, const float* x, float u, float* dx) {
  env_t e;
  env_init(&e, d, b);
  env_drift(&e, x, u, dx);
}
ORACLE_API void oracle_env_f_obs(const mtgp_rollout_desc* d, int b, float t, const float* x, float* y) {
  env_t e;
  env_init(&e, d, b);
  env_f_obs(&e, t, x, y);
}
ORACLE_API float oracle_env_cond(int env_id, const float* y, int n) { return env_cond(env_id, y, n); }
ORACLE_API float oracle_env_fitness(const mtgp_rollout_desc* d, int b, const float* xs, const float* us,
                                    const float* target, const float* ts, int K) {
  env_t e;
  env_init(&e, d, b);
  return env_fitness(&e, xs, us, target, ts, K);
}

/* the evaluators' vector field (_drift of feedforward_evaluate.py:104-110 / dynamic_evaluate.py:107-118 /
   SR_evaluator.py:85-88) at one point: y [D + state_size] -> dy, trajectory b of the descriptor */
ORACLE_API int oracle_vector_field(const float* cand /* [T,N,4] */, int T, int N, const mtgp_node_table* tab,
                                   const mtgp_rollout_desc* d, int b, float t, const float* y, float* dy) {
  rollout_ctx c;
  float scratch[MTGP_MAX_NODES + 1];
  if (N > MTGP_MAX_NODES) return -1;
  c.d = d; c.tab = tab; c.cand = cand; c.T = T; c.N = N;
  c.target = d->targets ? d->targets + (size_t)b * d->n_targets : NULL;
  c.scratch = scratch;
  if (d->kind != MTGP_EVAL_SR) env_init(&c.env, d, b);
  vector_field(&c, t, y, dy);
  return 0;
}
