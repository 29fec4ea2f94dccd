// mtgp_env.cuh -- device code shared by the rollout kernels: kernel parameters, mbarrier / bulk-copy
// helpers, jax.random (threefry, normal), and the environment functors (observation, drift, event,
// cost) restating MultiTreeGP/environments/control_environments/*.py.
#pragma once

#include "mtgp_internal.cuh"

namespace mtgp_dev {

#ifndef MTGP_ROLLOUT_WARPS
#define MTGP_ROLLOUT_WARPS 1
#endif
constexpr int kWarpsPerCta = MTGP_ROLLOUT_WARPS;  // one individual per warp; warps of a CTA share nothing but the launch
#ifndef MTGP_ROLLOUT_MIN_CTAS
#define MTGP_ROLLOUT_MIN_CTAS 32  // resident one-warp CTAs per SM the register allocation is tuned for (64 registers)
#endif

struct RolloutParams {
  const uint2* prog;
  const uint32_t* meta;
  const int32_t* live;
  const int32_t* order;  // schedule slot -> individual, costliest first (k_cost_order); NULL = identity
  int* next_slot;        // persistent warps draw their next schedule slot from this counter (NULL: one slot per warp)
  int lstride;
  int P, T;
  int kind;  // MTGP_EVAL_*
  int B, K, n_obs, n_targets, n_params, state_size, max_steps;
  float dt0, max_fitness, obs_noise, size_parsimony;
  int partitionable;
  const float *x0, *ts, *targets, *params, *ys_true;
  const uint32_t* obs_keys;
  float* fitness;
  float *xs_o, *ys_o, *us_o, *act_o, *tf_o;
  unsigned long long* steps_done;
  const float* noise;  // [cap_steps + K, n_obs, B] precomputed observation noise (NOISE == 1), see k_noise_table
  int cap_steps;       // rows of the step part of the noise table
  int fast_tmp, fast_th, want_fast;  // launch class: programs within / beyond the fast provisioning (temporaries, threaded instructions)
  int n_latent;        // state kept in the value file: SR -> T, DYNAMIC -> state_size, STATIC -> 0
  int slot_lat;        // data-vector slot of latent component 0
  int slot_u;          // data-vector slot of the control (control evaluators)
  int slot_target;     // data-vector slot of target 0
  int slot_out;        // n_latent slots: y1 of the latent state
  int slot_state;      // n_latent slots: master copy of the latent state
  int slots_per_warp;      // data + latent copies + temporaries + 1 scratch slot
  int th_cap;              // threaded-code instructions provisioned per warp (also bounds the compact staging area)
  float envc[12];          // UPAR kernels: the environment's derived constants (acrobot_derive), same for every trajectory
  // general Runge-Kutta path (mtgp_rollout_rk.cu): Dopri5 and/or PIDController
  int solver, controller;
  float rtol, atol, dtmin, safety, factormin, factormax, coeff1, coeff2, coeff3;
  int slot_rk;             // first of the 12 * (D + n_latent) slots: Y, Y1, YI, F0, KS[7], SV
};
constexpr int kMaxTrees = 32;
// launch classes of the rollout kernels: programs with at most kFastTmp temporaries and kFastTh threaded
// instructions per individual run with a small shared-memory footprint
constexpr int kFastTmp = 8;
constexpr int kFastTh = 136;  // full depth-6 trees (3 x 31 operators + the second readout copy) still fit; 32 CTAs x (6 KB + 1 KB reserved) per SM
// worst case of the threaded form: every compact instruction expands to MTGP_TH_MAX_EXPAND, plus the header
// and the ST per tree; the dynamic policy's readout is expanded twice
__host__ __device__ inline int mtgp_worst_threaded(int T, int lstride) { return (T + 1) * (MTGP_TH_MAX_EXPAND * lstride + 2); }

// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- per-warp shared memory of the rollout kernels ---------------------------------------------------
//   [compact programs, staged by cp.async.bulk | tree offsets | threaded-code programs | value file]
struct WarpArea {
  uint2* prog_s;     // th_cap compact instructions (sum over trees of (n + 2) & ~1 <= sum of n_threaded)
  int* woff;         // woff[t] = index of tree t's header in `th`
  mtgp_tinstr* th;   // th_cap threaded instructions
  float* vf;         // slots_per_warp * 32 floats; slot s of lane l at vf[s * 32 + l]
};
__host__ __device__ __forceinline__ size_t warp_area_bytes(int th_cap, int slots_per_warp) {
  return (size_t)th_cap * sizeof(uint2) + kMaxTrees * sizeof(int) + (size_t)th_cap * sizeof(mtgp_tinstr) +
         (size_t)slots_per_warp * 128;
}
__device__ __forceinline__ WarpArea warp_area(const RolloutParams& P, unsigned char* wbase) {
  WarpArea w;
  w.prog_s = reinterpret_cast<uint2*>(wbase);
  w.woff = reinterpret_cast<int*>(wbase + (size_t)P.th_cap * sizeof(uint2));
  w.th = reinterpret_cast<mtgp_tinstr*>(w.woff + kMaxTrees);
  w.vf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(w.th) + (size_t)P.th_cap * sizeof(mtgp_tinstr));
  return w;
}

// Threaded instructions an individual needs in shared memory (program layout below): every tree once,
// the dynamic policy's readout a second time (stand-alone copy for the save points), one ST per tree.
__device__ __forceinline__ int individual_th_need(const RolloutParams& P, int ind) {
  int total = 0;
  for (int t = 0; t < P.T; t++) total += (int)MTGP_META_TH(P.meta[(size_t)ind * P.T + t]) + 1;
  if (P.kind == MTGP_EVAL_DYNAMIC) total += (int)MTGP_META_TH(P.meta[(size_t)ind * P.T + P.state_size]);
  return total;
}
// Which launch owns an individual: by the temporaries and threaded instructions its programs need.
__device__ __forceinline__ bool individual_is_fast(const RolloutParams& P, int ind) {
  int need = 0;
  for (int t = 0; t < P.T; t++) need = max(need, (int)MTGP_META_TMP(P.meta[(size_t)ind * P.T + t]));
  return (need <= P.fast_tmp) && (individual_th_need(P, ind) <= P.fast_th);
}

// Stage the individual's compact programs (cp.async.bulk + mbarrier) and zero the value file.
__device__ __forceinline__ void warp_stage_programs(const RolloutParams& P, int ind, int lane, uint64_t* bar,
                                                    uint32_t& bar_phase, const WarpArea& w) {
  if (lane == 0) {
    uint32_t total = 0;
    for (int t = 0; t < P.T; t++) total += ((MTGP_META_N(P.meta[(size_t)ind * P.T + t]) + 2u) & ~1u) * 8u;
    mbar_expect_tx(bar, total);
    uint32_t soff = 0;  // compact programs are staged back to back (16-byte granules)
    for (int t = 0; t < P.T; t++) {
      const uint32_t n = MTGP_META_N(P.meta[(size_t)ind * P.T + t]);
      bulk_g2s(w.prog_s + soff, P.prog + ((size_t)ind * P.T + t) * P.lstride, ((n + 2u) & ~1u) * 8u, bar);
      soff += (n + 2u) & ~1u;
    }
  }
  __syncwarp();
  float* lb = w.vf + lane;
  for (int s = 0; s < P.slots_per_warp; s++) lb[s * 32] = 0.0f;
  mbar_wait(bar, bar_phase);
  bar_phase ^= 1u;
}

// Offset of tree t's compact program in the staging area.  meta_ind = P.meta + ind * T.
__device__ __forceinline__ uint32_t compact_offset(const uint32_t* meta_ind, int t) {
  uint32_t soff = 0;
  for (int q = 0; q < t; q++) soff += (MTGP_META_N(meta_ind[q]) + 2u) & ~1u;
  return soff;
}

// One threaded program = a CHAIN of trees run by a single interpreter call: [readout_tree,] first, first + 1, ...
// are expanded in turn; a tree's value is stored to its value-file slot (readout_store / store_slot0 + i, when
// >= 0) by giving its last instruction the ..._ST form (or an ST instruction).  The value of the last tree is
// also the call's result.  Writes [header]...[END] at dst, returns the instruction count.  Called by ONE lane;
// out of line: it runs once per individual and is expanded at several sites.
static __device__ __noinline__ int th_build_chain(const uint32_t* meta_ind, const uint2* prog_s, mtgp_tinstr* dst,
                                                  int first, int count, int store_slot0, int readout_tree,
                                                  int readout_store) {
  int m = 1;
  for (int i = (readout_tree >= 0 ? -1 : 0); i < count; i++) {
    const int t = (i < 0) ? readout_tree : first + i;
    const int store_slot = (i < 0) ? readout_store : (store_slot0 >= 0 ? store_slot0 + i : -1);
    const uint32_t soff = compact_offset(meta_ind, t);
    const int n = (int)MTGP_META_N(meta_ind[t]);  // without X_END
    for (int q = 0; q < n; q++) m += mtgp_th_expand(prog_s[soff + q], dst + m);
    if (n == 0) {  // (the lowering pass always emits at least a load)
      dst[m].next = TH_LOAD_I; dst[m].a = 0; dst[m].b = 0; dst[m].c = 0;
      m++;
    }
    if (store_slot >= 0) {
      mtgp_tinstr& last = dst[m - 1];
      if (last.next < TH_ST && (last.next & 1u) == 0u) {  // arithmetic / load without a store: fuse it
        last.next += 1u;
        last.c = (uint32_t)store_slot * 128u;
      } else {
        dst[m].next = TH_ST; dst[m].a = 0; dst[m].b = 0; dst[m].c = (uint32_t)store_slot * 128u;
        m++;
      }
    }
  }
  dst[m].next = TH_END; dst[m].a = 0; dst[m].b = 0; dst[m].c = 0;
  m++;
  mtgp_th_fuse(dst, m - 1);
  mtgp_th_link(dst, m - 1);
  return m;
}

// ---- jax.random on the device (THIRD-PARTY semantics; same as the oracle) -----
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& o0,
                                             uint32_t& o1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define TF_ROUND(r)                    \
  x0 += x1;                            \
  x1 = __funnelshift_l(x1, x1, r);     \
  x1 ^= x0;
  x0 += k0; x1 += k1;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
#undef TF_ROUND
  o0 = x0; o1 = x1;
}

__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;
  const float scale = 1.0f - lo;
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  float u = f * scale + lo;
  u = (u > lo) ? u : lo;
  return 1.41421356f * mtgp_erfinvf(u);
}

// Single out-of-line copies for the in-kernel noise generator (the adaptive kernel's hot loop is far larger
// than the instruction cache: five inlined threefry blocks and four erfinv polynomials are 1 000 instructions).
static __device__ __noinline__ uint2 threefry2x32_ool(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
  uint32_t o0, o1;
  threefry2x32(k0, k1, x0, x1, o0, o1);
  return make_uint2(o0, o1);
}
static __device__ __noinline__ float bits_to_normal_ool(uint32_t bits) { return bits_to_normal(bits); }

// normal(fold_in(key, bits(t)), (n,)) for n <= 4.  OOL: call the out-of-line helpers (same bits).
template <int NMAXOBS, bool OOL = false>
__device__ __forceinline__ void obs_normal(uint32_t key0, uint32_t key1, float t, int n, int partitionable,
                                           float* out) {
  auto tf = [](uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& o0, uint32_t& o1) {
    if (OOL) {
      const uint2 r = threefry2x32_ool(k0, k1, x0, x1);
      o0 = r.x; o1 = r.y;
    } else {
      threefry2x32(k0, k1, x0, x1, o0, o1);
    }
  };
  uint32_t f0, f1;
  tf(key0, key1, 0u, __float_as_uint(t), f0, f1);
  uint32_t bits[NMAXOBS];
  if (partitionable) {
#pragma unroll
    for (int i = 0; i < NMAXOBS; i++) {
      if (i < n) {
        uint32_t o0, o1;
        tf(f0, f1, 0u, (uint32_t)i, o0, o1);
        bits[i] = o0 ^ o1;
      }
    }
  } else {
    const int half = (n + 1) >> 1;
#pragma unroll
    for (int i = 0; i < (NMAXOBS + 1) / 2; i++) {
      if (i < half) {
        uint32_t o0, o1;
        const uint32_t c1 = (half + i < n) ? (uint32_t)(half + i) : 0u;
        tf(f0, f1, (uint32_t)i, c1, o0, o1);
#pragma unroll
        for (int j = 0; j < NMAXOBS; j++) {
          if (j == i) bits[j] = o0;
          if (j == half + i && j < n) bits[j] = o1;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NMAXOBS; i++)
    if (i < n) out[i] = OOL ? bits_to_normal_ool(bits[i]) : bits_to_normal(bits[i]);
}

// !isfinite(v) as ONE compare: |v| <= FLT_MAX is false for NaN and for +-inf
__device__ __forceinline__ bool nonfinite_cmp(float v) { return !(fabsf(v) <= 3.40282347e38f); }

// C @ x for C = eye(D)[:n_obs] and FINITE x, bit-identical to the literal dot product
//     y_i = ((1|0)*x_0 + (1|0)*x_1) + ...        (control_environment_base.py:46)
// every off-diagonal term is a zero with the sign of x_j, so y_i = x_i unless x_i is a zero, and a sum of
// zeros is -0 only if every term is -0:  y_i = x_i + z,  z = -0 if all x_j carry a sign bit, else +0.
template <int D>
__device__ __forceinline__ float obs_zero_term(const float* x) {
  uint32_t m = 0x80000000u;
#pragma unroll
  for (int j = 0; j < D; j++) m &= __float_as_uint(x[j]);
  return __uint_as_float(m);
}

// Acrobot.f_obs angle wrap (acrobot.py:31): jnp.remainder(o + pi, 2 pi) - pi.  t = o + pi.
// |t| < 2 pi: fmod returns t itself; a negative remainder gets the divisor added (python sign rule).
__device__ __forceinline__ float wrap_in_range(float t) {
  const float r = (t < 0.0f) ? t + 6.28318548f : t;
  return r - 3.14159274f;
}

__device__ __forceinline__ float clip_jnp(float v, float lo, float hi) {
  if (v != v) return v;
  v = v > lo ? v : lo;
  return v < hi ? v : hi;
}

__device__ __forceinline__ void sincos_contract(float x, float& s, float& c) {
  // identical to mtgp_sinf(x), mtgp_cosf(x): one shared argument reduction
  int q;
  const float r = mtgp_trig_reduce(x, &q);
  const float sk = mtgp_sin_kernel(r), ck = mtgp_cos_kernel(r);
  const float sv = (q & 1) ? ck : sk;
  const float cv = (q & 1) ? sk : ck;
  s = (q & 2) ? -sv : sv;
  c = ((q + 1) & 2) ? -cv : cv;
}

// Branch-free sin / cos for arguments the caller has range-checked (|x| < 1e5, warp-uniformly): the contract's
// Cody-Waite reduction, BOTH kernels, quadrant by select, sign by xor -- the same operations on the same values
// as mtgp_sinf / mtgp_cosf, hence the same bits, without per-call range branches and quadrant branches.
__device__ __forceinline__ float trig_pick(float sk, float ck, int n) {  // sin(r + n pi/2) from the two kernels
  const float v = (n & 1) ? ck : sk;
  return __uint_as_float(__float_as_uint(v) ^ (((uint32_t)n & 2u) << 30));
}
__device__ __forceinline__ float sin_checked(float x) {
  int q;
  const float r = mtgp_trig_reduce_cw(x, &q);
  return trig_pick(mtgp_sin_kernel(r), mtgp_cos_kernel(r), q);
}
__device__ __forceinline__ float cos_checked(float x) {
  int q;
  const float r = mtgp_trig_reduce_cw(x, &q);
  return trig_pick(mtgp_sin_kernel(r), mtgp_cos_kernel(r), q + 1);
}
__device__ __forceinline__ void sincos_checked(float x, float& s, float& c) {
  int q;
  const float r = mtgp_trig_reduce_cw(x, &q);
  const float sk = mtgp_sin_kernel(r), ck = mtgp_cos_kernel(r);
  s = trig_pick(sk, ck, q);
  c = trig_pick(sk, ck, q + 1);
}

// Cold paths of the Acrobot functor, OUT OF LINE: they run when an angle has left the contract's fast ranges
// (|theta| >= 4e4, |o + pi| >= 2 pi) -- practically never -- and would otherwise sit in the middle of the step
// loop, whose instruction-cache footprint is what limits the kernel (profiles/r02_*: `no_instruction` stalls).
#ifndef MTGP_COLD_NOINLINE
#define MTGP_COLD_NOINLINE 0  // measured on C5: out of line 76.9 ms, inline 76.3 ms
#endif
#if MTGP_COLD_NOINLINE
#define MTGP_COLD __noinline__
#else
#define MTGP_COLD __forceinline__
#endif
#ifndef MTGP_REACHED_FAST
#define MTGP_REACHED_FAST 0  // measured on C5: no gain (fewer instructions, larger loop)
#endif
struct Trig5 { float s2, c2, c12, s1, c1h; };
static __device__ MTGP_COLD Trig5 acrobot_trig_generic(float theta1, float theta2) {
  const float half_pi = 1.57079637f;
  Trig5 t;
  int q;  // sin / cos of theta2 share one reduction (same bits as mtgp_sinf / mtgp_cosf)
  const float r = mtgp_trig_reduce(theta2, &q);
  const float sk = mtgp_sin_kernel(r), ck = mtgp_cos_kernel(r);
  const float sv = (q & 1) ? ck : sk, cv = (q & 1) ? sk : ck;
  t.s2 = (q & 2) ? -sv : sv;
  t.c2 = ((q + 1) & 2) ? -cv : cv;
  t.c12 = mtgp_cosf((theta1 + theta2) - half_pi);
  t.s1 = mtgp_sinf(theta1);
  t.c1h = mtgp_cosf(theta1 - half_pi);
  return t;
}
static __device__ MTGP_COLD float wrap_generic(float o) {  // jnp.remainder(o + pi, 2 pi) - pi, any o
  return mtgp_remainderf_py(o + 3.14159274f, 6.28318548f) - 3.14159274f;
}
static __device__ MTGP_COLD bool acrobot_reached_generic(float theta1, float theta2) {
  return ((-mtgp_cosf(theta1)) - mtgp_cosf(theta1 + theta2)) > 1.5f;
}

// Loop-invariant sub-expressions of Acrobot.drift (acrobot.py:51-72), in the source's own grouping, so that
// the drift is the same sequence of float32 operations whether they are computed per lane on the device or
// once on the host (IEEE mul/add: same bits) for the trajectory-uniform-parameter kernels.
enum { AK_M2 = 0, AK_D1A, AK_D1B, AK_D1C, AK_D2A, AK_D2B, AK_PHI2, AK_P1A, AK_P1B, AK_P1C, AK_T2A, AK_DEN, AK_COUNT };
__host__ __device__ inline void acrobot_derive(float l1, float l2, float m1, float m2, float* k) {
  const float lc1 = 0.5f * l1, lc2 = 0.5f * l2, g = 9.81f, moi2 = 1.0f;
  k[AK_M2] = m2;
  k[AK_D1A] = m1 * (lc1 * lc1);
  k[AK_D1B] = (l1 * l1) + (lc2 * lc2);
  k[AK_D1C] = (2.0f * l1) * lc2;
  k[AK_D2A] = lc2 * lc2;
  k[AK_D2B] = l1 * lc2;
  k[AK_PHI2] = (m2 * lc2) * g;
  k[AK_P1A] = ((-m2) * l1) * lc2;
  k[AK_P1B] = ((2.0f * m2) * l1) * lc2;
  k[AK_P1C] = ((m1 * lc1) + (m2 * l1)) * g;
  k[AK_T2A] = (m2 * l1) * lc2;
  k[AK_DEN] = (m2 * (lc2 * lc2)) + moi2;
}

// ---- environments --------------------------------------------------------------------------------
template <int ENV>
struct EnvDim { static constexpr int D = 1, NT = 0; };  // D state components, NT targets (acrobot.py:11, harmonic_oscillator.py:12, reactor.py:11)
template <> struct EnvDim<MTGP_ENV_ACROBOT> { static constexpr int D = 4, NT = 0; };
template <> struct EnvDim<MTGP_ENV_HARMONIC> { static constexpr int D = 2, NT = 1; };
template <> struct EnvDim<MTGP_ENV_REACTOR> { static constexpr int D = 3, NT = 1; };

template <int ENV>
struct Env {
  static constexpr int D = EnvDim<ENV>::D;
  float p[ENV == MTGP_ENV_ACROBOT ? AK_COUNT : 8];  // Acrobot: the derived constants (acrobot_derive)
  float target;
  uint32_t key0, key1;
  // cost accumulators
  float sum, sum0, u_d;
  int fs;

  __device__ __forceinline__ void init(const RolloutParams& P, int b) {
    if (ENV == MTGP_ENV_ACROBOT) {
      acrobot_derive(P.params[b * P.n_params + 0], P.params[b * P.n_params + 1], P.params[b * P.n_params + 2],
                     P.params[b * P.n_params + 3], p);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = (i < P.n_params) ? P.params[b * P.n_params + i] : 0.0f;
    }
    target = (P.n_targets > 0) ? P.targets[b * P.n_targets] : 0.0f;
    key0 = P.obs_keys ? P.obs_keys[2 * b] : 0u;
    key1 = P.obs_keys ? P.obs_keys[2 * b + 1] : 0u;
    sum = 0.0f; sum0 = 0.0f; fs = -1;
    if (ENV == MTGP_ENV_HARMONIC) {
      // u_d = -pinv(b) @ A @ x_d, x_d = [target, 0]   (harmonic_oscillator.py:77-79)
      const float r0 = ((-0.0f) * 0.0f) + ((-1.0f) * (-p[0]));
      const float r1 = ((-0.0f) * 1.0f) + ((-1.0f) * (-p[1]));
      u_d = (r0 * target) + (r1 * 0.0f);
    }
  }

  // EnvironmentBase.f_obs (control_environment_base.py:43-48) + Acrobot.f_obs (acrobot.py:29-32)
  // normal(fold_in(key, bits(t)), (n_obs,)) @ W, W = obs_noise * eye (reactor.py:43 scales it)
  template <bool OOL = false>
  __device__ __forceinline__ void noise_term(float t, int n_obs, float obs_noise, int partitionable,
                                             float* out) const {
    float nrm[D];
    obs_normal<D, OOL>(key0, key1, t, n_obs, partitionable, nrm);
#pragma unroll
    for (int i = 0; i < D; i++) {
      if (i < n_obs) {
        float nacc = 0.0f;
#pragma unroll
        for (int j = 0; j < D; j++) {
          if (j < n_obs) {
            float w = (i == j) ? obs_noise : 0.0f;
            if (ENV == MTGP_ENV_REACTOR) w = w * (i == 2 ? 0.1f : 15.0f);
            const float term = nrm[j] * w;
            nacc = (j == 0) ? term : nacc + term;
          }
        }
        out[i] = nacc;
      }
    }
  }

  // out-of-line copy for the in-kernel generator (must not be replicated at every observation site), itself
  // built from single copies of threefry and erfinv
  struct Noise4 { float v[4]; };
  __device__ __noinline__ Noise4 noise_term_ool(float t, int n_obs, float obs_noise, int partitionable) const {
    float nt[D];
    noise_term<true>(t, n_obs, obs_noise, partitionable, nt);
    Noise4 r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.v[i] = (i < D) ? nt[i < D ? i : 0] : 0.0f;
    return r;
  }

  // NOISE: 0 none, 1 = noise term given in nz[0..3] (the caller read it from the precomputed table),
  //        2 = generated inline.  All three give the same bits.
  // FINITE_X: the caller guarantees (warp-uniformly) that every component of x is finite
  template <int NOISE, bool FINITE_X = false>
  __device__ __forceinline__ void f_obs(float t, const float* x, float* y, int n_obs, float obs_noise,
                                        int partitionable, const float* nz) const {
    float nt[D];
    if (NOISE == 2) {
      const Noise4 r = noise_term_ool(t, n_obs, obs_noise, partitionable);
#pragma unroll
      for (int i = 0; i < D; i++) nt[i] = r.v[i < 4 ? i : 3];
    }
    if (NOISE == 1) {
#pragma unroll
      for (int i = 0; i < D; i++) nt[i] = nz[i < 4 ? i : 3];
    }
    const float z = FINITE_X ? obs_zero_term<D>(x) : 0.0f;
#pragma unroll
    for (int i = 0; i < D; i++) {
      if (i < n_obs) {
        float acc = 0.0f;
        if (FINITE_X) {
          acc = x[i] + z;
        } else {
#pragma unroll
          for (int j = 0; j < D; j++) {  // C @ x, literally (0 * inf = NaN must survive)
            const float term = ((i == j) ? 1.0f : 0.0f) * x[j];
            acc = (j == 0) ? term : acc + term;
          }
        }
        if (NOISE != 0) acc = acc + nt[i];
        y[i] = acc;
      }
    }
    if (ENV == MTGP_ENV_ACROBOT) {
      const float pi = 3.14159274f, two_pi = 6.28318548f;
      const float w0 = y[0] + pi, w1 = (n_obs > 1) ? y[1] + pi : 0.0f;
      // warp-uniform choice: every lane's angles within one turn (the usual case) -> two adds each
      if (__all_sync(0xFFFFFFFFu, fabsf(w0) < two_pi && fabsf(w1) < two_pi)) {  // (every caller is warp-convergent)
        y[0] = wrap_in_range(w0);
        if (n_obs > 1) y[1] = wrap_in_range(w1);
      } else {  // (an exact three-subtraction path for |t| < 16 pi was measured: no gain over the generic remainder, 73.9 vs 73.5 ms)
#pragma unroll
        for (int i = 0; i < 2; i++)
          if (i < n_obs) y[i] = wrap_generic(y[i]);
      }
    }
  }

  // UPAR: every trajectory has the same parameters and the derived constants come from the kernel
  // parameters (constant bank operands, no registers) instead of this lane's p[].
  template <bool UPAR = false>
  __device__ __forceinline__ void drift(const float* s, float u_in, float* ds, const RolloutParams& P) const {
    if (ENV == MTGP_ENV_ACROBOT) {  // acrobot.py:51-72
      const float control = clip_jnp(u_in, -1.0f, 1.0f);
      const float theta1 = s[0], theta2 = s[1], theta1_dot = s[2], theta2_dot = s[3];
#define AK(i) (UPAR ? P.envc[i] : p[i])
      const float m2 = AK(AK_M2);
      const float moi1 = 1.0f, moi2 = 1.0f, half_pi = 1.57079637f;
      // the five trigonometric values: when every lane's angles are small enough that all four arguments are
      // inside the contract's fast range (|theta| < 4e4 => |theta1 + theta2 - pi/2| < 1e5), without range branches
      float s2, c2, c12, s1, c1h;
      if (__all_sync(0xFFFFFFFFu, fabsf(theta1) < 40000.0f && fabsf(theta2) < 40000.0f)) {
        sincos_checked(theta2, s2, c2);
        c12 = cos_checked((theta1 + theta2) - half_pi);
        s1 = sin_checked(theta1);
        c1h = cos_checked(theta1 - half_pi);
      } else {
        const Trig5 t = acrobot_trig_generic(theta1, theta2);
        s2 = t.s2; c2 = t.c2; c12 = t.c12; s1 = t.s1; c1h = t.c1h;
      }
      const float d1 = AK(AK_D1A) + m2 * (AK(AK_D1B) + AK(AK_D1C) * c2) + moi1 + moi2;
      const float d2 = m2 * (AK(AK_D2A) + AK(AK_D2B) * c2) + moi2;
      const float phi2 = AK(AK_PHI2) * c12;
      const float phi1 = (AK(AK_P1A) * (theta2_dot * theta2_dot)) * s2 -
                         ((AK(AK_P1B) * theta1_dot) * theta2_dot) * s1 + AK(AK_P1C) * c1h + phi2;
      const float theta2_acc =
          (((control + (d2 / d1) * phi1) - (AK(AK_T2A) * (theta1_dot * theta1_dot)) * s2) - phi2) /
          (AK(AK_DEN) - (d2 * d2) / d1);
#undef AK
      const float theta1_acc = (-((d2 * theta2_acc) + phi1)) / d1;
      ds[0] = theta1_dot; ds[1] = theta2_dot; ds[2] = theta1_acc; ds[3] = theta2_acc;
    } else if (ENV == MTGP_ENV_HARMONIC) {  // harmonic_oscillator.py:61-63,70-71
      const float A10 = -p[0], A11 = -p[1];
      ds[0] = ((0.0f * s[0]) + (1.0f * s[1])) + (0.0f * u_in);
      ds[1] = ((A10 * s[0]) + (A11 * s[1])) + (1.0f * u_in);
    } else if (ENV == MTGP_ENV_REACTOR) {  // reactor.py:78-87
      const float Vol = p[0], Cp = p[1], dHr = p[2], UA = p[3], q = p[4], Tf = p[5], Tcf = p[6], Volc = p[7];
      const float Tc = s[0], T = s[1], c = s[2];
      const float control = clip_jnp(u_in, 0.0f, 300.0f);
      const float mEaR = (float)(-72750.0 / 8.314);
      const float kT = 7.2e10f * mtgp_expf(mEaR / T);
      const float dc = ((q / Vol) * (1.0f - c)) - (kT * c);
      const float dT = (((q / Vol) * (Tf - T)) + ((((-dHr) / Cp) * kT) * c)) + (((UA / Vol) / Cp) * (Tc - T));
      const float dTc = ((control / Volc) * (Tcf - Tc)) + (((UA / Volc) / Cp) * (T - Tc));
      ds[0] = dTc; ds[1] = dT; ds[2] = dc;
    }
  }

  // cond_fn_nan on the env part of the state (the latent part is OR-ed in by the caller)
  __device__ __forceinline__ bool bad_state(const float* y) const {
    bool bad = false;
    if (ENV == MTGP_ENV_ACROBOT) {
      // !isfinite(v) | |v| > c  ==  !(|v| <= c): NaN and inf fail the comparison too
      bad = nonfinite_cmp(y[0]) | nonfinite_cmp(y[1]) | !(fabsf(y[2]) <= 25.1327419f) | !(fabsf(y[3]) <= 56.5486679f);
    } else {
#pragma unroll
      for (int i = 0; i < D; i++) bad |= nonfinite_cmp(y[i]);
    }
    return bad;
  }

  // one save point of env.fitness_function; fs_known >= 0 selects the exact second pass (Acrobot)
  __device__ __forceinline__ void cost_add(int k, float qk, const float* xs, float u, int fs_known) {
    if (ENV == MTGP_ENV_ACROBOT) {  // acrobot.py:77-84
      const float c = (u * 0.01f) * u;
      if (fs_known >= 0) {
        if (!(qk > (float)fs_known)) sum = sum + c;
        return;
      }
      // reached = -cos(theta1) - cos(theta1 + theta2) > 1.5 (acrobot.py:79).  A lane whose saved state is the inf
      // pre-fill (its event has fired) gives cos(inf) = NaN and "not reached" -- no reason to send the whole warp
      // through the Payne-Hanek call for it: when every lane is either non-finite or small, the branch-free path.
      const bool nonfin = nonfinite_cmp(xs[0]) | nonfinite_cmp(xs[1]);
      const bool small = fabsf(xs[0]) < 40000.0f && fabsf(xs[1]) < 40000.0f;
      bool reached;
      if (MTGP_REACHED_FAST && __all_sync(__activemask(), nonfin || small)) {  // (the adaptive kernel calls this under a per-lane condition)
        const float a0 = nonfin ? 0.0f : xs[0], a1 = nonfin ? 0.0f : xs[0] + xs[1];
        reached = !nonfin && (((-cos_checked(a0)) - cos_checked(a1)) > 1.5f);
      } else {
        reached = acrobot_reached_generic(xs[0], xs[1]);
      }
      if (!(qk > 0.0f)) sum0 = sum0 + c;
      if (fs < 0) {
        if (reached) {
          fs = k;
          if (!(qk > (float)k)) sum = sum + c;
        } else {
          sum = sum + c;
        }
      } else if (!(qk > (float)fs)) {
        sum = sum + c;
      }
    } else if (ENV == MTGP_ENV_HARMONIC) {  // harmonic_oscillator.py:76-81
      const float v0 = xs[0] - target, v1 = xs[1] - 0.0f;
      const float w0 = (v0 * 0.5f) + (v1 * 0.0f), w1 = (v0 * 0.0f) + (v1 * 0.0f);
      const float quad = (w0 * v0) + (w1 * v1);
      const float du = u - u_d;
      sum = sum + (quad + ((du * 0.5f) * du));
    } else if (ENV == MTGP_ENV_REACTOR) {  // reactor.py:92-96
      const float v[3] = {xs[0] - 0.0f, xs[1] - target, xs[2] - 0.0f};
      float w[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const float term = v[i] * ((i == j) ? (i == 1 ? 0.01f : 0.0f) : 0.0f);
          acc = (i == 0) ? term : acc + term;
        }
        w[j] = acc;
      }
      const float quad = ((w[0] * v[0]) + (w[1] * v[1])) + (w[2] * v[2]);
      sum = sum + (quad + ((u * 0.0001f) * u));
    }
  }

  __device__ __forceinline__ float cost_finish(int K, int fs_known) const {
    if (ENV == MTGP_ENV_ACROBOT) {
      if (fs_known >= 0) return (float)(fs_known + (fs_known == 0 ? K : 0)) + sum;
      if (fs < 0) return (float)K + sum0;
      return (float)(fs + (fs == 0 ? K : 0)) + sum;
    }
    return sum;
  }
};


}  // namespace mtgp_dev
