// mtgp_rollout.cu -- K2: the fused fixed-step rollout kernel.
//
// One warp = one individual, lanes = trajectories (initial conditions): every lane of a warp runs
// the SAME tree programs, so interpreter dispatch is warp-uniform and divergence-free.  Fused in
// one kernel, with the state in registers / the warp's shared-memory value file for ALL steps:
//
//   diffrax.diffeqsolve(Euler, ConstantStepSize, SaveAt(ts), Event(cond_fn_nan), throw=False)
//        (call sites: SR_evaluator.py:76-79, feedforward_evaluate.py:90-93, dynamic_evaluate.py:93-96;
//         loop semantics: SURVEY.md App. C -- THIRD-PARTY, restated identically in oracle/mtgp_oracle.c)
//   _drift: env.f_obs -> trees -> env.drift          (feedforward_evaluate.py:104-110,
//                                                      dynamic_evaluate.py:107-118, SR_evaluator.py:85-88)
//   post-pass at the save points + env.fitness_function
//                                                     (feedforward_evaluate.py:95-99, dynamic_evaluate.py:98-103)
//   where(nan|inf, max_fitness) -> mean over B -> clip (feedforward_evaluate.py:48-51 and siblings)
//   + size_parsimony * live rows                      (genetic_programming.py:424)
//
// Programs are staged into shared memory with cp.async.bulk (TMA 1-D bulk copy) + mbarrier.
// Arithmetic order follows the reference source literally (see the oracle for the same
// expressions); the file is compiled with -fmad=false so nothing is contracted.
// The fixed-step kernel inlines the interpreter at its three call sites (measured on the C5 workload, one
// warp per CTA: 130 -> 122 ms); the adaptive kernel keeps it out of line for code size.
#ifndef MTGP_INTERP_INLINE
#define MTGP_INTERP_INLINE __forceinline__
#endif
#include "mtgp_env.cuh"

using namespace mtgp_dev;

int mtgp_launch_rollout_rk(RolloutParams rp, int kind, int env_id, int tmp_base, int bound, cudaStream_t stream);

namespace {

// ---- observation-noise table -------------------------------------------------------------------------
// The noise term of f_obs depends only on (trajectory key, time): it is the same for every individual
// of the population.  Two tiny kernels compute it once per evaluation:
//   k_schedule     walks the diffrax time loop (float32, same arithmetic as the rollout kernel) and
//                  records tprev of every step;
//   k_noise_table  evaluates normal(fold_in(key_b, bits(t))) @ W for every (step | save point, b).
// Layout: row r = step s (r < cap) or save point k (r = cap + k); element [r][b][i], i padded to 4.
__global__ void k_schedule(const float* __restrict__ ts, int K, float dt0, int max_steps, int cap,
                           float* __restrict__ tprev_out, int* __restrict__ n_steps_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const float t0 = ts[0], t1 = ts[K - 1];
  float tprev = t0, tnext = t0 + dt0;
  tnext = tnext < t1 ? tnext : t1;
  int n = 0;
  while (tprev < t1 && n < max_steps && n < cap) {
    tprev_out[n] = tprev;
    const float new_tprev = tnext < t1 ? tnext : t1;
    float new_tnext = tnext + dt0;
    new_tnext = (new_tnext > t1 - 1e-6f) ? t1 : new_tnext;
    tprev = new_tprev;
    tnext = new_tnext;
    n++;
  }
  *n_steps_out = n;
}

template <int ENV>
__global__ void __launch_bounds__(256) k_noise_table(const RolloutParams P, const float* __restrict__ tprev,
                                                    const int* __restrict__ n_steps, float* __restrict__ table) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = P.cap_steps + P.K;
  if (idx >= (long long)rows * P.B) return;
  const int r = (int)(idx / P.B), b = (int)(idx % P.B);
  float t;
  if (r < P.cap_steps) {
    if (r >= *n_steps) return;
    t = tprev[r];
  } else {
    t = P.ts[r - P.cap_steps];
  }
  Env<ENV> env;
  env.key0 = P.obs_keys[2 * b];
  env.key1 = P.obs_keys[2 * b + 1];
  float nt[Env<ENV>::D];
  env.noise_term(t, P.n_obs, P.obs_noise, P.partitionable, nt);
#pragma unroll
  for (int i = 0; i < Env<ENV>::D; i++)
    if (i < P.n_obs) table[((size_t)r * P.B + b) * 4 + i] = nt[i];
}

// ---- the kernel ----------------------------------------------------------------------------------
// SPEC: launch-time facts promoted to compile time for the configurations the notebooks use
//   0 generic; 1 every state component is observed (n_obs == D); 2 = 1 + a dynamic policy with two latent
//   states (DynamicPolicy.ipynb: state_size = 2).  The per-component predicates and the latent loops vanish.
// TRAJ_OUT: the evaluate_candidate outputs (xs, ys, us, activities, per-trajectory fitness) are requested.
// Population evaluation never asks for them, and without their code the hot loop is 120 instructions shorter.
// UPAR: the environment parameters are the same for every trajectory (the reference's samplers: acrobot.py:24-27)
// and their derived constants are kernel parameters (constant bank) instead of per-lane registers.
template <int KIND, int ENV, int NOISE, int SPEC, bool TRAJ_OUT, bool UPAR>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MTGP_ROLLOUT_MIN_CTAS) k_rollout(const RolloutParams P) {
  constexpr int D = (KIND == MTGP_EVAL_SR) ? 0 : EnvDim<ENV>::D;
  constexpr int DR = D > 0 ? D : 1;
  constexpr bool FULLOBS = SPEC >= 1;
  constexpr bool FIXED_LATENT = (KIND == MTGP_EVAL_STATIC) || (KIND == MTGP_EVAL_DYNAMIC && SPEC == 2);
  const int n_obs = FULLOBS ? D : P.n_obs;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wic = (kWarpsPerCta == 1) ? 0 : (int)(threadIdx.x >> 5);
  const int slot0 = blockIdx.x * kWarpsPerCta + wic;

  // shared memory: [mbarriers][per warp: compact programs | tree offsets | threaded programs | value file]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  const WarpArea wa =
      warp_area(P, smem_raw + 16 * kWarpsPerCta + (size_t)wic * warp_area_bytes(P.th_cap, P.slots_per_warp));
  // Persistent warps: every warp draws schedule slots (costliest individual first) from a global counter
  // until the population is exhausted -- no CTA turnover between individuals.
  uint64_t* bar = bars + 2 * wic;
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  uint32_t bar_phase = 0;
  bool first_round = true;
  for (;;) {
    int slot = slot0;  // the first individual comes with the launch; later ones from the counter
    if (!first_round) {
      if (!P.next_slot) return;
      if (lane == 0) slot = (int)(gridDim.x * kWarpsPerCta) + atomicAdd(P.next_slot, 1);
      slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
    }
    first_round = false;
    if (slot >= P.P) return;  // whole warp leaves together (no block-wide barrier anywhere)
    const int ind = P.order ? P.order[slot] : slot;
    if (individual_is_fast(P, ind) != (P.want_fast != 0)) continue;  // the other launch's individual

    // ---- stage this individual's programs (cp.async.bulk -> shared, mbarrier) ----
    warp_stage_programs(P, ind, lane, bar, bar_phase, wa);
    float* lb = wa.vf + lane;  // lane base of the value file (R = 1)
    mtgp_saddr lbytes = mtgp_saddr_of(lb);
    // opaque to the compiler: it would otherwise re-derive this address (special registers, four uniform
    // operations) at every use in the step loop instead of keeping it in a register
    asm volatile("mov.u32 %0, %0;" : "+r"(lbytes));
    bool readout_reads_y = false;
    if (KIND == MTGP_EVAL_DYNAMIC) {
      // does the readout read an observation or the control slot?  (conservative: unused operand fields count)
      const uint32_t soff = compact_offset(P.meta + (size_t)ind * P.T, P.state_size);
      const int n = (int)MTGP_META_N(P.meta[(size_t)ind * P.T + P.state_size]);
      for (int i = lane; i < n; i += 32) {
        const uint2 c = wa.prog_s[soff + i];
        const uint32_t xk = (c.x >> 4) & 3u, yk = (c.x >> 6) & 3u, xs = (c.x >> 16) & 0xFFu, ys = c.x >> 24;
        const bool rx = xk == OK_MEM && ((int)xs < n_obs || (int)xs == P.slot_u);
        const bool ry = yk == OK_MEM && ((int)ys < n_obs || (int)ys == P.slot_u);
        readout_reads_y |= (rx || ry);
      }
    }
    // The dynamic readout sees zeros for the observations and the control inside the step
    // (dynamic_evaluate.py:113) and the saved observations in the post-pass (:100-101).  With the
    // notebooks' variable lists it reads neither (only the latent state), and the slots need no care.
    const bool readout_sees_y = (KIND == MTGP_EVAL_DYNAMIC) && (__any_sync(0xFFFFFFFFu, readout_reads_y) != 0);
    // ---- expand to threaded code.  One interpreter call per Euler step runs the whole CHAIN of trees:
    //   static : [readout]                                          -> u
    //   dynamic: [readout -> slot_u][state eq j -> slot_out + j]...   (readout first: the state equations read u)
    //            split in two calls when the readout reads y / u (it must see zeros there, dynamic_evaluate.py:113)
    //   SR     : [tree j -> slot_out + j]...
    // plus a stand-alone readout for the save points of the dynamic policy.
    if (lane == 0) {
      const uint32_t* mi = P.meta + (size_t)ind * P.T;
      int off = 0;
      if (KIND == MTGP_EVAL_STATIC) {
        wa.woff[0] = wa.woff[1] = wa.woff[2] = 0;
        off += th_build_chain(mi, wa.prog_s, wa.th, 0, 1, -1, -1, -1);
      } else if (KIND == MTGP_EVAL_SR) {
        wa.woff[0] = wa.woff[1] = wa.woff[2] = 0;
        off += th_build_chain(mi, wa.prog_s, wa.th, 0, P.T, P.slot_out, -1, -1);
      } else {
        wa.woff[2] = 0;  // stand-alone readout
        off += th_build_chain(mi, wa.prog_s, wa.th, P.state_size, 1, -1, -1, -1);
        wa.woff[0] = readout_sees_y ? 0 : off;
        wa.woff[1] = off;
        // the state equations, preceded by the readout (-> slot_u) unless it has to see zeros
        off += th_build_chain(mi, wa.prog_s, wa.th + off, 0, P.state_size, P.slot_out,
                              readout_sees_y ? -1 : P.state_size, P.slot_u);
      }
    }
    __syncwarp();

    const float* ts = P.ts;
    const int K = P.K;
    const int NL = FIXED_LATENT ? (KIND == MTGP_EVAL_STATIC ? 0 : 2) : P.n_latent;
    // slot map of the data vector (mtgp_rollout_slot_map); compile-time for SPEC == 2: [y(D) | a(2) | u | targets | out(2) | state(2)]
    constexpr bool FIXED_SLOTS = (KIND == MTGP_EVAL_DYNAMIC && SPEC == 2);
    constexpr int NT = (KIND == MTGP_EVAL_SR) ? 0 : EnvDim<ENV>::NT;
    const int slot_lat = FIXED_SLOTS ? D : P.slot_lat, slot_u = FIXED_SLOTS ? D + 2 : P.slot_u,
              slot_target = FIXED_SLOTS ? D + 3 : P.slot_target, slot_out = FIXED_SLOTS ? D + 3 + NT : P.slot_out,
              slot_state = FIXED_SLOTS ? D + 5 + NT : P.slot_state;
    // shared-space addresses of the latent-state slots of this lane (component j is 128 bytes further)
    const mtgp_saddr a_state = lbytes + 128u * (uint32_t)slot_state, a_out = lbytes + 128u * (uint32_t)slot_out,
                     a_lat = lbytes + 128u * (uint32_t)slot_lat;
    const float t0 = ts[0], t1 = ts[K - 1];
    const float dts = (K > 1) ? ts[1] - ts[0] : 0.0f;  // ts[1]-ts[0], acrobot.py:82
    const mtgp_saddr th_s = mtgp_saddr_of(wa.th);
    const mtgp_saddr prog_step = th_s + 16u * (uint32_t)wa.woff[0];   // the step's chain (first call)
    const mtgp_saddr prog_step2 = th_s + 16u * (uint32_t)wa.woff[1];  // dynamic, readout_sees_y: the state equations
    const mtgp_saddr prog_save = th_s + 16u * (uint32_t)wa.woff[2];   // stand-alone readout (save points)
    float fit_sum = 0.0f;
    unsigned int my_steps = 0;  // Euler steps integrated by this lane (< 2^32: max_steps is capped at 2^22 per solve)

    for (int b0 = 0; b0 < P.B; b0 += 32) {  // B > 32: the warp takes 32 trajectories at a time
      const int b = b0 + lane;
      const bool lane_valid = b < P.B;
      const int bb = lane_valid ? b : P.B - 1;  // idle lanes shadow a valid trajectory
      Env<ENV> env;
      if (KIND != MTGP_EVAL_SR) env.init(P, bb);
      if (KIND != MTGP_EVAL_SR)
        for (int i = 0; i < P.n_targets; i++) lb[(slot_target + i) * 32] = P.targets[bb * P.n_targets + i];
      const float* nz_base = (NOISE == 1) ? P.noise + (size_t)bb * 4 : nullptr;
      const int nzs = P.B * 4;  // floats per table row

      int fs_known = -1;
      bool viol = false;
      float traj_fitness = 0.0f;
      for (int pass = 0; pass < 2; pass++) {
        // ---------------- one trajectory solve ----------------
        float x[DR], x1[DR];
        // FIXED_SLOTS: the latent state lives in registers (aS current, aO candidate); slot_lat of the value
        // file holds what the trees read.  Otherwise it lives in the value file (slot_state / slot_out).
        float aS[2] = {0.0f, 0.0f}, aO[2] = {0.0f, 0.0f};
        if (KIND != MTGP_EVAL_SR) {
#pragma unroll
          for (int i = 0; i < D; i++) { x[i] = P.x0[bb * D + i]; x1[i] = x[i]; }
        }
        for (int j = 0; j < NL; j++) {
          const float v = (KIND == MTGP_EVAL_SR) ? P.x0[bb * NL + j] : 0.0f;  // a0 = 0 (dynamic_evaluate.py:89)
          lb[(slot_state + j) * 32] = v;
          lb[(slot_lat + j) * 32] = v;
          lb[(slot_out + j) * 32] = v;
        }
        if (KIND != MTGP_EVAL_SR) { env.sum = 0.0f; env.sum0 = 0.0f; env.fs = -1; }
        float sr_total = 0.0f;
        bool ev_old_bad;
        bool x_nonfinite = false;  // some component of the env state x is NaN/inf (C @ x must then be literal)
        {
          bool bad = false;
          if (KIND != MTGP_EVAL_SR) {
            bad = env.bad_state(x);
#pragma unroll
            for (int i = 0; i < D; i++) x_nonfinite |= nonfinite_cmp(x[i]);
          }
          for (int j = 0; j < NL; j++) bad |= !mtgp_isfinitef(lb[(slot_state + j) * 32]);
          ev_old_bad = bad;  // Event: cond_fn at (t0, y0)
        }
        // a lane beyond B shadows a valid trajectory and never integrates: "dead" from the start, so that the
        // loop's one ballot covers validity too (its outputs are discarded below)
        bool alive = lane_valid;
        float tprev = t0, tnext = t0 + P.dt0;
        tnext = tnext < t1 ? tnext : t1;
        int steps = 0, k = 0;
        unsigned int lane_steps = 0;  // steps this lane integrated before its event fired
        float ts_k = ts[0];  // ts[k], reloaded only when k advances
        // NOISE == 1: the table row of the coming step, fetched one step ahead (an L2 hit is ~250 cycles)
        float4 nz_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (NOISE == 1) nz_next = __ldg(reinterpret_cast<const float4*>(nz_base));
        // One iteration = one Euler step while any lane is still integrating, then the save points that
        // step covers.  Once the solve is over (t1 reached, max_steps, or every lane's event fired) the
        // remaining save points are emitted from the inf pre-fill of SaveAt -- by the same code.
        while (k < K) {
          const bool stepping = (tprev < t1) && (steps < P.max_steps) &&
                                (__ballot_sync(0xFFFFFFFFu, alive) != 0u);
          float save_until = mtgp_inff();
          if (stepping) {
            const float dt = tnext - tprev;
            float u = 0.0f;
            if (KIND != MTGP_EVAL_SR) {
              float y[DR];
              // a dead lane's x is frozen at its last committed (finite) value; only a non-finite x0 takes the literal path
              const float nzv[4] = {nz_next.x, nz_next.y, nz_next.z, nz_next.w};
              if (__any_sync(0xFFFFFFFFu, x_nonfinite))
                env.template f_obs<NOISE, false>(tprev, x, y, n_obs, P.obs_noise, P.partitionable, nzv);
              else
                env.template f_obs<NOISE, true>(tprev, x, y, n_obs, P.obs_noise, P.partitionable, nzv);
              if (KIND == MTGP_EVAL_STATIC) {
#pragma unroll
                for (int i = 0; i < D; i++)
                  if (i < n_obs) mtgp_sts_f32(lbytes + 128u * (uint32_t)(i), y[i]);
              } else if (readout_sees_y) {
#pragma unroll
                for (int i = 0; i < D; i++)
                  if (i < n_obs) mtgp_sts_f32(lbytes + 128u * (uint32_t)(i), 0.0f);
                mtgp_sts_f32(lbytes + 128u * (uint32_t)(slot_u), 0.0f);
              }
              if (KIND == MTGP_EVAL_DYNAMIC && !readout_sees_y) {
                // the whole chain in one call: the observations are already in place (the readout does not
                // read them), the readout's ST puts u in its slot before the state equations run
#pragma unroll
                for (int i = 0; i < D; i++)
                  if (i < n_obs) mtgp_sts_f32(lbytes + 128u * (uint32_t)(i), y[i]);
                mtgp_run_threaded(prog_step, lbytes);
                u = mtgp_lds_f32(lbytes + 128u * (uint32_t)(slot_u));
              } else {
                u = mtgp_run_threaded(prog_step, lbytes);  // feedforward :107 / dynamic readout :113
              }
              // the next step's noise row, most of a step ahead of its use (after the interpreter call: a call
              // waits for every outstanding load)
              if (NOISE == 1 && steps + 1 < P.cap_steps)
                nz_next = __ldg(reinterpret_cast<const float4*>(nz_base + (size_t)(steps + 1) * nzs));
              float dx[DR];
              env.template drift<UPAR>(x, u, dx, P);
#pragma unroll
              for (int i = 0; i < D; i++) x1[i] = x[i] + dt * dx[i];
              if (KIND == MTGP_EVAL_DYNAMIC && readout_sees_y) {
#pragma unroll
                for (int i = 0; i < D; i++)
                  if (i < n_obs) mtgp_sts_f32(lbytes + 128u * (uint32_t)(i), y[i]);
                mtgp_sts_f32(lbytes + 128u * (uint32_t)(slot_u), u);
                mtgp_run_threaded(prog_step2, lbytes);
              }
            } else {
              mtgp_run_threaded(prog_step, lbytes);  // SR: the vector field, component j in slot_out + j
            }
            // latent Euler update (state equation :116 / SR vector field): f_j is in slot_out + j
            if (FIXED_SLOTS) {
#pragma unroll
              for (int j = 0; j < 2; j++) aO[j] = aS[j] + dt * mtgp_lds_f32(a_out + 128u * j);
            } else {
              for (int j = 0; j < NL; j++)
                mtgp_sts_f32(a_out + 128u * j, mtgp_lds_f32(a_state + 128u * j) + dt * mtgp_lds_f32(a_out + 128u * j));
            }
            save_until = tnext;
          }
          const bool live_save = stepping && alive;
          // ---- SaveAt(ts): dense (linear) output for every ts[k] <= tnext ----
          // (once the solve is over every remaining save point is drained, whatever ts holds -- a NaN in ts must not stall the loop)
          while (k < K && (!stepping || ts_k <= save_until)) {
            const float tk = ts_k;
            const float coeff = (tprev == tnext) ? 0.0f : (tk - tprev) / (tnext - tprev);
            float xs[DR];
            if (KIND != MTGP_EVAL_SR) {
#pragma unroll
              for (int i = 0; i < D; i++) xs[i] = live_save ? x[i] + coeff * (x1[i] - x[i]) : mtgp_inff();
            }
            float row = 0.0f;
            for (int j = 0; j < NL; j++) {
              const float a0 = FIXED_SLOTS ? aS[j & 1] : mtgp_lds_f32(a_state + 128u * j);
              const float a1 = FIXED_SLOTS ? aO[j & 1] : mtgp_lds_f32(a_out + 128u * j);
              const float lat = live_save ? a0 + coeff * (a1 - a0) : mtgp_inff();
              if (KIND == MTGP_EVAL_SR) {
                const float e = lat - P.ys_true[((size_t)bb * K + k) * NL + j];
                row = row + e * e;
                if (TRAJ_OUT && P.xs_o && lane_valid && pass == 0) P.xs_o[(((size_t)ind * P.B + b) * K + k) * NL + j] = lat;
              } else {
                mtgp_sts_f32(a_lat + 128u * j, lat);
                if (TRAJ_OUT && P.act_o && lane_valid && pass == 0) P.act_o[(((size_t)ind * P.B + b) * K + k) * NL + j] = lat;
              }
            }
            if (KIND == MTGP_EVAL_SR) {
              sr_total = sr_total + row;
            } else {
              float ys[DR];
#pragma unroll
              for (int i = 0; i < DR; i++) ys[i] = 0.0f;
              // the observation at the save point is needed only when the readout reads it or the caller asked for
              // it; the dynamic policy's readout of the notebooks reads the latent state alone
#ifndef MTGP_SAVE_SKIP_YS
#define MTGP_SAVE_SKIP_YS 1
#endif
              const bool need_ys = !MTGP_SAVE_SKIP_YS || (KIND == MTGP_EVAL_STATIC) || readout_sees_y || (TRAJ_OUT && P.ys_o != nullptr);
              if (need_ys) {
                float nzk[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (NOISE == 1) {
                  const float4 v = __ldg(reinterpret_cast<const float4*>(nz_base + (size_t)(P.cap_steps + k) * nzs));
                  nzk[0] = v.x; nzk[1] = v.y; nzk[2] = v.z; nzk[3] = v.w;
                }
                env.template f_obs<NOISE, false>(tk, xs, ys, n_obs, P.obs_noise, P.partitionable, nzk);
              }
              if (KIND == MTGP_EVAL_STATIC || readout_sees_y) {
#pragma unroll
                for (int i = 0; i < D; i++)
                  if (i < n_obs) mtgp_sts_f32(lbytes + 128u * (uint32_t)(i), ys[i]);
                if (KIND == MTGP_EVAL_DYNAMIC) mtgp_sts_f32(lbytes + 128u * (uint32_t)(slot_u), 0.0f);  // zeros(control_size) (dynamic_evaluate.py:101)
              }
              const float uk = mtgp_run_threaded(prog_save, lbytes);
              const float qk = (ENV == MTGP_ENV_ACROBOT) ? tk / dts : 0.0f;
              if (ENV == MTGP_ENV_ACROBOT && qk > (float)(k + 1)) viol = true;
              env.cost_add(k, qk, xs, uk, fs_known);
              if (TRAJ_OUT && lane_valid && pass == 0 && (P.xs_o || P.ys_o || P.us_o)) {
                const size_t o = ((size_t)ind * P.B + b) * K + k;
                if (P.xs_o)
#pragma unroll
                  for (int i = 0; i < D; i++) P.xs_o[o * D + i] = xs[i];
                if (P.ys_o)
#pragma unroll
                  for (int i = 0; i < D; i++)
                    if (i < n_obs) P.ys_o[o * n_obs + i] = ys[i];
                if (P.us_o) P.us_o[o] = uk;
              }
            }
            k++;
            ts_k = ts[k < K ? k : K - 1];
          }
          if (stepping) {
            // ---- Event(cond_fn): sign change of cond_fn evaluated at (tnext, y1) ----
            bool bad = false;
            if (KIND != MTGP_EVAL_SR) bad = env.bad_state(x1);
            if (FIXED_SLOTS) {
              bad |= nonfinite_cmp(aO[0]) | nonfinite_cmp(aO[1]);
            } else {
              for (int j = 0; j < NL; j++) bad |= !mtgp_isfinitef(mtgp_lds_f32(a_out + 128u * j));
            }
            const bool fired = (bad != ev_old_bad);
            // ---- commit ----  (a lane whose event fires never reads its state again -- its remaining save
            // points are the inf pre-fill -- so it keeps the last finite state and stays on the fast paths)
            const bool adv = alive && !fired;
            if (alive) {
              ev_old_bad = bad;
              lane_steps++;
            }
            if (KIND != MTGP_EVAL_SR) {
#pragma unroll
              for (int i = 0; i < D; i++) x[i] = adv ? x1[i] : x[i];
              if (adv && bad) {  // integrating on in a state that was already bad at t0 (rare)
                x_nonfinite = false;
#pragma unroll
                for (int i = 0; i < D; i++) x_nonfinite |= nonfinite_cmp(x[i]);
              }
            }
            if (FIXED_SLOTS) {
#pragma unroll
              for (int j = 0; j < 2; j++) {
                aS[j] = adv ? aO[j] : aS[j];
                mtgp_sts_f32(a_lat + 128u * j, aS[j]);
              }
            } else {
              for (int j = 0; j < NL; j++) {
                const float cur = mtgp_lds_f32(a_state + 128u * j);
                const float nv = adv ? mtgp_lds_f32(a_out + 128u * j) : cur;
                mtgp_sts_f32(a_state + 128u * j, nv);
                mtgp_sts_f32(a_lat + 128u * j, nv);
              }
            }
            if (alive && fired) alive = false;
            {  // ConstantStepSize: accumulated tnext + dt0, clipped to t1 (_clip_to_end)
              const float new_tprev = tnext < t1 ? tnext : t1;
              float new_tnext = tnext + P.dt0;
              new_tnext = (new_tnext > t1 - 1e-6f) ? t1 : new_tnext;
              tprev = new_tprev;
              tnext = new_tnext;
            }
            steps++;
          }
        }
        if (pass == 0 && lane_valid) my_steps += lane_steps;
        traj_fitness = (KIND == MTGP_EVAL_SR) ? sr_total / (float)K : env.cost_finish(K, fs_known);
        // Acrobot with a ts grid for which the running-prefix cost is not exact (ts[k]/dts > k+1
        // somewhere): redo the solve with first_success known.  Warp-uniform decision.
        if (ENV != MTGP_ENV_ACROBOT || pass == 1) break;
        if (__ballot_sync(0xFFFFFFFFu, viol) == 0u) break;
        fs_known = env.fs < 0 ? 0 : env.fs;
      }
      if (lane_valid) {
        if (TRAJ_OUT && P.tf_o) P.tf_o[(size_t)ind * P.B + b] = traj_fitness;
        fit_sum = fit_sum + (mtgp_isfinitef(traj_fitness) ? traj_fitness : P.max_fitness);
      }
    }
    // mean over B (fixed shuffle tree), clip, parsimony
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fit_sum = fit_sum + __shfl_xor_sync(0xFFFFFFFFu, fit_sum, o);
    unsigned long long warp_steps = my_steps;
    for (int o = 16; o > 0; o >>= 1) warp_steps += __shfl_xor_sync(0xFFFFFFFFu, warp_steps, o);
    if (lane == 0) {
      float f = clip_jnp(fit_sum / (float)P.B, 0.0f, P.max_fitness);
      if (P.live) {
        int cnt = 0;
        for (int t = 0; t < P.T; t++) cnt += P.live[(size_t)ind * P.T + t];
        f = f + P.size_parsimony * (float)cnt;
      }
      P.fitness[ind] = f;
      if (P.steps_done) atomicAdd(P.steps_done, warp_steps);
    }
    __syncwarp();  // the value file and the program area are reused by the next individual
  }  // persistent loop
}

int sm_count() {
  static int n = 0;  // one device model per process
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

size_t rollout_smem(const RolloutParams& rp) {
  return 16 * kWarpsPerCta + (size_t)kWarpsPerCta * warp_area_bytes(rp.th_cap, rp.slots_per_warp);
}

template <int KIND, int ENV, int NOISE, int SPEC, bool TRAJ_OUT, bool UPAR>
int launch_one(RolloutParams rp, int tmp_base, int n_tmp, int th_cap, int want_fast, cudaStream_t stream) {
  rp.slots_per_warp = tmp_base + n_tmp + 1;
  rp.th_cap = (th_cap + 1) & ~1;
  rp.want_fast = want_fast;
  const size_t smem = rollout_smem(rp);
  if (smem > 227 * 1024) {
    mtgp_set_error("rollout needs %zu bytes of shared memory per CTA: over the 227 KB limit", smem);
    return MTGP_ERR_UNSUPPORTED;
  }
  auto kern = k_rollout<KIND, ENV, NOISE, SPEC, TRAJ_OUT, UPAR>;
  MTGP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned blocks = (unsigned)((rp.P + kWarpsPerCta - 1) / kWarpsPerCta);
  rp.next_slot = nullptr;
#ifndef MTGP_NO_PERSISTENT
  if (rp.order && blocks > (unsigned)(sm_count() * MTGP_ROLLOUT_MIN_CTAS)) {  // persistent: as many CTAs as can be resident, each warp loops over schedule slots
    rp.next_slot = const_cast<int*>(rp.order) + rp.P + (want_fast ? 0 : 1);
    const unsigned resident = (unsigned)(sm_count() * MTGP_ROLLOUT_MIN_CTAS);
    blocks = blocks < resident ? blocks : resident;
  }
#endif
  kern<<<blocks, kWarpsPerCta * 32, smem, stream>>>(rp);
  return mtgp_check_launch("k_rollout");
}

// Individuals whose programs need at most kFastTmp temporaries and kFastTh threaded instructions
// (every tree of ordinary depth) run with a small shared-memory footprint and high occupancy; the rest
// (long combs grown by mutation, very large trees) run in a second launch provisioned with the
// worst-case bounds.  Each launch skips the other's individuals.  (kFastTmp, kFastTh: mtgp_env.cuh)

template <int KIND, int ENV>
int launch_kind_env(RolloutParams rp, int tmp_base, int bound, void* noise_ws, bool upar, cudaStream_t stream) {
  int noise = 0;
  if (rp.obs_noise != 0.0f && KIND != MTGP_EVAL_SR) {
    noise = noise_ws ? 1 : 2;
    if (noise == 1) {
      float* tprev = static_cast<float*>(noise_ws);
      int* n_steps = reinterpret_cast<int*>(tprev + rp.cap_steps);
      float* table = tprev + ((rp.cap_steps + 64 + 3) & ~3);  // 16-byte aligned rows
      k_schedule<<<1, 32, 0, stream>>>(rp.ts, rp.K, rp.dt0, rp.max_steps, rp.cap_steps, tprev, n_steps);
      const long long n = (long long)(rp.cap_steps + rp.K) * rp.B;
      k_noise_table<ENV><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rp, tprev, n_steps, table);
      int rc = mtgp_check_launch("k_noise_table");
      if (rc) return rc;
      rp.noise = table;
    }
  }
  const int worst_th = mtgp_worst_threaded(rp.T, rp.lstride);
  const bool all_fast = (bound <= kFastTmp) && (worst_th <= kFastTh);
  rp.fast_tmp = all_fast ? bound : kFastTmp;
  rp.fast_th = all_fast ? worst_th : kFastTh;
  int rc;
  constexpr int ENVD = (KIND == MTGP_EVAL_SR) ? 0 : EnvDim<ENV>::D;
  const bool full_obs = (KIND != MTGP_EVAL_SR) && rp.n_obs == ENVD;
  int spec = !full_obs ? 0 : ((KIND == MTGP_EVAL_DYNAMIC && rp.state_size == 2) ? 2 : 1);
  if (spec == 2) {  // the kernel's compile-time slot map must be the launch's
    constexpr int NT = (KIND == MTGP_EVAL_SR) ? 0 : EnvDim<ENV>::NT;
    if (rp.slot_lat != ENVD || rp.slot_u != ENVD + 2 || rp.slot_target != ENVD + 3 || rp.slot_out != ENVD + 3 + NT ||
        rp.slot_state != ENVD + 5 + NT || rp.n_latent != 2)
      spec = 1;
  }
  const bool traj_out = rp.xs_o || rp.ys_o || rp.us_o || rp.act_o || rp.tf_o;
#define LAUNCH1U(NZ, SPEC, TR, UP)                                                                    \
  {                                                                                                   \
    rc = launch_one<KIND, ENV, NZ, SPEC, TR, UP>(rp, tmp_base, rp.fast_tmp, rp.fast_th, 1, stream);   \
    if (rc == MTGP_OK && !all_fast)                                                                   \
      rc = launch_one<KIND, ENV, NZ, SPEC, TR, UP>(rp, tmp_base, bound, worst_th, 0, stream);         \
  }
  // the uniform-parameter instantiation exists for the Acrobot (the environment whose drift constants matter)
#define LAUNCH1T(NZ, SPEC, TR)                                                                        \
  {                                                                                                   \
    if constexpr (ENV == MTGP_ENV_ACROBOT) { if (upar) LAUNCH1U(NZ, SPEC, TR, true) else LAUNCH1U(NZ, SPEC, TR, false) } \
    else LAUNCH1U(NZ, SPEC, TR, false)                                                                \
  }
#define LAUNCH1(NZ, SPEC) { if (traj_out) LAUNCH1T(NZ, SPEC, true) else LAUNCH1T(NZ, SPEC, false) }
#define LAUNCH2(NZ)                                                                                   \
  if constexpr (KIND == MTGP_EVAL_SR) LAUNCH1(NZ, 0)                                                  \
  else if constexpr (KIND == MTGP_EVAL_STATIC) { if (spec) LAUNCH1(NZ, 1) else LAUNCH1(NZ, 0) }       \
  else { if (spec == 2) LAUNCH1(NZ, 2) else if (spec == 1) LAUNCH1(NZ, 1) else LAUNCH1(NZ, 0) }
  if (noise == 0) { LAUNCH2(0) } else if (noise == 1) { LAUNCH2(1) } else { LAUNCH2(2) }
#undef LAUNCH2
#undef LAUNCH1
#undef LAUNCH1T
#undef LAUNCH1U
  return rc;
}

}  // namespace

// Slot map of the value file (shared with the lowering pass through tmp_base):
//   [0, n_data)                      the data vector the trees read
//   [n_data, n_data + n_latent)      y1 of the latent state
//   [.., + n_latent)                 master copy of the latent state
//   [tmp_base, tmp_base + n_tmp)     temporaries of the programs
int mtgp_rollout_slot_map(const mtgp_rollout_desc& d, int T, int* n_data, int* n_latent) {
  if (d.kind == MTGP_EVAL_SR) {
    *n_data = T;
    *n_latent = T;
  } else if (d.kind == MTGP_EVAL_STATIC) {
    *n_data = d.n_obs + d.n_targets;
    *n_latent = 0;
  } else {
    *n_data = d.n_obs + d.state_size + (T - d.state_size) + d.n_targets;
    *n_latent = d.state_size;
  }
  if (d.solver != MTGP_SOLVER_EULER || d.controller != MTGP_CONTROLLER_CONSTANT) {
    // general Runge-Kutta kernel: 12 arrays of (D + n_latent) components (mtgp_rollout_rk.cu)
    const int D = (d.kind == MTGP_EVAL_SR) ? 0 : d.D;
    return *n_data + 12 * (D + *n_latent);
  }
  return *n_data + 2 * (*n_latent);
}

// Rows of the step part of the noise table, and its size in bytes (0 = generate the noise inline).
static const size_t kNoiseTableMaxBytes = (size_t)512 << 20;
int mtgp_noise_cap_steps(const mtgp_rollout_desc& d) { return d.max_steps < (1 << 22) ? d.max_steps : (1 << 22); }
size_t mtgp_noise_workspace_bytes(const mtgp_rollout_desc& d) {
  if (d.kind == MTGP_EVAL_SR || d.obs_noise == 0.0f) return 0;
  if (d.solver != MTGP_SOLVER_EULER || d.controller != MTGP_CONTROLLER_CONSTANT) return 0;  // per-lane times
  if (d.flags & MTGP_DESC_NOISE_INLINE) return 0;
  if (d.max_steps > mtgp_noise_cap_steps(d)) return 0;  // the table could not hold every step row
  const size_t cap = (size_t)mtgp_noise_cap_steps(d);
  const size_t bytes = ((cap + 64 + 3) & ~(size_t)3) * sizeof(float) + (cap + d.K) * (size_t)4 * d.B * sizeof(float);
  return bytes <= kNoiseTableMaxBytes ? bytes : 0;
}

int mtgp_launch_rollout(const mtgp_programs& progs, const int32_t* live, int P, int T, int N,
                        const mtgp_rollout_desc& d, float size_parsimony, float* fitness,
                        const mtgp_trajectory_out* traj, unsigned long long* steps_done, void* noise_ws,
                        cudaStream_t stream) {
  if (P == 0) return MTGP_OK;
  RolloutParams rp{};
  rp.prog = progs.prog; rp.meta = progs.meta; rp.live = live; rp.lstride = progs.lstride;
  rp.P = P; rp.T = T; rp.kind = d.kind;
  if (progs.order) {
    if (int rc = mtgp_launch_cost_order(progs.meta, P, T, progs.order, stream)) return rc;
    rp.order = progs.order;
  }
  rp.B = d.B; rp.K = d.K; rp.n_obs = d.n_obs; rp.n_targets = d.n_targets; rp.n_params = d.n_params;
  rp.state_size = d.state_size; rp.max_steps = d.max_steps;
  rp.dt0 = d.dt0; rp.max_fitness = d.max_fitness; rp.obs_noise = d.obs_noise; rp.size_parsimony = size_parsimony;
  rp.partitionable = d.threefry_partitionable;
  rp.x0 = d.x0; rp.ts = d.ts; rp.targets = d.targets; rp.params = d.params; rp.ys_true = d.ys_true;
  rp.obs_keys = d.obs_keys;
  rp.fitness = fitness;
  if (traj) {
    rp.xs_o = traj->xs; rp.ys_o = traj->ys; rp.us_o = traj->us; rp.act_o = traj->activities;
    rp.tf_o = traj->traj_fitness;
  }
  rp.steps_done = steps_done;
  rp.cap_steps = mtgp_noise_cap_steps(d);
  if (mtgp_noise_workspace_bytes(d) == 0) noise_ws = nullptr;
  int n_data, n_latent;
  const int reserved = mtgp_rollout_slot_map(d, T, &n_data, &n_latent);
  rp.n_latent = n_latent;
  if (d.kind == MTGP_EVAL_SR) {
    rp.slot_lat = 0; rp.slot_u = 0; rp.slot_target = 0;
  } else if (d.kind == MTGP_EVAL_STATIC) {
    rp.slot_lat = 0; rp.slot_u = 0; rp.slot_target = d.n_obs;
  } else {
    rp.slot_lat = d.n_obs;
    rp.slot_u = d.n_obs + d.state_size;
    rp.slot_target = d.n_obs + d.state_size + (T - d.state_size);
  }
  rp.slot_out = n_data;
  rp.slot_state = n_data + n_latent;
  if (reserved != progs.tmp_base) {
    mtgp_set_error("internal: slot map mismatch (%d vs %d)", reserved, progs.tmp_base);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  const int bound = progs.n_tmp;
  if (d.solver != MTGP_SOLVER_EULER || d.controller != MTGP_CONTROLLER_CONSTANT) {
    // Dopri5 and/or PIDController: the general kernel (same two provisioning classes)
    rp.solver = d.solver;
    rp.controller = d.controller;
    rp.rtol = d.rtol; rp.atol = d.atol; rp.dtmin = d.dtmin;
    rp.safety = d.pid_safety > 0.0f ? d.pid_safety : 0.9f;
    rp.factormin = d.pid_factormin > 0.0f ? d.pid_factormin : 0.2f;
    rp.factormax = d.pid_factormax > 0.0f ? d.pid_factormax : 10.0f;
    const bool dflt = (d.pcoeff == 0.0f && d.icoeff == 0.0f && d.dcoeff == 0.0f);
    const float pc = dflt ? 0.0f : d.pcoeff, ic = dflt ? 1.0f : d.icoeff, dc = dflt ? 0.0f : d.dcoeff;
    const float error_order = 5.0f;  // Dopri5.order
    rp.coeff1 = (ic + pc + dc) / error_order;
    rp.coeff2 = -(pc + 2.0f * dc) / error_order;
    rp.coeff3 = dc / error_order;
    rp.slot_rk = n_data;
    rp.slots_per_warp = progs.tmp_base + bound + 1;
    rp.th_cap = mtgp_worst_threaded(T, progs.lstride);
    return mtgp_launch_rollout_rk(rp, d.kind, d.env_id, progs.tmp_base, bound, stream);
  }
  // trajectory-uniform environment parameters, visible on the host: derived constants into the kernel parameters
  bool upar = false;
  if (d.env_id == MTGP_ENV_ACROBOT && d.kind != MTGP_EVAL_SR && d.params_host && d.n_params == 4) {
    upar = true;
    for (int b = 1; b < d.B && upar; b++)
      for (int i = 0; i < 4; i++) upar &= (mtgp_f2u(d.params_host[b * 4 + i]) == mtgp_f2u(d.params_host[i]));
    if (upar) acrobot_derive(d.params_host[0], d.params_host[1], d.params_host[2], d.params_host[3], rp.envc);
  }
#define DISPATCH(KIND, ENV) return launch_kind_env<KIND, ENV>(rp, progs.tmp_base, bound, noise_ws, upar, stream)
#ifdef MTGP_DEV_ONLY_DYNAMIC_ACROBOT  // tuning builds (tools/build_variant.sh): one evaluator kind and environment, a fraction of the compile time
  if (d.kind == MTGP_EVAL_DYNAMIC && d.env_id == MTGP_ENV_ACROBOT) DISPATCH(MTGP_EVAL_DYNAMIC, MTGP_ENV_ACROBOT);
  mtgp_set_error("tuning build: dynamic-policy Acrobot only");
  return MTGP_ERR_UNSUPPORTED;
#else
  if (d.kind == MTGP_EVAL_SR) DISPATCH(MTGP_EVAL_SR, MTGP_ENV_NONE);
  if (d.kind == MTGP_EVAL_STATIC) {
    if (d.env_id == MTGP_ENV_ACROBOT) DISPATCH(MTGP_EVAL_STATIC, MTGP_ENV_ACROBOT);
    if (d.env_id == MTGP_ENV_HARMONIC) DISPATCH(MTGP_EVAL_STATIC, MTGP_ENV_HARMONIC);
    if (d.env_id == MTGP_ENV_REACTOR) DISPATCH(MTGP_EVAL_STATIC, MTGP_ENV_REACTOR);
  }
  if (d.kind == MTGP_EVAL_DYNAMIC) {
    if (d.env_id == MTGP_ENV_ACROBOT) DISPATCH(MTGP_EVAL_DYNAMIC, MTGP_ENV_ACROBOT);
    if (d.env_id == MTGP_ENV_HARMONIC) DISPATCH(MTGP_EVAL_DYNAMIC, MTGP_ENV_HARMONIC);
    if (d.env_id == MTGP_ENV_REACTOR) DISPATCH(MTGP_EVAL_DYNAMIC, MTGP_ENV_REACTOR);
  }
  mtgp_set_error("unknown evaluator kind %d / environment %d", d.kind, d.env_id);
  return MTGP_ERR_INVALID_ARGUMENT;
#endif
#undef DISPATCH
}
