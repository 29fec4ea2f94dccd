// mtgp_rollout_rk.cu -- K2b: the fused rollout for the solver settings the reference's notebooks
// actually pass: diffrax.Dopri5() with PIDController(atol, rtol, dtmin) (examples/StaticPolicy.ipynb,
// examples/DynamicPolicy.ipynb, examples/SymbolicRegression.ipynb), and Dopri5 with a constant step.
// (Euler + ConstantStepSize, the evaluators' class default, has its own tuned kernel in mtgp_rollout.cu.)
//
// Same mapping as K2 -- one warp per individual, lanes = trajectories, warp-uniform programs -- but
// every lane now carries its OWN time grid: step sizes, accept/reject decisions, save indices and
// the observation-noise keys (fold_in(key, bits(t)) at the stage times) are per lane.  The warp loops
// until its slowest lane is done, exactly what the reference's vmapped while_loop does.  All solver
// state (Y, stage increments, candidate) lives in the per-warp shared-memory value file.
//
// Semantics restated from diffrax (THIRD-PARTY, identical in oracle/mtgp_oracle.c: diffeqsolve):
// ERK stepping with FSAL, b_error, PIDController.adapt_step_size with rms_norm, force_dtmin,
// _clip_to_end, 4th-order dense output of _Dopri5Interpolation for SaveAt(ts), Event after accepted steps.
// one register set for the interpreter's instruction fetch (mtgp_threaded.cuh): this kernel's hot loop is several
// times the instruction cache, and half the handler code is worth 5 % here (C3: 72.5 -> 68.6 ms)
#ifndef MTGP_TH_SETS
#define MTGP_TH_SETS 1
#endif
#include "mtgp_env.cuh"

using namespace mtgp_dev;

namespace {

__constant__ float DP_A[6][6] = {
    {(float)(1.0 / 5), 0, 0, 0, 0, 0},
    {(float)(3.0 / 40), (float)(9.0 / 40), 0, 0, 0, 0},
    {(float)(44.0 / 45), (float)(-56.0 / 15), (float)(32.0 / 9), 0, 0, 0},
    {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), 0, 0},
    {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176), (float)(-5103.0 / 18656), 0},
    {(float)(35.0 / 384), 0.0f, (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784), (float)(11.0 / 84)}};
__constant__ float DP_C[6] = {(float)(1.0 / 5), (float)(3.0 / 10), (float)(4.0 / 5), (float)(8.0 / 9), 1.0f, 1.0f};
__constant__ float DP_BERR[7] = {(float)(35.0 / 384 - 1951.0 / 21600), 0.0f, (float)(500.0 / 1113 - 22642.0 / 50085),
                                 (float)(125.0 / 192 - 451.0 / 720), (float)(-2187.0 / 6784 - -12231.0 / 42400),
                                 (float)(11.0 / 84 - 649.0 / 6300), (float)(-1.0 / 60)};
__constant__ float DP_CMID[7] = {(float)(6025192743.0 / 30085553152.0 / 2), 0.0f,
                                 (float)(51252292925.0 / 65400821598.0 / 2), (float)(-2691868925.0 / 45128329728.0 / 2),
                                 (float)(187940372067.0 / 1594534317056.0 / 2), (float)(-1776094331.0 / 19743644256.0 / 2),
                                 (float)(11237099.0 / 235043384.0 / 2)};

constexpr int kRkWarps = 1;  // one warp per CTA: the adaptive workloads are small populations, fine-grained CTAs balance the SMs
// 12 solver arrays per lane: Y, Y1, YI, F0, KS[0..6], SV

// SPEC as in k_rollout: 0 generic; 2 = every state component observed and the latent count known (static: 0,
// dynamic: 2), so that D + n_latent, the solver-array offsets and the component loops are compile-time.
template <int KIND, int ENV, bool NOISE, int SPEC>
// the solver's component loops stay rolled: the kernel's hot loop is far larger than the instruction cache,
// and code size costs more than loop overhead here
#ifdef MTGP_RK_UNROLL
#define MTGP_RK_LOOP
#else
#define MTGP_RK_LOOP _Pragma("unroll 1")
#endif
#ifndef MTGP_RK_MIN_CTAS
#define MTGP_RK_MIN_CTAS 16  // 122 registers, no spills; shared memory allows 15 resident CTAs for the fast class
#endif
__global__ void __launch_bounds__(kRkWarps * 32, MTGP_RK_MIN_CTAS) k_rollout_rk(const RolloutParams P) {
  constexpr int D = (KIND == MTGP_EVAL_SR) ? 0 : EnvDim<ENV>::D;
  constexpr int DR = D > 0 ? D : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int slot0 = blockIdx.x * kRkWarps + wic;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  const WarpArea wa = warp_area(P, smem_raw + 16 * kRkWarps + (size_t)wic * warp_area_bytes(P.th_cap, P.slots_per_warp));
  // persistent warps, as in k_rollout: the first individual comes with the launch, later ones from the counter
  uint64_t* bar = bars + 2 * wic;
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  uint32_t bar_phase = 0;
  bool first_round = true;
  for (;;) {
    int slot = slot0;
    if (!first_round) {
      if (!P.next_slot) return;
      if (lane == 0) slot = (int)(gridDim.x * kRkWarps) + atomicAdd(P.next_slot, 1);
      slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
    }
    first_round = false;
    if (slot >= P.P) return;
    const int ind = P.order ? P.order[slot] : slot;
    // which launch owns this individual (as in k_rollout): small programs run with a small footprint
    if (individual_is_fast(P, ind) != (P.want_fast != 0)) continue;

    // ---- stage the programs (same as K2) and expand every tree to its own threaded program (lane t: tree t) ----
    warp_stage_programs(P, ind, lane, bar, bar_phase, wa);
    float* lb = wa.vf + lane;
    const mtgp_saddr lbytes = mtgp_saddr_of(lb);
    // Dynamic policy: when the readout reads neither an observation nor the control slot (the notebooks'
    // variable lists: it reads the latent state alone), one interpreter call per vector-field evaluation runs
    // the whole chain [readout -> slot_u][state eq j -> slot_f + j] as in k_rollout (three calls otherwise:
    // the readout has to see zeros there, dynamic_evaluate.py:113).  slot_f: the latent part of the SV array,
    // dead outside the save-point loop.
    bool chain = false;
    if (KIND == MTGP_EVAL_DYNAMIC) {
      const int n_obs_rt = P.n_obs;
      const uint32_t soff = compact_offset(P.meta + (size_t)ind * P.T, P.state_size);
      const int n = (int)MTGP_META_N(P.meta[(size_t)ind * P.T + P.state_size]);
      bool reads = false;
      for (int i = lane; i < n; i += 32) {
        const uint2 c = wa.prog_s[soff + i];
        const uint32_t xk = (c.x >> 4) & 3u, yk = (c.x >> 6) & 3u, xs = (c.x >> 16) & 0xFFu, ys = c.x >> 24;
        reads |= (xk == OK_MEM && ((int)xs < n_obs_rt || (int)xs == P.slot_u)) ||
                 (yk == OK_MEM && ((int)ys < n_obs_rt || (int)ys == P.slot_u));
      }
      chain = __any_sync(0xFFFFFFFFu, reads) == 0;
    }
    if (chain) {
      if (lane == 0) {
        const uint32_t* mi = P.meta + (size_t)ind * P.T;
        const int Dc = (KIND == MTGP_EVAL_SR) ? 0 : EnvDim<ENV>::D;
        const int slot_f = P.slot_rk + 11 * (Dc + P.n_latent) + Dc;
        const int off = th_build_chain(mi, wa.prog_s, wa.th, P.state_size, 1, -1, -1, -1);  // stand-alone readout (save points)
        wa.woff[P.state_size] = 0;
        wa.woff[0] = off;
        th_build_chain(mi, wa.prog_s, wa.th + off, 0, P.state_size, slot_f, P.state_size, P.slot_u);
      }
    } else {
      for (int t = lane; t < P.T; t += 32) {
        int off = 0;
        for (int q = 0; q < t; q++) off += (int)MTGP_META_TH(P.meta[(size_t)ind * P.T + q]);
        wa.woff[t] = off;
        th_build_chain(P.meta + (size_t)ind * P.T, wa.prog_s, wa.th + off, t, 1, -1, -1, -1);
      }
    }
    __syncwarp();

    const float* ts = P.ts;
    constexpr bool FIXED = (SPEC == 2) && (KIND != MTGP_EVAL_SR);
    const int n_obs = FIXED ? D : P.n_obs;
    const int K = P.K, NL = FIXED ? (KIND == MTGP_EVAL_STATIC ? 0 : 2) : P.n_latent, DT = D + NL;
    const float t0 = ts[0], t1 = ts[K - 1];
    const float dts = (K > 1) ? ts[1] - ts[0] : 0.0f;
    const bool dopri = P.solver == MTGP_SOLVER_DOPRI5, pid = P.controller == MTGP_CONTROLLER_PID;
    const mtgp_saddr th_s = mtgp_saddr_of(wa.th);
    const mtgp_saddr prog_readout = th_s + 16u * (uint32_t)wa.woff[(KIND == MTGP_EVAL_DYNAMIC) ? NL : 0];
    const mtgp_saddr prog_chain = th_s + 16u * (uint32_t)wa.woff[0];  // (chain only)
    // solver arrays in the value file: component i of array a is slot slot_rk + a*DT + i
    float* const Y = lb + (size_t)(P.slot_rk + 0 * DT) * 32;
    float* const Y1 = lb + (size_t)(P.slot_rk + 1 * DT) * 32;
    float* const YI = lb + (size_t)(P.slot_rk + 2 * DT) * 32;
    float* const F0 = lb + (size_t)(P.slot_rk + 3 * DT) * 32;
    float* const KS = lb + (size_t)(P.slot_rk + 4 * DT) * 32;  // KS[st][i] at (st*DT + i)*32
    float* const SV = lb + (size_t)(P.slot_rk + 11 * DT) * 32; // state at a save point
    float fit_sum = 0.0f;
    unsigned int my_steps = 0;

    for (int b0 = 0; b0 < P.B; b0 += 32) {
      const int b = b0 + lane;
      const bool lane_valid = b < P.B;
      const int bb = lane_valid ? b : P.B - 1;
      Env<ENV> env;
      if (KIND != MTGP_EVAL_SR) env.init(P, bb);
      if (KIND != MTGP_EVAL_SR)
        for (int i = 0; i < P.n_targets; i++) lb[(P.slot_target + i) * 32] = P.targets[bb * P.n_targets + i];

      // vector field of the evaluator at (t, src): writes dst[i] = scale * f_i  (i < DT)
      auto vector_field = [&](float t, const float* src, float* dst, float scale, float* raw) {
        float x[DR], y[DR], dx[DR];
        for (int j = 0; j < NL; j++) lb[(P.slot_lat + j) * 32] = src[(D + j) * 32];
        if (KIND != MTGP_EVAL_SR) {
#pragma unroll
          for (int i = 0; i < D; i++) x[i] = src[i * 32];
          env.template f_obs<NOISE ? 2 : 0>(t, x, y, n_obs, P.obs_noise, P.partitionable, nullptr);
          if (KIND == MTGP_EVAL_STATIC || chain) {
#pragma unroll
            for (int i = 0; i < D; i++)
              if (i < n_obs) lb[i * 32] = y[i];
          } else {
#pragma unroll
            for (int i = 0; i < D; i++)
              if (i < n_obs) lb[i * 32] = 0.0f;
            lb[P.slot_u * 32] = 0.0f;
          }
          float u;
          if (chain) {  // readout and state equations in one call; the readout's ST leaves u in its slot
            mtgp_run_threaded(prog_chain, lbytes);
            u = lb[P.slot_u * 32];
          } else {
            u = mtgp_run_threaded(prog_readout, lbytes);
          }
          env.template drift<false>(x, u, dx, P);
#pragma unroll
          for (int i = 0; i < D; i++) {
            dst[i * 32] = scale * dx[i];
            if (raw) raw[i * 32] = dx[i];
          }
          if (KIND == MTGP_EVAL_DYNAMIC && !chain) {
#pragma unroll
            for (int i = 0; i < D; i++)
              if (i < n_obs) lb[i * 32] = y[i];
            lb[P.slot_u * 32] = u;
          }
        }
        for (int j = 0; j < NL; j++) {
          const float f = chain ? SV[(D + j) * 32] : mtgp_run_threaded(th_s + 16u * (uint32_t)wa.woff[j], lbytes);
          dst[(D + j) * 32] = scale * f;
          if (raw) raw[(D + j) * 32] = f;
        }
      };
      auto bad_state = [&](const float* src) {
        bool bad = false;
        if (KIND != MTGP_EVAL_SR) {
          float x[DR];
#pragma unroll
          for (int i = 0; i < D; i++) x[i] = src[i * 32];
          bad = env.bad_state(x);
        }
        for (int j = 0; j < NL; j++) bad |= !mtgp_isfinitef(src[(D + j) * 32]);
        return bad;
      };

      int fs_known = -1;
      bool viol = false;
      float traj_fitness = 0.0f;
      for (int pass = 0; pass < 2; pass++) {
        if (KIND != MTGP_EVAL_SR) {
#pragma unroll
          for (int i = 0; i < D; i++) Y[i * 32] = P.x0[bb * D + i];
        }
        for (int j = 0; j < NL; j++) Y[(D + j) * 32] = (KIND == MTGP_EVAL_SR) ? P.x0[bb * NL + j] : 0.0f;
        if (KIND != MTGP_EVAL_SR) { env.sum = 0.0f; env.sum0 = 0.0f; env.fs = -1; }
        float sr_total = 0.0f;
        bool ev_old_bad = bad_state(Y);
        bool active = true, fsal_valid = false, at_dtmin = false;  // fsal_valid: F0 holds f(tprev, Y)
        float tprev = t0, tnext = t0 + P.dt0;
        tnext = tnext < t1 ? tnext : t1;
        float prev_inv = 1.0f, prev_prev_inv = 1.0f;
        int num_steps = 0, k = 0;

        // one save point of the post-pass for the lanes with `doit`; xs/lat hold the saved state
        auto save_point = [&](bool doit, const float* sv /* DT components at sv[i*32] */) {
          const int kk = k < K ? k : K - 1;
          const float tk = ts[kk];
          if (KIND == MTGP_EVAL_SR) {
            float row = 0.0f;
            for (int j = 0; j < NL; j++) {
              const float lat = sv[j * 32];
              const float e = lat - P.ys_true[((size_t)bb * K + kk) * NL + j];
              row = row + e * e;
              if (doit && P.xs_o && lane_valid && pass == 0) P.xs_o[(((size_t)ind * P.B + b) * K + kk) * NL + j] = lat;
            }
            if (doit) sr_total = sr_total + row;
          } else {
            float xs[DR], ys[DR];
#pragma unroll
            for (int i = 0; i < D; i++) xs[i] = sv[i * 32];
            for (int j = 0; j < NL; j++) {
              const float lat = sv[(D + j) * 32];
              lb[(P.slot_lat + j) * 32] = lat;
              if (doit && P.act_o && lane_valid && pass == 0) P.act_o[(((size_t)ind * P.B + b) * K + kk) * NL + j] = lat;
            }
            env.template f_obs<NOISE ? 2 : 0>(tk, xs, ys, n_obs, P.obs_noise, P.partitionable, nullptr);
#pragma unroll
            for (int i = 0; i < D; i++)
              if (i < n_obs) lb[i * 32] = ys[i];
            if (KIND == MTGP_EVAL_DYNAMIC) lb[P.slot_u * 32] = 0.0f;
            const float uk = mtgp_run_threaded(prog_readout, lbytes);
            const float qk = (ENV == MTGP_ENV_ACROBOT) ? tk / dts : 0.0f;
            if (doit) {
              if (ENV == MTGP_ENV_ACROBOT && qk > (float)(kk + 1)) viol = true;
              env.cost_add(kk, qk, xs, uk, fs_known);
              if (lane_valid && pass == 0 && (P.xs_o || P.ys_o || P.us_o)) {
                const size_t o = ((size_t)ind * P.B + b) * K + kk;
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
          }
        };

        // One iteration = one attempted step of every lane that is still integrating, followed by the
        // save points that step covers (per lane).  When no lane integrates any more, the remaining save
        // points of every lane are drained from the inf pre-fill of SaveAt -- through the same code.
        while (true) {
          const bool stepping = __ballot_sync(0xFFFFFFFFu, active && lane_valid) != 0u;
          if (!stepping && __ballot_sync(0xFFFFFFFFu, k < K) == 0u) break;
          const float dt = tnext - tprev;
          bool keep = true, nan_err = false;
          float new_tprev = tprev, new_tnext = tnext;
          if (stepping) {
            // ---- stages (Euler: one; Dopri5: seven, the first one reused under FSAL) ----
            const int n_stages = dopri ? 7 : 1;
            const bool need_f0 = !dopri || (__ballot_sync(0xFFFFFFFFu, !fsal_valid) != 0u);
            for (int st = 0; st < n_stages; st++) {
              const float* src = Y;
              float tst = tprev;
              if (st > 0) {
                MTGP_RK_LOOP
                for (int i = 0; i < DT; i++) {
                  float acc = 0.0f;
                  MTGP_RK_LOOP
                  for (int j = 0; j < st; j++) {
                    const float term = DP_A[st - 1][j] * KS[(j * DT + i) * 32];
                    acc = (j == 0) ? term : acc + term;
                  }
                  YI[i * 32] = Y[i * 32] + acc;
                }
                if (st == 6)
                  for (int i = 0; i < DT; i++) Y1[i * 32] = YI[i * 32];
                const float cst = DP_C[st - 1];
                tst = (cst == 1.0f) ? tnext : tprev + cst * dt;
                src = YI;
              }
              if (st == 0 && !need_f0) {
                for (int i = 0; i < DT; i++) KS[i * 32] = dt * F0[i * 32];  // FSAL / same point after a rejection
              } else {
                // unscaled copies: stage 0 -> F0; the last stage overwrites YI (dead by then) and becomes
                // the next step's F0 when this step is accepted
                float* raw = dopri ? (st == 0 ? F0 : (st == 6 ? YI : nullptr)) : nullptr;
                vector_field(tst, src, KS + (size_t)st * DT * 32, dt, raw);
              }
            }
            fsal_valid = true;
            if (!dopri)
              for (int i = 0; i < DT; i++) Y1[i * 32] = Y[i * 32] + KS[i * 32];
            // ---- step-size controller ----
            if (!pid) {
              new_tprev = tnext;
              new_tnext = tnext + P.dt0;
            } else {
              float sumsq = 0.0f;
              MTGP_RK_LOOP
              for (int i = 0; i < DT; i++) {
                float yerr = 0.0f;
                for (int j = 0; j < 7; j++) {
                  const float term = DP_BERR[j] * KS[(j * DT + i) * 32];
                  yerr = (j == 0) ? term : yerr + term;
                }
                const float ym = mtgp_maximumf(mtgp_absf(Y[i * 32]), mtgp_absf(Y1[i * 32]));
                const float sc = yerr / (P.atol + ym * P.rtol);
                sumsq = sumsq + sc * sc;
              }
              const float scaled_error = mtgp_sqrtf(sumsq / (float)DT);
              keep = (scaled_error < 1.0f) || at_dtmin;
              const float inv_err = (scaled_error == 0.0f) ? mtgp_inff() : 1.0f / scaled_error;
              const float factor1 = (P.coeff1 == 0.0f) ? 1.0f : mtgp_expf(P.coeff1 * mtgp_logf(inv_err));
              const float factor2 = (P.coeff2 == 0.0f) ? 1.0f : mtgp_expf(P.coeff2 * mtgp_logf(prev_inv));
              const float factor3 = (P.coeff3 == 0.0f) ? 1.0f : mtgp_expf(P.coeff3 * mtgp_logf(prev_prev_inv));
              const float fmin = keep ? 1.0f : P.factormin;
              float factor = ((P.safety * factor1) * factor2) * factor3;
              factor = clip_jnp(factor, fmin, P.factormax);
              float ndt = dt * factor;
              if (P.dtmin > 0.0f) {
                at_dtmin = (ndt <= P.dtmin);
                ndt = (ndt != ndt) ? ndt : (ndt > P.dtmin ? ndt : P.dtmin);
              }
              new_tprev = keep ? tnext : tprev;
              new_tnext = new_tprev + ndt;
              if (keep && active) { prev_prev_inv = prev_inv; prev_inv = inv_err; }
              nan_err = (scaled_error != scaled_error);
            }
            new_tprev = new_tprev < t1 ? new_tprev : t1;
            {
              const bool clip = new_tnext > t1 - 1e-6f;
              const float tclip = keep ? t1 : new_tprev + 0.5f * (t1 - new_tprev);
              new_tnext = clip ? tclip : new_tnext;
            }
          }
          const bool commit = stepping && active && keep && !nan_err;
          // ---- SaveAt(ts): every lane has its own pending save points ----
          while (true) {
            const int kk = k < K ? k : K - 1;
            const float tk = ts[kk];
            const bool pending = (k < K) && (stepping ? (commit && tk <= tnext) : true);
            if (__ballot_sync(0xFFFFFFFFu, pending) == 0u) break;
            const float th = (tprev == tnext) ? 0.0f : (tk - tprev) / (tnext - tprev);
            MTGP_RK_LOOP
            for (int i = 0; i < DT; i++) {
              const float y0 = Y[i * 32], y1 = Y1[i * 32];
              float v;
              if (!dopri) {
                v = y0 + th * (y1 - y0);
              } else {
                float acc = 0.0f;
                for (int j = 0; j < 7; j++) {
                  const float term = DP_CMID[j] * KS[(j * DT + i) * 32];
                  acc = (j == 0) ? term : acc + term;
                }
                const float ymid = y0 + acc;
                const float f0 = KS[i * 32], f1 = KS[(6 * DT + i) * 32];
                const float a = ((2.0f * (f1 - f0)) - (8.0f * (y1 + y0))) + (16.0f * ymid);
                const float bq = ((((5.0f * f0) - (3.0f * f1)) + (18.0f * y0)) + (14.0f * y1)) - (32.0f * ymid);
                const float cq = ((((f1 - (4.0f * f0)) - (11.0f * y0)) - (5.0f * y1))) + (16.0f * ymid);
                v = ((((a * th + bq) * th + cq) * th + f0) * th) + y0;
              }
              SV[i * 32] = stepping ? v : mtgp_inff();  // drained save points keep the inf pre-fill
            }
            save_point(pending, SV);
            if (pending) k++;
          }
          if (stepping) {
            // ---- Event + commit ----
            const bool bad = bad_state(Y1);
            const bool fired = (bad != ev_old_bad);
            if (commit) {
              ev_old_bad = bad;
              my_steps += (pass == 0 && lane_valid) ? 1u : 0u;
              for (int i = 0; i < DT; i++) {
                Y[i * 32] = Y1[i * 32];
                if (dopri) F0[i * 32] = YI[i * 32];  // unscaled f at (tnext, y1)
              }
            }
            if (active) {
              num_steps++;
              tprev = new_tprev;
              tnext = new_tnext;
              if ((commit && fired) || nan_err || !(tprev < t1) || num_steps >= P.max_steps) active = false;
            }
          }
        }
        traj_fitness = (KIND == MTGP_EVAL_SR) ? sr_total / (float)K : env.cost_finish(K, fs_known);
        if (ENV != MTGP_ENV_ACROBOT || pass == 1) break;
        if (__ballot_sync(0xFFFFFFFFu, viol) == 0u) break;
        fs_known = env.fs < 0 ? 0 : env.fs;
      }
      if (lane_valid) {
        if (P.tf_o) P.tf_o[(size_t)ind * P.B + b] = traj_fitness;
        fit_sum = fit_sum + (mtgp_isfinitef(traj_fitness) ? traj_fitness : P.max_fitness);
      }
    }
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

// Two provisioning classes as in the fixed-step kernel (kFastTmp, kFastTh: mtgp_env.cuh); each launch skips the
// other's individuals.

template <int KIND, int ENV>
int launch_rk_class(RolloutParams rp, int tmp_base, int n_tmp, int th_cap, int want_fast, cudaStream_t stream) {
  rp.slots_per_warp = tmp_base + n_tmp + 1;
  rp.th_cap = (th_cap + 1) & ~1;
  rp.want_fast = want_fast;
  const size_t smem = 16 * kRkWarps + (size_t)kRkWarps * warp_area_bytes(rp.th_cap, rp.slots_per_warp);
  if (smem > 227 * 1024) {
    mtgp_set_error("adaptive rollout needs %zu bytes of shared memory per CTA: over the 227 KB limit", smem);
    return MTGP_ERR_UNSUPPORTED;
  }
  unsigned blocks = (unsigned)((rp.P + kRkWarps - 1) / kRkWarps);
  rp.next_slot = nullptr;
  {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    const size_t per_cta = smem + 1024;  // + the per-CTA reservation
    size_t fit = ((size_t)228 * 1024) / per_cta;
    if (fit > MTGP_RK_MIN_CTAS) fit = MTGP_RK_MIN_CTAS;
    if (fit < 1) fit = 1;
    const unsigned resident = (unsigned)(n_sm * fit);
    if (rp.order && blocks > resident) {  // persistent warps (see k_rollout)
      rp.next_slot = const_cast<int*>(rp.order) + rp.P + (want_fast ? 0 : 1);
      blocks = resident;
    }
  }
  constexpr int ENVD = (KIND == MTGP_EVAL_SR) ? 0 : EnvDim<ENV>::D;
  const bool fixed = (KIND != MTGP_EVAL_SR) && rp.n_obs == ENVD &&
                     (KIND == MTGP_EVAL_STATIC ? rp.n_latent == 0 : rp.n_latent == 2);
  const bool noise = rp.obs_noise != 0.0f && KIND != MTGP_EVAL_SR;
#define RK_LAUNCH(NZ, SPEC)                                                                             \
  {                                                                                                     \
    auto kern = k_rollout_rk<KIND, ENV, NZ, SPEC>;                                                      \
    MTGP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<blocks, kRkWarps * 32, smem, stream>>>(rp);                                                  \
  }
  if constexpr (KIND == MTGP_EVAL_SR) {
    RK_LAUNCH(false, 0)
  } else if (fixed) {
    if (noise) RK_LAUNCH(true, 2) else RK_LAUNCH(false, 2)
  } else {
    if (noise) RK_LAUNCH(true, 0) else RK_LAUNCH(false, 0)
  }
#undef RK_LAUNCH
  return mtgp_check_launch("k_rollout_rk");
}

template <int KIND, int ENV>
int launch_rk(RolloutParams rp, int tmp_base, int bound, cudaStream_t stream) {
  // The adaptive kernel's value file is large (12 solver arrays), shared memory decides how many warps an SM holds,
  // and the kernel is latency-bound: its fast class is provisioned for the programs of ordinary notebook trees
  // (kRkFastTh threaded instructions: 16 resident warps per SM on the DynamicPolicy shapes instead of 14).
#ifndef MTGP_RK_FAST_TH
#define MTGP_RK_FAST_TH 64
#endif
#ifndef MTGP_RK_FAST_TMP
#define MTGP_RK_FAST_TMP kFastTmp
#endif
  constexpr int kRkFastTh = MTGP_RK_FAST_TH, kRkFastTmp = MTGP_RK_FAST_TMP;
  const int worst_th = mtgp_worst_threaded(rp.T, rp.lstride);
  const bool all_fast = (bound <= kRkFastTmp) && (worst_th <= kRkFastTh);
  rp.fast_tmp = all_fast ? bound : kRkFastTmp;
  rp.fast_th = all_fast ? worst_th : kRkFastTh;
  int rc = launch_rk_class<KIND, ENV>(rp, tmp_base, rp.fast_tmp, rp.fast_th, 1, stream);
  if (rc == MTGP_OK && !all_fast) rc = launch_rk_class<KIND, ENV>(rp, tmp_base, bound, worst_th, 0, stream);
  return rc;
}

}  // namespace

int mtgp_launch_rollout_rk(RolloutParams rp, int kind, int env_id, int tmp_base, int bound, cudaStream_t stream) {
#define DISPATCH(KIND, ENV) return launch_rk<KIND, ENV>(rp, tmp_base, bound, stream)
  if (kind == MTGP_EVAL_SR) DISPATCH(MTGP_EVAL_SR, MTGP_ENV_NONE);
  if (kind == MTGP_EVAL_STATIC) {
    if (env_id == MTGP_ENV_ACROBOT) DISPATCH(MTGP_EVAL_STATIC, MTGP_ENV_ACROBOT);
    if (env_id == MTGP_ENV_HARMONIC) DISPATCH(MTGP_EVAL_STATIC, MTGP_ENV_HARMONIC);
    if (env_id == MTGP_ENV_REACTOR) DISPATCH(MTGP_EVAL_STATIC, MTGP_ENV_REACTOR);
  }
  if (kind == MTGP_EVAL_DYNAMIC) {
    if (env_id == MTGP_ENV_ACROBOT) DISPATCH(MTGP_EVAL_DYNAMIC, MTGP_ENV_ACROBOT);
    if (env_id == MTGP_ENV_HARMONIC) DISPATCH(MTGP_EVAL_DYNAMIC, MTGP_ENV_HARMONIC);
    if (env_id == MTGP_ENV_REACTOR) DISPATCH(MTGP_EVAL_DYNAMIC, MTGP_ENV_REACTOR);
  }
#undef DISPATCH
  mtgp_set_error("unknown evaluator kind %d / environment %d", kind, env_id);
  return MTGP_ERR_INVALID_ARGUMENT;
}
