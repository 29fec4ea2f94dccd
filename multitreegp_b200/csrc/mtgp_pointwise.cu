// mtgp_pointwise.cu -- K1: pointwise tree evaluation over a data set (symbolic-regression scale-up).
//
// Oracle: pred = vmap(vmap_foriloop, in_axes=[None, 0])(candidate, X)   (genetic_programming.py:390-401)
//         fitness = mean_n sum_t (pred - Y)^2, where(nan|inf, max_fitness), clip   (SR_evaluator.py:24,42-45)
//
// Mapping: a CTA owns G individuals and ALL data points; its warps stride over 32*R-point chunks.
// For each chunk a warp loads the variables once (coalesced float4 from the transposed data set,
// which is L2-resident) into its shared-memory value file and then runs the programs of all G
// individuals on it, so every data byte fetched from L2 is used G times.  Each lane evaluates R
// points per dispatched instruction, which amortises the (warp-uniform) dispatch over 32*R points.
// R = 8 (the two regular launch classes): programs are staged by cp.async.bulk + mbarrier, expanded once per
// CTA to threaded code and run by the eight-values-per-lane interpreter of mtgp_threaded_r8.cuh.
// R = 1 (worst-case provisioning for trees of extreme size): the compact-format interpreter.
#include "mtgp_env.cuh"  // mbarrier / bulk-copy helpers
#include "mtgp_threaded_r8.cuh"

namespace {

#ifndef MTGP_PW_R
#define MTGP_PW_R 8  // points per lane per dispatched instruction in the fast pointwise launch
#endif
#ifndef MTGP_PW_MIN_CTAS
#define MTGP_PW_MIN_CTAS 5  // resident CTAs per SM the register allocation is tuned for (shared memory allows 5 in the fast class)
#endif
constexpr int kPwWarps = 4;
#ifndef MTGP_PW_GROUP
#define MTGP_PW_GROUP 8  // individuals sharing one CTA (and every data chunk a warp loads)
#endif
constexpr int kPwGroup = MTGP_PW_GROUP;

struct PointwiseParams {
  const uint2* prog;
  const uint32_t* meta;
  const int32_t* live;
  const int32_t* order;  // schedule slot -> individual (cost-ordered), NULL = identity
  int lstride;
  int P, T;
  const float* Xt;  // [n_vars, npad]
  const float* Yt;  // [T, npad]
  int n_points, npad, n_vars;
  float max_fitness, size_parsimony;
  float* fitness;
  float* pred;
  int tmp_base;    // = n_vars + T
  // this launch handles the individuals whose programs fit (cap_tmp temporaries, cap_instr instructions
  // incl. END per tree) but do not fit the previous, smaller class (prev_tmp, prev_instr; -1 = none)
  int cap_tmp, cap_instr, prev_tmp, prev_instr;
  int pstride;     // instructions reserved per tree in shared memory (compact and threaded form alike)
  int slots_per_warp;  // data + temporaries (+ 1 scratch slot for R = 8)
};


template <int R>
__global__ void __launch_bounds__(kPwWarps * 32, MTGP_PW_MIN_CTAS) k_pointwise(const PointwiseParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_need[kPwGroup];
  __shared__ int s_ind[kPwGroup];  // the individuals of this CTA's group
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int ind0 = blockIdx.x * kPwGroup;
  const int G = min(kPwGroup, P.P - ind0);
  const int words_per_ind = P.T * P.pstride;
  uint2* prog_s = reinterpret_cast<uint2*>(smem_raw);
  float* acc_s = reinterpret_cast<float*>(smem_raw + (size_t)kPwGroup * words_per_ind * (sizeof(uint2) + (R == 8 ? sizeof(mtgp_tinstr) : 0)));
  float* vf = acc_s + kPwGroup * kPwWarps * 32 + (size_t)wic * P.slots_per_warp * 32 * R;

  // which individuals of the group belong to this launch (by temporaries and program length)
  if (threadIdx.x < kPwGroup) {
    int mine = 0;
    if ((int)threadIdx.x < G) {
      const int ind = P.order ? P.order[ind0 + threadIdx.x] : ind0 + (int)threadIdx.x;
      s_ind[threadIdx.x] = ind;
      int need = 0, len = 0;
      for (int t = 0; t < P.T; t++) {
        const uint32_t m = P.meta[(size_t)ind * P.T + t];
        need = max(need, (int)MTGP_META_TMP(m));
        len = max(len, (int)MTGP_META_TH(m));  // threaded instructions incl. header and END (>= compact length + 1)
      }
      const bool fits = (need <= P.cap_tmp) && (len <= P.cap_instr);
      const bool fits_prev = (need <= P.prev_tmp) && (len <= P.prev_instr);
      mine = (fits && !fits_prev) ? 1 : 0;
    }
    s_need[threadIdx.x] = mine;
  }
  __syncthreads();
  bool any = false;
  for (int g = 0; g < G; g++) any |= (s_need[g] != 0);
  if (!any) return;

  // stage the programs of the group -- only the instructions each tree actually has -- with cp.async.bulk
  // completing on one mbarrier (16-byte granules: program rows are 16-byte aligned, lstride is even)
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x == 0) {
    mtgp_dev::mbar_init(&s_bar, 1);
    uint32_t total = 0;
    for (int gt = 0; gt < G * P.T; gt++)
      if (s_need[gt / P.T]) total += ((MTGP_META_N(P.meta[(size_t)s_ind[gt / P.T] * P.T + gt % P.T]) + 2u) & ~1u) * 8u;
    mtgp_dev::mbar_expect_tx(&s_bar, total);
    for (int gt = 0; gt < G * P.T; gt++) {
      if (!s_need[gt / P.T]) continue;
      const size_t tree = (size_t)s_ind[gt / P.T] * P.T + gt % P.T;
      mtgp_dev::bulk_g2s(prog_s + (size_t)gt * P.pstride, P.prog + tree * P.lstride,
                         ((MTGP_META_N(P.meta[tree]) + 2u) & ~1u) * 8u, &s_bar);
    }
  }
  __syncthreads();
  mtgp_dev::mbar_wait(&s_bar, 0);
  mtgp_tinstr* th_s = reinterpret_cast<mtgp_tinstr*>(prog_s + (size_t)kPwGroup * words_per_ind);
  if constexpr (R == 8) {
    // expand to threaded code (slot stride 1024 B, no in-block trig): one thread per tree
    for (int gt = threadIdx.x; gt < G * P.T; gt += blockDim.x) {
      if (!s_need[gt / P.T]) continue;
      const size_t tree = (size_t)s_ind[gt / P.T] * P.T + gt % P.T;
      const int n = (int)MTGP_META_N(P.meta[tree]) + 1;  // with X_END
      const uint2* src = prog_s + (size_t)gt * P.pstride;
      mtgp_tinstr* dst = th_s + (size_t)gt * P.pstride;
      int m = 1;
      for (int i = 0; i < n; i++) m += mtgp_th_expand<1024, false, true>(src[i], dst + m);
      mtgp_th_link(dst, m - 1);
    }
    __syncthreads();
  }

  float* lb = vf + lane * (R % 4 == 0 ? 4 : R);
  constexpr int chunk_pts = 32 * R;
  auto point_of = [&](int c, int r) { return c * chunk_pts + (R % 4 == 0 ? (r / 4) * 128 + lane * 4 + (r % 4) : lane * R + r); };
  const int n_chunks = P.npad / chunk_pts;
  // per-lane partial sums of squared error, one per individual of the group, in shared memory
  for (int g = 0; g < kPwGroup; g++) acc_s[(g * kPwWarps + wic) * 32 + lane] = 0.0f;

  for (int c = wic; c < n_chunks; c += kPwWarps) {
    const bool ragged = (c + 1) * chunk_pts > P.n_points;  // warp-uniform: only the last chunk has padding
    // variables and targets of this chunk -> value file slots [0, n_vars) and [n_vars, n_vars+T)
    for (int v = 0; v < P.n_vars + P.T; v++) {
      const float* src = (v < P.n_vars) ? P.Xt + (size_t)v * P.npad : P.Yt + (size_t)(v - P.n_vars) * P.npad;
      Vec<R> d;
      if constexpr (R % 4 == 0) {
#pragma unroll
        for (int i = 0; i < R / 4; i++) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(src + c * chunk_pts + i * 128 + lane * 4));
          d.v[4 * i] = t4.x; d.v[4 * i + 1] = t4.y; d.v[4 * i + 2] = t4.z; d.v[4 * i + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < R; i++) d.v[i] = __ldg(src + point_of(c, i));
      }
      vf_store<R>(lb, v, d);
    }
    for (int g = 0; g < G; g++) {
      if (!s_need[g]) continue;
      float row[R];
#pragma unroll
      for (int r = 0; r < R; r++) row[r] = 0.0f;
      for (int t = 0; t < P.T; t++) {
        Vec<R> pr;
        if constexpr (R == 8)
          pr = mtgp_run_threaded8(mtgp_saddr_of(th_s + (size_t)g * words_per_ind + (size_t)t * P.pstride), mtgp_saddr_of(lb),
                                  (uint32_t)(P.slots_per_warp - 1) * 1024u);
        else
          pr = mtgp_interpret<R>(prog_s + (size_t)g * words_per_ind + (size_t)t * P.pstride, lb);
        const Vec<R> yt = vf_load<R>(lb, P.n_vars + t);
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float e = pr.v[r] - yt.v[r];
          row[r] = row[r] + e * e;
        }
        if (P.pred) {
#pragma unroll
          for (int r = 0; r < R; r++)
            if (point_of(c, r) < P.n_points) P.pred[((size_t)s_ind[g] * P.n_points + point_of(c, r)) * P.T + t] = pr.v[r];
        }
      }
      float s = 0.0f;
      if (ragged) {
#pragma unroll
        for (int r = 0; r < R; r++) s = s + ((point_of(c, r) < P.n_points) ? row[r] : 0.0f);
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) s = s + row[r];
      }
      acc_s[(g * kPwWarps + wic) * 32 + lane] += s;
    }
  }
  __syncthreads();
  for (int g = wic; g < G; g += kPwWarps) {
    if (!s_need[g]) continue;
    float s = 0.0f;
    for (int w = 0; w < kPwWarps; w++) s = s + acc_s[(g * kPwWarps + w) * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = s + __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0) {
      float f = s / (float)P.n_points;
      if (!mtgp_isfinitef(f)) f = P.max_fitness;
      f = f < 0.0f ? 0.0f : (f > P.max_fitness ? P.max_fitness : f);
      if (P.live) {
        int cnt = 0;
        for (int t = 0; t < P.T; t++) cnt += P.live[(size_t)s_ind[g] * P.T + t];
        f = f + P.size_parsimony * (float)cnt;
      }
      P.fitness[s_ind[g]] = f;
    }
  }
}

// X [n, n_vars] -> Xt [n_vars, npad], Y [n, T] -> Yt [T, npad]; the pad is zero-filled
__global__ void k_transpose_points(const float* __restrict__ X, const float* __restrict__ Y, int n, int npad,
                                   int n_vars, int T, float* __restrict__ Xt, float* __restrict__ Yt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  for (int v = 0; v < n_vars; v++) Xt[(size_t)v * npad + i] = (i < n) ? X[(size_t)i * n_vars + v] : 0.0f;
  for (int t = 0; t < T; t++) Yt[(size_t)t * npad + i] = (i < n) ? Y[(size_t)i * T + t] : 0.0f;
}

template <int R>
int launch_pw(PointwiseParams pp, int n_tmp, int pstride, int prev_tmp, int prev_instr, cudaStream_t stream) {
  pp.cap_tmp = n_tmp;
  pp.cap_instr = pstride;
  pp.prev_tmp = prev_tmp;
  pp.prev_instr = prev_instr;
  pp.pstride = pstride;
  pp.slots_per_warp = pp.tmp_base + n_tmp + (R == 8 ? 1 : 0);  // R = 8: one scratch slot for the slow operators
  const size_t smem = (size_t)kPwGroup * pp.T * pstride * (sizeof(uint2) + (R == 8 ? sizeof(mtgp_tinstr) : 0)) +
                      (size_t)kPwGroup * kPwWarps * 32 * 4 + (size_t)kPwWarps * pp.slots_per_warp * 32 * R * 4;
  if (smem > 227 * 1024) {
    mtgp_set_error("pointwise kernel needs %zu bytes of shared memory per CTA: over the 227 KB limit", smem);
    return MTGP_ERR_UNSUPPORTED;
  }
  auto kern = k_pointwise<R>;
  MTGP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned blocks = (unsigned)((pp.P + kPwGroup - 1) / kPwGroup);
  kern<<<blocks, kPwWarps * 32, smem, stream>>>(pp);
  return mtgp_check_launch("k_pointwise");
}

}  // namespace

int mtgp_launch_transpose_points(const float* X, const float* Y, int n_points, int n_points_padded, int n_vars,
                                 int T, float* Xt, float* Yt, cudaStream_t stream) {
  k_transpose_points<<<(n_points_padded + 255) / 256, 256, 0, stream>>>(X, Y, n_points, n_points_padded, n_vars, T,
                                                                       Xt, Yt);
  return mtgp_check_launch("k_transpose_points");
}

int mtgp_launch_pointwise(const mtgp_programs& progs, const int32_t* live, int P, int T, int N, const float* Xt,
                          const float* Yt, int n_points, int n_points_padded, int n_vars, float max_fitness,
                          float size_parsimony, float* fitness, float* pred, cudaStream_t stream) {
  if (P == 0) return MTGP_OK;
  PointwiseParams pp{};
  pp.prog = progs.prog; pp.meta = progs.meta; pp.live = live; pp.lstride = progs.lstride;
  pp.P = P; pp.T = T;
  pp.Xt = Xt; pp.Yt = Yt; pp.n_points = n_points; pp.npad = n_points_padded; pp.n_vars = n_vars;
  pp.max_fitness = max_fitness; pp.size_parsimony = size_parsimony;
  pp.fitness = fitness; pp.pred = pred;
  pp.tmp_base = progs.tmp_base;
  if (progs.order) {  // groups of similar cost, costliest first
    if (int rc = mtgp_launch_cost_order(progs.meta, P, T, progs.order, stream)) return rc;
    pp.order = progs.order;
  }
  // Three provisioning classes, smallest first; each launch skips the individuals of the others.
  //   fast   : <= 4 temporaries, <= 32 instructions per tree (every tree of ordinary size and depth):
  //            R points per lane, small shared-memory footprint, 6-8 resident CTAs per SM
  //   medium : <= 8 temporaries, <= 128 instructions: still R points per lane, fewer resident CTAs
  //   worst  : anything a max_nodes-row array can hold (long combs grown by mutation): one point per lane
  // (lengths are threaded-code instructions per tree, header and END included: MTGP_META_TH)
  const int bound = mtgp_tmp_bound(N), worst_instr = MTGP_TH_MAX_EXPAND * progs.lstride + 2;
#ifndef MTGP_PW_FAST_TMP
#define MTGP_PW_FAST_TMP 4
#endif
#ifndef MTGP_PW_FAST_TH
#define MTGP_PW_FAST_TH 40
#endif
  const int t0 = bound < MTGP_PW_FAST_TMP ? bound : MTGP_PW_FAST_TMP, i0 = worst_instr < MTGP_PW_FAST_TH ? worst_instr : MTGP_PW_FAST_TH;
  int rc = launch_pw<MTGP_PW_R>(pp, t0, i0, -1, -1, stream);
  if (rc != MTGP_OK || (t0 == bound && i0 == worst_instr)) return rc;
  const int t1 = bound < 8 ? bound : 8, i1 = worst_instr < 160 ? worst_instr : 160;
  rc = launch_pw<MTGP_PW_R>(pp, t1, i1, t0, i0, stream);
  if (rc != MTGP_OK || (t1 == bound && i1 == worst_instr)) return rc;
  rc = launch_pw<1>(pp, bound, worst_instr, t1, i1, stream);
  return rc;
}
