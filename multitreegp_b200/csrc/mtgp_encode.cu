// mtgp_encode.cu -- K0: population encoder / decoder and the lowering pass.
//
//   k_encode : f32[P,T,N,4] tree arrays (reference format, SURVEY.md App. A;
//              initialization.py:96-98, genetic_programming.py:368-374)  ->  level-1 SoA bytecode
//   k_decode : the inverse (bit-exact unless MTGP_BC_LOSSY)
//   k_lower  : level-1 bytecode -> accumulator-machine programs (mtgp_program.cuh)
//
// HBM-bound, coalesced: k_encode reads 16 B and writes 9 B per row with one warp per tree.
#include "mtgp_internal.cuh"
#include "mtgp_lower.cuh"

namespace {

__device__ __forceinline__ int f2i_trunc(float v) {
  // .astype(int) (XLA convert f32->s32): truncation; NaN -> 0, saturating
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 2147483647;
  if (v <= -2147483648.0f) return -2147483647 - 1;
  return (int)v;
}

// One warp per tree.
__global__ void __launch_bounds__(256) k_encode(const float* __restrict__ nodes, int ns,
                                               const float* __restrict__ coef, int cs, long long n_trees,
                                               int N, int F, mtgp_bytecode bc) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= n_trees) return;
  const long long row0 = warp * N;
  int live = 0;
  unsigned lossy = 0;
  for (int i = lane; i < N; i += 32) {
    const long long r = row0 + i;
    float f, a, b, c;
    if (ns == 4 && cs == 4 && coef == nodes + 3) {  // combined [.,4] array: one 16-byte load per row
      const float4 v = __ldg(reinterpret_cast<const float4*>(nodes) + r);
      f = v.x; a = v.y; b = v.z; c = v.w;
    } else {
      f = __ldg(nodes + r * ns);
      a = __ldg(nodes + r * ns + 1);
      b = __ldg(nodes + r * ns + 2);
      c = __ldg(coef + r * cs);
    }
    live += (f != 0.0f);  // jnp.sum(array[:,:,0] != 0), genetic_programming.py:424
    int fi = f2i_trunc(f);
    int fc = min(max(fi, 0), F - 1);  // lax.switch clamps its index
    if ((float)fi != f || fc != fi) lossy = 1;
    // f_idx == 1 selects the coefficient only for exactly 1.0 (genetic_programming.py:372);
    // 1 < f < 2 truncates to node function 1, which returns 0.0 -> encode as the empty node
    if (fc == 1 && f != 1.0f) fc = 0;
    int ai = f2i_trunc(a), bi = f2i_trunc(b);
    int ac = min(max(ai, -32768), 32767), bcl = min(max(bi, -32768), 32767);
    if ((float)ai != a || ac != ai || (float)bi != b || bcl != bi) lossy = 1;
    bc.op[r] = (uint8_t)fc;
    bc.a[r] = (int16_t)ac;
    bc.b[r] = (int16_t)bcl;
    bc.coef[r] = c;
  }
  live = __reduce_add_sync(0xFFFFFFFFu, live);
  lossy = __reduce_or_sync(0xFFFFFFFFu, lossy);
  if (lane == 0) {
    bc.live[warp] = live;
    bc.flags[warp] = lossy ? MTGP_BC_LOSSY : 0;
  }
}

__global__ void __launch_bounds__(256) k_decode(mtgp_bytecode bc, long long n_rows, float4* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  out[r] = make_float4((float)bc.op[r], (float)bc.a[r], (float)bc.b[r], bc.coef[r]);
}

// One thread per tree.  NMAX bounds the per-thread scratch (local memory).
template <int NMAX>
__global__ void __launch_bounds__(128) k_lower(mtgp_bytecode bc, long long n_trees, int N,
                                              const mtgp_node_table tab, int tmp_base, int lstride,
                                              uint2* __restrict__ prog, uint32_t* __restrict__ meta) {
  const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= n_trees) return;
  bool canonical;
  meta[tree] = mtgp_lower_tree<NMAX>(bc.op + tree * N, bc.a + tree * N, bc.b + tree * N, bc.coef + tree * N, N, tab,
                                     tmp_base, prog + tree * lstride, &canonical, mtgp_tmp_bound(N));
  bc.flags[tree] = (uint8_t)((bc.flags[tree] & MTGP_BC_LOSSY) | (canonical ? MTGP_BC_CANONICAL : 0));
}

}  // namespace

// ---- host launchers ---------------------------------------------------------------------------
int mtgp_launch_encode(const float* nodes, int ns, const float* coef, int cs, int P, int T, int N,
                       const mtgp_node_table* table, mtgp_bytecode out, cudaStream_t stream) {
  const long long n_trees = (long long)P * T;
  if (n_trees == 0) return MTGP_OK;
  const int warps_per_block = 8;
  const long long blocks = (n_trees + warps_per_block - 1) / warps_per_block;
  k_encode<<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(nodes, ns, coef, cs, n_trees, N, table->n_ids, out);
  return mtgp_check_launch("k_encode");
}

int mtgp_launch_decode(mtgp_bytecode in, int P, int T, int N, float* trees_out, cudaStream_t stream) {
  const long long n_rows = (long long)P * T * N;
  if (n_rows == 0) return MTGP_OK;
  const long long blocks = (n_rows + 255) / 256;
  k_decode<<<(unsigned)blocks, 256, 0, stream>>>(in, n_rows, reinterpret_cast<float4*>(trees_out));
  return mtgp_check_launch("k_decode");
}

int mtgp_launch_lower(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table, int tmp_base,
                      int lstride, uint2* prog, uint32_t* meta, cudaStream_t stream) {
  const long long n_trees = (long long)P * T;
  if (n_trees == 0) return MTGP_OK;
  const unsigned blocks = (unsigned)((n_trees + 127) / 128);
#define LOWER(NMAX) k_lower<NMAX><<<blocks, 128, 0, stream>>>(bc, n_trees, N, *table, tmp_base, lstride, prog, meta)
  if (N <= 32) LOWER(32);
  else if (N <= 64) LOWER(64);
  else if (N <= 128) LOWER(128);
  else if (N <= 256) LOWER(256);
  else LOWER(MTGP_MAX_NODES);
#undef LOWER
  return mtgp_check_launch("k_lower");
}

// ---- cost-ordered schedule ------------------------------------------------------------------------------
// A warp's (rollout) or CTA's (pointwise) run time grows with the instructions of its individuals'
// programs: a 60-instruction individual takes several times as long as a 5-instruction one.  CTAs are
// dispatched in blockIdx order, so handing out the individuals costliest-first (longest-processing-
// time-first) keeps the last wave short; it also makes neighbours alike, which the pointwise kernel's
// groups of 8 and its launch classes profit from.  One CTA, counting sort on the total instruction
// count: order[slot] = individual.
namespace {
constexpr int kCostBins = 256;
__global__ void __launch_bounds__(1024) k_cost_order(const uint32_t* __restrict__ meta, int P, int T,
                                                     int32_t* __restrict__ order) {
  __shared__ int bin_start[kCostBins];
  for (int i = threadIdx.x; i < kCostBins; i += blockDim.x) bin_start[i] = 0;
  __syncthreads();
  auto cost_bin = [&](int ind) {
    int c = 0;
    for (int t = 0; t < T; t++) c += (int)MTGP_META_N(meta[(size_t)ind * T + t]) + 1;
    return kCostBins - 1 - min(c, kCostBins - 1);  // bin 0 = costliest
  };
  for (int i = threadIdx.x; i < P; i += blockDim.x) atomicAdd(&bin_start[cost_bin(i)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < kCostBins; b++) {
      const int n = bin_start[b];
      bin_start[b] = run;
      run += n;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) order[atomicAdd(&bin_start[cost_bin(i)], 1)] = i;
  if (threadIdx.x < 2) order[P + threadIdx.x] = 0;  // slot counters of the persistent rollout launches
}
}  // namespace

int mtgp_launch_cost_order(const uint32_t* meta, int P, int T, int32_t* order, cudaStream_t stream) {
  if (P == 0) return MTGP_OK;
  k_cost_order<<<1, 1024, 0, stream>>>(meta, P, T, order);
  return mtgp_check_launch("k_cost_order");
}
