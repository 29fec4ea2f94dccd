// mtgp_select.cu -- what the strategy does with the gathered fitness vector, on the device:
//   best = argmin(fitness); best_solution = flat_populations[best]     (genetic_programming.py:426-431)
//   elite = argsort(fitness_of_sub_population)[:elite_size]             (reproduction.py:82)
// so that a generation needs no host round trip between evaluation and reproduction.
//
// Ordering rules are numpy's / jax.numpy's: argmin returns the FIRST index of the minimum and a NaN
// counts as the minimum; argsort is stable, -0.0 == +0.0, and NaNs sort last.
#include "mtgp_internal.cuh"

namespace {

constexpr int kSelThreads = 1024;

// total order for argsort: ascending float, NaN last, -0 == +0; ties by index (stable)
__device__ __forceinline__ unsigned long long sort_key(float f, int idx) {
  uint32_t u;
  if (f != f) {
    u = 0xFFFFFFFFu;
  } else {
    u = __float_as_uint(f + 0.0f);  // -0 -> +0
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  return ((unsigned long long)u << 32) | (uint32_t)idx;
}

// total order for argmin: NaN first, then ascending; ties by index
__device__ __forceinline__ unsigned long long min_key(float f, int idx) {
  uint32_t u;
  if (f != f) {
    u = 0u;
  } else {
    u = __float_as_uint(f + 0.0f);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  return ((unsigned long long)u << 32) | (uint32_t)idx;
}

__device__ __forceinline__ unsigned long long block_min(unsigned long long v, unsigned long long* sm) {
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
    v = w < v ? w : v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // sm may still be read from the previous round
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  unsigned long long r = sm[lane < (int)(blockDim.x >> 5) ? lane : 0];
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, r, o);
    r = w < r ? w : r;
  }
  return r;
}

// one CTA per sub-population: elite_size rounds of "smallest key above the previous one"
__global__ void __launch_bounds__(kSelThreads) k_elite(const float* __restrict__ fitness, int population_size,
                                                       int elite_size, int32_t* __restrict__ elite_out) {
  __shared__ unsigned long long sm[32];
  const float* f = fitness + (size_t)blockIdx.x * population_size;
  unsigned long long prev = 0;
  bool have_prev = false;
  for (int e = 0; e < elite_size; e++) {
    unsigned long long best = ~0ull;
    for (int i = threadIdx.x; i < population_size; i += blockDim.x) {
      const unsigned long long k = sort_key(f[i], i);
      if ((!have_prev || k > prev) && k < best) best = k;
    }
    best = block_min(best, sm);
    if (threadIdx.x == 0) elite_out[(size_t)blockIdx.x * elite_size + e] = (int32_t)(uint32_t)best;
    prev = best;
    have_prev = true;
  }
}

// one CTA: argmin over the flattened population, then the copy of the winning candidate
__global__ void __launch_bounds__(kSelThreads) k_best(const float* __restrict__ fitness, int P,
                                                      int32_t* __restrict__ best_index, float* __restrict__ best_fitness,
                                                      const float* __restrict__ population, int row_floats,
                                                      float* __restrict__ best_solution) {
  __shared__ unsigned long long sm[32];
  unsigned long long best = ~0ull;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const unsigned long long k = min_key(fitness[i], i);
    best = k < best ? k : best;
  }
  best = block_min(best, sm);
  const int idx = (int)(uint32_t)best;
  if (threadIdx.x == 0) {
    if (best_index) *best_index = idx;
    if (best_fitness) *best_fitness = fitness[idx];
  }
  if (population && best_solution) {
    const float* src = population + (size_t)idx * row_floats;
    for (int i = threadIdx.x; i < row_floats; i += blockDim.x) best_solution[i] = src[i];
  }
}

}  // namespace

extern "C" int mtgp_select(const float* fitness, int num_populations, int population_size, int elite_size,
                           int32_t* best_index, float* best_fitness, int32_t* elite,
                           const float* population, int row_floats, float* best_solution, void* stream) {
  if (num_populations < 1 || population_size < 1) {
    mtgp_set_error("mtgp_select: need num_populations >= 1 and population_size >= 1 (got %d, %d)", num_populations,
                   population_size);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if ((long long)num_populations * population_size > 0x7FFFFFFFll) {
    mtgp_set_error("mtgp_select: flattened population does not fit an int32 index");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (elite_size < 0 || elite_size > population_size) {
    mtgp_set_error("mtgp_select: elite_size must be in [0, population_size] (got %d)", elite_size);
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if (!fitness || (elite_size > 0 && !elite)) {
    mtgp_set_error("mtgp_select: fitness (and elite when elite_size > 0) must not be NULL");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  if ((population != nullptr) != (best_solution != nullptr) || (population && row_floats < 1)) {
    mtgp_set_error("mtgp_select: population, row_floats and best_solution go together");
    return MTGP_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (elite_size > 0) {
    k_elite<<<num_populations, kSelThreads, 0, st>>>(fitness, population_size, elite_size, elite);
    if (int rc = mtgp_check_launch("k_elite")) return rc;
  }
  if (best_index || best_fitness || best_solution) {
    k_best<<<1, kSelThreads, 0, st>>>(fitness, num_populations * population_size, best_index, best_fitness, population,
                                      row_floats, best_solution);
    if (int rc = mtgp_check_launch("k_best")) return rc;
  }
  return MTGP_OK;
}
