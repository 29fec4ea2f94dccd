// mtgp_internal.cuh -- declarations shared by the translation units of libmtgp_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mtgp_threaded.cuh"

// ---- error reporting (thread-local message behind mtgp_last_error) ----------------------------
void mtgp_set_error(const char* fmt, ...);
int mtgp_check_launch(const char* what);  // cudaGetLastError -> MTGP_ERR_CUDA
#define MTGP_CUDA_TRY(expr)                                                       \
  do {                                                                            \
    cudaError_t e_ = (expr);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      mtgp_set_error("%s failed: %s", #expr, cudaGetErrorString(e_));             \
      return MTGP_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

// ---- launchers ------------------------------------------------------------------------------
int mtgp_launch_encode(const float* nodes, int ns, const float* coef, int cs, int P, int T, int N,
                       const mtgp_node_table* table, mtgp_bytecode out, cudaStream_t stream);
int mtgp_launch_decode(mtgp_bytecode in, int P, int T, int N, float* trees_out, cudaStream_t stream);
int mtgp_launch_lower(mtgp_bytecode bc, int P, int T, int N, const mtgp_node_table* table, int tmp_base,
                      int lstride, uint2* prog, uint32_t* meta, cudaStream_t stream);

// order[slot] = individual, costliest (most instructions) first; see mtgp_encode.cu
int mtgp_launch_cost_order(const uint32_t* meta, int P, int T, int32_t* order, cudaStream_t stream);

struct mtgp_programs {
  const uint2* prog;     // [P*T, lstride]
  const uint32_t* meta;  // [P*T]  n_instr | n_tmp_slots << 16
  int lstride;           // instructions reserved per tree (even)
  int tmp_base;          // first temporary slot
  int n_tmp;             // temporary slots provisioned per warp
  int32_t* order;        // [P + 2] scratch: cost-ordered schedule + two zeroed slot counters (may be NULL)
};

int mtgp_launch_rollout(const mtgp_programs& progs, const int32_t* live, int P, int T, int N,
                        const mtgp_rollout_desc& desc, float size_parsimony, float* fitness,
                        const mtgp_trajectory_out* traj, unsigned long long* steps_done, void* noise_ws,
                        cudaStream_t stream);
size_t mtgp_noise_workspace_bytes(const mtgp_rollout_desc& d);  // 0: noise generated inline

int mtgp_launch_pointwise(const mtgp_programs& progs, const int32_t* live, int P, int T, int N,
                          const float* Xt, const float* Yt, int n_points, int n_points_padded,
                          int n_vars, float max_fitness, float size_parsimony, float* fitness,
                          float* pred, cudaStream_t stream);
int mtgp_launch_transpose_points(const float* X, const float* Y, int n_points, int n_points_padded,
                                 int n_vars, int T, float* Xt, float* Yt, cudaStream_t stream);

// slot bookkeeping shared by the host code
static __host__ __device__ inline int mtgp_tmp_bound(int N) { return (N + 1) / 2 + 1; }  // ceil(N/2)+1
static inline int mtgp_lstride(int N) { return (N + 2 + 1) & ~1; }
