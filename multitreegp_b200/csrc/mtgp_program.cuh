// mtgp_program.cuh -- level-2 ("executable") program format and the warp-uniform interpreter.
//
// A level-1 bytecode tree (mtgp_bytecode, include/mtgp.h) is lowered by k_lower (mtgp_encode.cu)
// into a short program for an ACCUMULATOR MACHINE:
//
//   * one instruction per OPERATOR node of the tree -- leaves never cost a dispatch: a variable
//     leaf becomes a memory operand (M) that names a slot of the per-warp value file in shared
//     memory, a coefficient leaf becomes a 32-bit immediate (I) inside the instruction;
//   * the result of the previous instruction lives in a register (A); in postfix order the left
//     operand of an operator is, whenever it is itself an operator, exactly that register;
//   * a result that is needed later (the right operand of an operator whose both children are
//     operators) is stored to a temporary slot by the instruction that produced it;
//   * sub-trees without variables are folded into immediates by the lowering pass with the SAME
//     device functions the interpreter uses, so folding is bit-exact.
//
// The reference evaluates every one of the max_nodes rows through lax.switch over all node
// functions (genetic_programming.py:356-388); this executes only live operator nodes, with the
// same float32 result for every tree.
//
// Instruction = 8 bytes:  .x = handler | store_slot << 8 | x_slot << 16 | y_slot << 24
//                         .y = immediate (float bits)
// Slot s of a warp's value file is the 32*R floats at  base + s*32*R ; lane l owns elements
// [l*R, l*R+R) -- conflict-free for R = 1 (4-byte) and R = 4 (16-byte) accesses.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mtgp.h"
#include "../../include/mtgp_math.h"

#define MTGP_NO_SLOT 0xFFu
#define MTGP_MAX_SLOTS 254

enum mtgp_handler : uint32_t {
  H_END = 0,
  H_LOADI = 1,  // acc = imm
  H_LOADM = 2,  // acc = M[x]
  // unary, operand in A (9) then operand in M (9); order follows mtgp_opcode SIN..SQUARE
  H_UA_BASE = 3,
  H_UM_BASE = 12,
  // commutative binary: ADD, MUL, MAX, MIN x {AM, AI, MM, MI}
  H_ADD_BASE = 21,
  H_MUL_BASE = 25,
  H_MAX_BASE = 29,
  H_MIN_BASE = 33,
  // non-commutative: SUB, DIV x {AM, AI, MA, IA, MM, MI, IM}
  H_SUB_BASE = 37,
  H_DIV_BASE = 44,
  H_COUNT = 51
};
// kind offsets inside a binary group
enum { K_AM = 0, K_AI = 1, K_MM = 2, K_MI = 3 };                                       // commutative
enum { N_AM = 0, N_AI = 1, N_MA = 2, N_IA = 3, N_MM = 4, N_MI = 5, N_IM = 6 };           // non-commutative

__host__ __device__ __forceinline__ uint2 mtgp_make_instr(uint32_t h, uint32_t st, uint32_t xs,
                                                          uint32_t ys, float imm) {
  uint2 w;
  w.x = (h & 0xFFu) | ((st & 0xFFu) << 8) | ((xs & 0xFFu) << 16) | ((ys & 0xFFu) << 24);
  w.y = mtgp_f2u(imm);
  return w;
}

// ---- node functions (one definition, used by the interpreter AND by constant folding) --------
__host__ __device__ __forceinline__ float mtgp_apply_unary(int opcode, float x) {
  switch (opcode) {
    case MTGP_OP_SIN: return mtgp_sinf(x);
    case MTGP_OP_COS: return mtgp_cosf(x);
    case MTGP_OP_EXP: return mtgp_expf(x);
    case MTGP_OP_LOG: return mtgp_logf(x);
    case MTGP_OP_SQRT: return mtgp_sqrtf(x);
    case MTGP_OP_TANH: return mtgp_tanhf(x);
    case MTGP_OP_NEG: return -x;
    case MTGP_OP_ABS: return mtgp_absf(x);
    default: return x * x;  // MTGP_OP_SQUARE
  }
}
__host__ __device__ __forceinline__ float mtgp_apply_binary(int opcode, float x, float y) {
  switch (opcode) {
    case MTGP_OP_ADD: return x + y;
    case MTGP_OP_SUB: return x - y;
    case MTGP_OP_MUL: return x * y;
    case MTGP_OP_DIV: return x / y;
    case MTGP_OP_MAX: return mtgp_maximumf(x, y);
    default: return mtgp_minimumf(x, y);
  }
}

// ---- value-file access ----------------------------------------------------------------------
template <int R>
struct Vec {
  float v[R];
};

template <int R>
__host__ __device__ __forceinline__ Vec<R> vf_load(const float* lane_base, uint32_t slot) {
  Vec<R> r;
  const float* p = lane_base + slot * (32 * R);
  if constexpr (R % 4 == 0) {
#pragma unroll
    for (int i = 0; i < R / 4; i++) {
      float4 t = reinterpret_cast<const float4*>(p)[i];
      r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < R; i++) r.v[i] = p[i];
  }
  return r;
}
template <int R>
__host__ __device__ __forceinline__ void vf_store(float* lane_base, uint32_t slot, const Vec<R>& a) {
  float* p = lane_base + slot * (32 * R);
  if constexpr (R % 4 == 0) {
#pragma unroll
    for (int i = 0; i < R / 4; i++)
      reinterpret_cast<float4*>(p)[i] = make_float4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < R; i++) p[i] = a.v[i];
  }
}

#define MTGP_FOR_R _Pragma("unroll") for (int r = 0; r < R; r++)

// ---- the interpreter ------------------------------------------------------------------------
// prog: shared-memory program (terminated by H_END), identical for all lanes of the warp, so
// every branch below is warp-uniform.  lane_base = value file + lane*R.
// Returns the value of the tree's root for the R points / trajectories this lane owns.
template <int R>
__host__ __device__ __noinline__ Vec<R> mtgp_interpret(const uint2* __restrict__ prog, float* lane_base) {
  Vec<R> acc;
  MTGP_FOR_R acc.v[r] = 0.0f;
  uint2 w = prog[0];
  int pc = 0;
  while (true) {
    const uint32_t h = w.x & 0xFFu;
    if (h == H_END) break;
    const uint32_t st = (w.x >> 8) & 0xFFu, xs = (w.x >> 16) & 0xFFu, ys = w.x >> 24;
    const float imm = mtgp_u2f(w.y);
    w = prog[++pc];  // prefetch the next instruction under this one's arithmetic
    switch (h) {
      case H_LOADI: MTGP_FOR_R acc.v[r] = imm; break;
      case H_LOADM: acc = vf_load<R>(lane_base, xs); break;
#define U_CASES(OFF, OPC)                                                                        \
  case H_UA_BASE + OFF: MTGP_FOR_R acc.v[r] = mtgp_apply_unary(OPC, acc.v[r]); break;            \
  case H_UM_BASE + OFF: {                                                                        \
    Vec<R> x = vf_load<R>(lane_base, xs);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_unary(OPC, x.v[r]);                                         \
  } break;
      U_CASES(0, MTGP_OP_SIN)
      U_CASES(1, MTGP_OP_COS)
      U_CASES(2, MTGP_OP_EXP)
      U_CASES(3, MTGP_OP_LOG)
      U_CASES(4, MTGP_OP_SQRT)
      U_CASES(5, MTGP_OP_TANH)
      U_CASES(6, MTGP_OP_NEG)
      U_CASES(7, MTGP_OP_ABS)
      U_CASES(8, MTGP_OP_SQUARE)
#undef U_CASES
#define C_CASES(BASE, OPC)                                                                       \
  case BASE + K_AM: {                                                                            \
    Vec<R> y = vf_load<R>(lane_base, ys);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, acc.v[r], y.v[r]);                              \
  } break;                                                                                       \
  case BASE + K_AI: MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, acc.v[r], imm); break;          \
  case BASE + K_MM: {                                                                            \
    Vec<R> x = vf_load<R>(lane_base, xs);                                                        \
    Vec<R> y = vf_load<R>(lane_base, ys);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, x.v[r], y.v[r]);                                \
  } break;                                                                                       \
  case BASE + K_MI: {                                                                            \
    Vec<R> x = vf_load<R>(lane_base, xs);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, x.v[r], imm);                                   \
  } break;
      C_CASES(H_ADD_BASE, MTGP_OP_ADD)
      C_CASES(H_MUL_BASE, MTGP_OP_MUL)
      C_CASES(H_MAX_BASE, MTGP_OP_MAX)
      C_CASES(H_MIN_BASE, MTGP_OP_MIN)
#undef C_CASES
#define N_CASES(BASE, OPC)                                                                       \
  case BASE + N_AM: {                                                                            \
    Vec<R> y = vf_load<R>(lane_base, ys);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, acc.v[r], y.v[r]);                              \
  } break;                                                                                       \
  case BASE + N_AI: MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, acc.v[r], imm); break;          \
  case BASE + N_MA: {                                                                            \
    Vec<R> x = vf_load<R>(lane_base, xs);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, x.v[r], acc.v[r]);                              \
  } break;                                                                                       \
  case BASE + N_IA: MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, imm, acc.v[r]); break;          \
  case BASE + N_MM: {                                                                            \
    Vec<R> x = vf_load<R>(lane_base, xs);                                                        \
    Vec<R> y = vf_load<R>(lane_base, ys);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, x.v[r], y.v[r]);                                \
  } break;                                                                                       \
  case BASE + N_MI: {                                                                            \
    Vec<R> x = vf_load<R>(lane_base, xs);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, x.v[r], imm);                                   \
  } break;                                                                                       \
  case BASE + N_IM: {                                                                            \
    Vec<R> y = vf_load<R>(lane_base, ys);                                                        \
    MTGP_FOR_R acc.v[r] = mtgp_apply_binary(OPC, imm, y.v[r]);                                   \
  } break;
      N_CASES(H_SUB_BASE, MTGP_OP_SUB)
      N_CASES(H_DIV_BASE, MTGP_OP_DIV)
#undef N_CASES
      default: break;
    }
    if (st != MTGP_NO_SLOT) vf_store<R>(lane_base, st, acc);
  }
  return acc;
}
