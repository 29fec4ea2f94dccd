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
// Instruction = 8 bytes:  .x = opcode | x kind << 4 | y kind << 6 | store_slot << 8 | x_slot << 16 | y_slot << 24
//                         .y = immediate (float bits)          kinds: accumulator / memory slot / immediate
// Slot s of a warp's value file is the 32*R floats at  base + s*32*R.  R = 1: lane l owns element l.
// R = 4k: the slot is k PLANES of 32 float4, lane l owns float4 l of every plane (lane base = base + 4 l,
// plane i at + 128 i floats) -- every 16-byte access of a warp covers 512 contiguous bytes, conflict-free.
// (Round 1 gave a lane its R values contiguously: a 32-byte lane stride, 2-way conflicts on every LDS.128.)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mtgp.h"
#include "../../include/mtgp_math.h"

#define MTGP_NO_SLOT 0xFFu
#define MTGP_MAX_SLOTS 254

// meta word of a lowered tree (k_lower): n_instr (compact instructions, X_END excluded) in bits 0-9,
// n_threaded (threaded-code instructions incl. header and END, mtgp_threaded.cuh) in bits 10-21,
// n_tmp (temporary value-file slots) in bits 22-31
#define MTGP_META_N(m) ((uint32_t)(m) & 0x3FFu)
#define MTGP_META_TH(m) (((uint32_t)(m) >> 10) & 0xFFFu)
#define MTGP_META_TMP(m) ((uint32_t)(m) >> 22)
#define MTGP_META_MAKE(n, th, tmp) ((uint32_t)(n) | ((uint32_t)(th) << 10) | ((uint32_t)(tmp) << 22))

// Instruction word .x = opcode (bits 0-3) | x kind (bits 4-5) | y kind (bits 6-7) | store slot << 8
//                       | x slot << 16 | y slot << 24;   .y = immediate (float bits).
// The opcode set is closed and small on purpose: dispatch is one short compare chain and the three
// most frequent operators (+ - *) share one branch-free data path.
enum mtgp_xop : uint32_t {
  X_END = 0,
  X_LOAD = 1,  // acc = x
  X_ADD = 2,
  X_SUB = 3,
  X_MUL = 4,   // MTGP_OP_SQUARE is lowered to x * x
  X_DIV = 5,
  X_MAX = 6,
  X_MIN = 7,
  X_SIN = 8,
  X_COS = 9,
  X_EXP = 10,
  X_LOG = 11,
  X_SQRT = 12,
  X_TANH = 13,
  X_NEG = 14,
  X_ABS = 15
};
enum { OK_ACC = 0, OK_MEM = 1, OK_IMM = 2 };  // operand kinds

__host__ __device__ __forceinline__ uint2 mtgp_make_instr(uint32_t op, uint32_t xk, uint32_t yk, uint32_t st,
                                                          uint32_t xs, uint32_t ys, float imm) {
  uint2 w;
  w.x = (op & 15u) | ((xk & 3u) << 4) | ((yk & 3u) << 6) | ((st & 0xFFu) << 8) | ((xs & 0xFFu) << 16) |
        ((ys & 0xFFu) << 24);
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
      float4 t = *reinterpret_cast<const float4*>(p + 128 * i);
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
      *reinterpret_cast<float4*>(p + 128 * i) = make_float4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < R; i++) p[i] = a.v[i];
  }
}

#define MTGP_FOR_R _Pragma("unroll") for (int r = 0; r < R; r++)

// ---- compact-format interpreter (R values per lane; used by the pointwise kernel) ---------------
// prog: shared-memory program (terminated by X_END), identical for all lanes of the warp, so every
// branch below is warp-uniform and its cost is amortised over the 32*R values a warp evaluates per
// dispatched instruction.  lane_base = value file + lane*R.
#ifndef MTGP_COMPACT_INLINE
#define MTGP_COMPACT_INLINE __noinline__
#endif
template <int R>
__host__ __device__ MTGP_COMPACT_INLINE Vec<R> mtgp_interpret(const uint2* __restrict__ prog, float* lane_base) {
  Vec<R> acc;
  MTGP_FOR_R acc.v[r] = 0.0f;
  uint2 w = prog[0];
  int pc = 0;
  while ((w.x & 15u) != X_END) {
    const uint32_t c = w.x;
    const float imm = mtgp_u2f(w.y);
    w = prog[++pc];  // prefetch the next instruction under this one's arithmetic
    const uint32_t op = c & 15u, xk = (c >> 4) & 3u, yk = (c >> 6) & 3u, st = (c >> 8) & 0xFFu;
    const uint32_t xs = (c >> 16) & 0xFFu, ys = c >> 24;
    // + - * and loads: one straight-line leaf per operand-kind combination, operands streamed from
    // the value file straight into the arithmetic -- no control-flow merge, so no register copies
#define MTGP_LEAF(EXPR) { MTGP_FOR_R acc.v[r] = (EXPR); }
#define MTGP_BINOP(F)                                                                                       \
  if (xk == OK_ACC) {                                                                                       \
    if (yk == OK_MEM) { const Vec<R> y = vf_load<R>(lane_base, ys); MTGP_LEAF(F(acc.v[r], y.v[r])) }        \
    else if (yk == OK_IMM) MTGP_LEAF(F(acc.v[r], imm))                                                      \
    else MTGP_LEAF(F(acc.v[r], acc.v[r]))                                                                   \
  } else if (xk == OK_MEM) {                                                                                \
    const Vec<R> x = vf_load<R>(lane_base, xs);                                                             \
    if (yk == OK_MEM) { const Vec<R> y = vf_load<R>(lane_base, ys); MTGP_LEAF(F(x.v[r], y.v[r])) }          \
    else if (yk == OK_IMM) MTGP_LEAF(F(x.v[r], imm))                                                        \
    else MTGP_LEAF(F(x.v[r], acc.v[r]))                                                                     \
  } else {                                                                                                  \
    if (yk == OK_MEM) { const Vec<R> y = vf_load<R>(lane_base, ys); MTGP_LEAF(F(imm, y.v[r])) }             \
    else MTGP_LEAF(F(imm, acc.v[r]))                                                                        \
  }
#define MTGP_F_ADD(a, b) ((a) + (b))
#define MTGP_F_SUB(a, b) ((a) - (b))
#define MTGP_F_MUL(a, b) ((a) * (b))
    if (op == X_ADD) {
      MTGP_BINOP(MTGP_F_ADD)
    } else if (op == X_MUL) {
      MTGP_BINOP(MTGP_F_MUL)
    } else if (op == X_SUB) {
      MTGP_BINOP(MTGP_F_SUB)
    } else if (op == X_LOAD) {
      if (xk == OK_MEM) { const Vec<R> x = vf_load<R>(lane_base, xs); MTGP_LEAF(x.v[r]) }
      else MTGP_LEAF(imm)
    } else {
      Vec<R> x = acc, y = acc;
      if (xk == OK_MEM) x = vf_load<R>(lane_base, xs);
      if (yk == OK_MEM) y = vf_load<R>(lane_base, ys);
      if (xk == OK_IMM) { MTGP_FOR_R x.v[r] = imm; }
      if (yk == OK_IMM) { MTGP_FOR_R y.v[r] = imm; }
      switch (op) {
        case X_DIV: MTGP_FOR_R acc.v[r] = x.v[r] / y.v[r]; break;
        case X_MAX: MTGP_FOR_R acc.v[r] = mtgp_maximumf(x.v[r], y.v[r]); break;
        case X_MIN: MTGP_FOR_R acc.v[r] = mtgp_minimumf(x.v[r], y.v[r]); break;
        case X_SIN: MTGP_FOR_R acc.v[r] = mtgp_sinf(x.v[r]); break;
        case X_COS: MTGP_FOR_R acc.v[r] = mtgp_cosf(x.v[r]); break;
        case X_EXP: MTGP_FOR_R acc.v[r] = mtgp_expf(x.v[r]); break;
        case X_LOG: MTGP_FOR_R acc.v[r] = mtgp_logf(x.v[r]); break;
        case X_SQRT: MTGP_FOR_R acc.v[r] = mtgp_sqrtf(x.v[r]); break;
        case X_TANH: MTGP_FOR_R acc.v[r] = mtgp_tanhf(x.v[r]); break;
        case X_NEG: MTGP_FOR_R acc.v[r] = -x.v[r]; break;
        default: MTGP_FOR_R acc.v[r] = mtgp_absf(x.v[r]); break;  // X_ABS
      }
    }
#undef MTGP_LEAF
#undef MTGP_BINOP
    if (st != MTGP_NO_SLOT) vf_store<R>(lane_base, st, acc);
  }
  return acc;
}

// Shared-memory accessors on 32-bit shared-space addresses (device) / plain pointers (host build of
// the unit-test harness).  One IADD + LDS/STS per value-file access, one LDS.128 per half instruction.
#if defined(__CUDA_ARCH__)
typedef uint32_t mtgp_saddr;
__device__ __forceinline__ mtgp_saddr mtgp_saddr_of(const void* p) { return (mtgp_saddr)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float mtgp_lds_f32(mtgp_saddr a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void mtgp_sts_f32(mtgp_saddr a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ uint4 mtgp_lds_v4(mtgp_saddr a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
#else
typedef uintptr_t mtgp_saddr;
static inline mtgp_saddr mtgp_saddr_of(const void* p) { return (mtgp_saddr)p; }
static inline float mtgp_lds_f32(mtgp_saddr a) { float v; memcpy(&v, (const void*)a, 4); return v; }
static inline void mtgp_sts_f32(mtgp_saddr a, float v) { memcpy((void*)a, &v, 4); }
static inline uint4 mtgp_lds_v4(mtgp_saddr a) { uint4 v; memcpy(&v, (const void*)a, 16); return v; }
#endif

