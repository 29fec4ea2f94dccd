// mtgp_threaded_r8.cuh -- the pointwise kernel's interpreter: the threaded code of mtgp_threaded.cuh with
// EIGHT values per lane (a warp evaluates one tree node on 256 data points per dispatched instruction).
//
// Same instruction format, handler set and dispatch (one brx.idx per handler over a table that the build
// merges, see mtgp_threaded.cuh / sass_patch.py); what differs is the data path:
//   * the accumulator is eight registers; a memory operand is two LDS.128 from the two float4 planes of a
//     value-file slot (slot stride 1024 B, plane stride 512 B, lane stride 16 B: conflict-free);
//   * x / y (the symbolic-regression operator set has it) runs in the block: LDY_* loads the divisor vector,
//     DIV_* divides (div.rn.f32, IEEE);
//   * every other non-arithmetic operator leaves the block with its x vector in the accumulator and its y
//     vector in a scratch slot of the value file; ordinary C code computes the eight results with the
//     contract's functions and the block is re-entered.  (sin / cos included: no in-block trig for eight values per lane.)
// Round 1's interpreter for this kernel decoded 8-byte instructions through compare trees and spent ~65
// SASS instructions per node and 256 points; a handler here is ~17 (2 LDS.128, 8 FP, fetch, dispatch).
#pragma once

#include "mtgp_threaded.cuh"

#if defined(__CUDA_ARCH__)
// operands: %0..%7 acc  %8 pc  %9 nh  %10 status  %11 lane (byte address of the lane's base in the value file)
//           %12 parity  %13 scratch (byte offset of the scratch slot that receives the y vector of a slow operator)
#ifndef MTGP_TH8_SETS
#define MTGP_TH8_SETS 1  // register sets of the instruction fetch (see MTGP_TH_SETS): 1 = one copy of every handler (C4: 15.2 ms; 2 sets: 19.7 ms -- the eight-value handlers are long and two copies do not fit the instruction cache)
#endif
#if MTGP_TH8_SETS == 1
#define T8_NEXT(N) " mov.u32 nx,i" #N "0;\n add.u32 %8,%8,16;\n ld.shared.v4.u32 {i" #N "0,i" #N "1,i" #N "2,i" #N "3},[%8];\n"
#define T8_GO(C, N) " brx.idx nx, TBLa;\n"
#ifndef MTGP_TH8_FETCH_LATE
#define MTGP_TH8_FETCH_LATE 1  // arithmetic handlers fetch their successor last (see TH_NEXT_LATE in mtgp_threaded.cuh)
#endif
#else
#define T8_NEXT(N) " add.u32 %8,%8,16;\n ld.shared.v4.u32 {i" #N "0,i" #N "1,i" #N "2,i" #N "3},[%8];\n"
#ifdef MTGP_TH_CENTRAL
#define T8_GO(C, N) " bra DISP_" #C ";\n"
#else
#define T8_GO(C, N) " brx.idx i" #C "0, TBL" #N ";\n"
#endif
#endif
#define T8_LDV(V, AD) " ld.shared.v4.f32 {" V "0," V "1," V "2," V "3},[" AD "];\n ld.shared.v4.f32 {" V "4," V "5," V "6," V "7},[" AD "+512];\n"
#define T8_STV(V0, V1, V2, V3, V4, V5, V6, V7, AD) \
  " st.shared.v4.f32 [" AD "],{" V0 "," V1 "," V2 "," V3 "};\n st.shared.v4.f32 [" AD "+512],{" V4 "," V5 "," V6 "," V7 "};\n"
#define T8_XM(C) " add.u32 ad,%11,i" #C "1;\n" T8_LDV("x", "ad")
#define T8_YM(C) " add.u32 ae,%11,i" #C "2;\n" T8_LDV("y", "ae")
#define T8_XI(C) " mov.b32 xi,i" #C "1;\n"
#define T8_YI(C) " mov.b32 yi,i" #C "2;\n"
#define T8_NONE(C) ""
#define T8_XMYM(C) T8_XM(C) T8_YM(C)
#define T8_XMYI(C) T8_XM(C) T8_YI(C)
#define T8_STORE(C) " add.u32 ad,%11,i" #C "3;\n" T8_STV("%0", "%1", "%2", "%3", "%4", "%5", "%6", "%7", "ad")
#define T8_NOSTORE(C) ""
// the store address is formed before the successor is fetched (with one register set the fetch overwrites the fields)
#define T8_STORE_A(C) " add.u32 sa,%11,i" #C "3;\n"
#define T8_STORE_D T8_STV("%0", "%1", "%2", "%3", "%4", "%5", "%6", "%7", "sa")
#define T8_NOSTORE_A(C) ""
#define T8_NOSTORE_D ""
// operand name of value r:  A = accumulator, X / Y = loaded vectors, XI / YI = broadcast immediates
#define T8_A(r) "%" #r
#define T8_X(r) "x" #r
#define T8_Y(r) "y" #r
#define T8_XIM(r) "xi"
#define T8_YIM(r) "yi"
#define T8_L(OP, XF, YF, r) " " OP " %" #r "," XF(r) "," YF(r) ";\n"
#define T8_ALL(OP, XF, YF) \
  T8_L(OP, XF, YF, 0) T8_L(OP, XF, YF, 1) T8_L(OP, XF, YF, 2) T8_L(OP, XF, YF, 3) T8_L(OP, XF, YF, 4) T8_L(OP, XF, YF, 5) \
  T8_L(OP, XF, YF, 6) T8_L(OP, XF, YF, 7)
#if MTGP_TH8_SETS == 1 && MTGP_TH8_FETCH_LATE
#define T8_NEXT_EARLY(N) ""
#define T8_NEXT_LATE(N) T8_NEXT(N)
#else
#define T8_NEXT_EARLY(N) T8_NEXT(N)
#define T8_NEXT_LATE(N) ""
#endif
#define T8_OP(PRE, OP, XF, YF, C, N, ST) PRE(C) ST##_A(C) T8_NEXT_EARLY(N) T8_ALL(OP, XF, YF) ST##_D T8_NEXT_LATE(N) T8_GO(C, N)

#define T8_B_ADD_AM(C, N, ST) T8_OP(T8_YM, "add.rn.f32", T8_A, T8_Y, C, N, ST)
#define T8_B_ADD_AI(C, N, ST) T8_OP(T8_YI, "add.rn.f32", T8_A, T8_YIM, C, N, ST)
#define T8_B_ADD_MM(C, N, ST) T8_OP(T8_XMYM, "add.rn.f32", T8_X, T8_Y, C, N, ST)
#define T8_B_ADD_MI(C, N, ST) T8_OP(T8_XMYI, "add.rn.f32", T8_X, T8_YIM, C, N, ST)
#define T8_B_ADD_AA(C, N, ST) T8_OP(T8_NONE, "add.rn.f32", T8_A, T8_A, C, N, ST)
#define T8_B_MUL_AM(C, N, ST) T8_OP(T8_YM, "mul.rn.f32", T8_A, T8_Y, C, N, ST)
#define T8_B_MUL_AI(C, N, ST) T8_OP(T8_YI, "mul.rn.f32", T8_A, T8_YIM, C, N, ST)
#define T8_B_MUL_MM(C, N, ST) T8_OP(T8_XMYM, "mul.rn.f32", T8_X, T8_Y, C, N, ST)
#define T8_B_MUL_MI(C, N, ST) T8_OP(T8_XMYI, "mul.rn.f32", T8_X, T8_YIM, C, N, ST)
#define T8_B_MUL_AA(C, N, ST) T8_OP(T8_NONE, "mul.rn.f32", T8_A, T8_A, C, N, ST)
#define T8_B_SUB_AM(C, N, ST) T8_OP(T8_YM, "sub.rn.f32", T8_A, T8_Y, C, N, ST)
#define T8_B_RSB_AM(C, N, ST) T8_OP(T8_YM, "sub.rn.f32", T8_Y, T8_A, C, N, ST)
#define T8_B_RSB_AI(C, N, ST) T8_OP(T8_YI, "sub.rn.f32", T8_YIM, T8_A, C, N, ST)
#define T8_B_SUB_MM(C, N, ST) T8_OP(T8_XMYM, "sub.rn.f32", T8_X, T8_Y, C, N, ST)
#define T8_B_RSB_MI(C, N, ST) T8_OP(T8_XMYI, "sub.rn.f32", T8_YIM, T8_X, C, N, ST)
#define T8_MOV_X(r) " mov.f32 %" #r ",x" #r ";\n"
#define T8_MOV_XI(r) " mov.f32 %" #r ",xi;\n"
#define T8_REP(F) F(0) F(1) F(2) F(3) F(4) F(5) F(6) F(7)
#define T8_B_LOAD_M(C, N, ST) T8_XM(C) ST##_A(C) T8_NEXT(N) T8_REP(T8_MOV_X) ST##_D T8_GO(C, N)
#define T8_B_LOAD_I(C, N, ST) T8_XI(C) ST##_A(C) T8_NEXT(N) T8_REP(T8_MOV_XI) ST##_D T8_GO(C, N)

#define T8_B_ST(C, N) T8_STORE(C) T8_NEXT(N) T8_GO(C, N)
// the divisor / second operand vector, kept in y0..y7 for the DIV_* or SLOW_* that follows
#define T8_MOV_YA(r) " mov.f32 y" #r ",%" #r ";\n"
#define T8_MOV_YI(r) " mov.f32 y" #r ",yi;\n"
#define T8_B_LDY_M(C, N) T8_YM(C) T8_NEXT(N) T8_GO(C, N)
#define T8_B_LDY_I(C, N) T8_YI(C) T8_REP(T8_MOV_YI) T8_NEXT(N) T8_GO(C, N)
#define T8_B_LDY_A(C, N) T8_REP(T8_MOV_YA) T8_NEXT(N) T8_GO(C, N)
#define T8_B_DIV_A(C, N) T8_NEXT(N) T8_ALL("div.rn.f32", T8_A, T8_Y) T8_GO(C, N)
#define T8_B_DIV_M(C, N) T8_XM(C) T8_NEXT(N) T8_ALL("div.rn.f32", T8_X, T8_Y) T8_GO(C, N)
#define T8_B_DIV_I(C, N) T8_XI(C) T8_NEXT(N) T8_ALL("div.rn.f32", T8_XIM, T8_Y) T8_GO(C, N)
// leaving the block for a slow operator: the x vector goes to the accumulator (which the operator overwrites
// anyway), the y vector to the scratch slot
#define T8_PARNUM_a "0"
#define T8_PARNUM_b "1"
#define T8_LEAVE(C, N) \
  " add.u32 ae,%11,%13;\n" T8_STV("y0", "y1", "y2", "y3", "y4", "y5", "y6", "y7", "ae") \
  " mov.u32 %10,i" #C "3;\n mov.u32 %9,i" #C "0;\n add.u32 %8,%8,16;\n mov.u32 %12," T8_PARNUM_##N ";\n bra DONE;\n"
#define T8_B_SLOW_A(C, N) T8_LEAVE(C, N)
#define T8_B_SLOW_M(C, N) T8_XM(C) T8_REP(T8_MOV_X) T8_LEAVE(C, N)
#define T8_B_SLOW_I(C, N) T8_XI(C) T8_REP(T8_MOV_XI) T8_LEAVE(C, N)
// sin / cos have no in-block form here: the expansion (TRIG_INBLOCK = false) never emits these handlers; the
// labels exist because the jump table covers the whole handler enumeration
#define T8_B_SIN_A(C, N) " mov.u32 %10,0xFFFF;\n bra DONE;\n"
#define T8_B_SIN_M(C, N) " mov.u32 %10,0xFFFF;\n bra DONE;\n"
#define T8_B_COS_A(C, N) " mov.u32 %10,0xFFFF;\n bra DONE;\n"
#define T8_B_COS_M(C, N) " mov.u32 %10,0xFFFF;\n bra DONE;\n"
#define T8_B_END(C, N) " mov.u32 %10,0;\n bra DONE;\n"
#define T8_B_FDATA(C, N) " mov.u32 %10,0xFFFF;\n bra DONE;\n"  // (pair fusion is the rollout kernels'; never emitted here)

#if MTGP_TH8_SETS == 1
#define T8_CODE_ARITH(n) "H_" #n "_a:\n" T8_B_##n(a, a, T8_NOSTORE) "H_" #n "_ST_a:\n" T8_B_##n(a, a, T8_STORE)
#define T8_CODE_OTHER(n) "H_" #n "_a:\n" T8_B_##n(a, a)
#else
#define T8_CODE_ARITH(n)                                                                       \
  "H_" #n "_a:\n" T8_B_##n(a, b, T8_NOSTORE) "H_" #n "_b:\n" T8_B_##n(b, a, T8_NOSTORE)        \
  "H_" #n "_ST_a:\n" T8_B_##n(a, b, T8_STORE) "H_" #n "_ST_b:\n" T8_B_##n(b, a, T8_STORE)
#define T8_CODE_OTHER(n) "H_" #n "_a:\n" T8_B_##n(a, b) "H_" #n "_b:\n" T8_B_##n(b, a)
#endif

__device__ __forceinline__ uint32_t mtgp_th8_block(float (&acc)[8], uint32_t& pc, uint32_t& nh, uint32_t& parity,
                                                   uint32_t lane, uint32_t scratch) {
  uint32_t status;
  asm volatile(
      "{\n"
      " .reg .u32 ia0,ia1,ia2,ia3,ib0,ib1,ib2,ib3,ad,ae,nx,sa;\n"
      " .reg .f32 x0,x1,x2,x3,x4,x5,x6,x7,y0,y1,y2,y3,y4,y5,y6,y7,xi,yi;\n"
      " .reg .pred pq;\n"
      "TBLa: .branchtargets " MTGP_TH_ARITH(TH_LABELS_ARITH_a) MTGP_TH_OTHER(TH_LABELS_OTHER_a) "DONE;\n"
#if MTGP_TH8_SETS == 1
      " ld.shared.v4.u32 {ia0,ia1,ia2,ia3},[%8];\n"
      " brx.idx %9, TBLa;\n"
#else
      "TBLb: .branchtargets " MTGP_TH_ARITH(TH_LABELS_ARITH_b) MTGP_TH_OTHER(TH_LABELS_OTHER_b) "DONE;\n"
      " setp.ne.u32 pq,%12,0;\n"
      " @pq bra ENTRY_b;\n"
      " ld.shared.v4.u32 {ia0,ia1,ia2,ia3},[%8];\n"
      " brx.idx %9, TBLa;\n"
      "ENTRY_b:\n"
      " ld.shared.v4.u32 {ib0,ib1,ib2,ib3},[%8];\n"
      " brx.idx %9, TBLb;\n"
#endif
#ifdef MTGP_TH_CENTRAL
      "DISP_a:\n brx.idx ia0, TBLb;\n"
      "DISP_b:\n brx.idx ib0, TBLa;\n"
#endif
      MTGP_TH_ARITH(T8_CODE_ARITH)
      MTGP_TH_OTHER(T8_CODE_OTHER)
      "DONE:\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7]),
        "+r"(pc), "+r"(nh), "=r"(status), "+r"(lane), "+r"(parity), "+r"(scratch)
      :
      : "memory");
  return status;
}

// the slow operators on eight values (out of line: rare, and large)
static __device__ __noinline__ Vec<8> mtgp_slow8(uint32_t sop, Vec<8> x, Vec<8> y) {
  Vec<8> r;
#pragma unroll 1
  for (int i = 0; i < 8; i++) r.v[i] = mtgp_th_slow_op(sop, x.v[i], y.v[i]);
  return r;
}

// One tree on eight values per lane.  entry: shared address of the program header; lane: byte address of the
// lane's base of the value file (base + 16 * lane); scratch: byte offset of one free slot.
static __device__ __noinline__ Vec<8> mtgp_run_threaded8(mtgp_saddr entry, mtgp_saddr lane, uint32_t scratch) {
  uint32_t nh, pc = entry + 16u, parity = 0u;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nh) : "r"(entry));
  Vec<8> acc;
#pragma unroll
  for (int r = 0; r < 8; r++) acc.v[r] = 0.0f;
  for (;;) {
    uint32_t sop = mtgp_th8_block(acc.v, pc, nh, parity, lane, scratch);
    if (sop == S_END) return acc;
    asm volatile("" : "+r"(sop));
    Vec<8> y;  // the second operand vector from the scratch slot (two float4 planes)
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const uint4 ya = mtgp_lds_v4(lane + scratch + 512u * i);
      y.v[4 * i] = mtgp_u2f(ya.x); y.v[4 * i + 1] = mtgp_u2f(ya.y); y.v[4 * i + 2] = mtgp_u2f(ya.z); y.v[4 * i + 3] = mtgp_u2f(ya.w);
    }
    acc = mtgp_slow8(sop, acc, y);
  }
}
#else
// host compilation pass: declaration only (the kernel bodies that call it are device code)
__device__ Vec<8> mtgp_run_threaded8(mtgp_saddr entry, mtgp_saddr lane, uint32_t scratch);
#endif
