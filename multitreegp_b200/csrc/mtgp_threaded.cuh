// mtgp_threaded.cuh -- the rollout kernels' interpreter: THREADED CODE over 16-byte instructions.
//
// In a rollout the same tree program runs thousands of times per trajectory with ONE value per lane,
// so the cost of an operator node is its dispatch, not its arithmetic.  Round 1's select-mask data
// path cost ~30 SASS instructions per node (profiles/r01_rollout_c5_by_line.txt).  Here every
// (operator x operand kinds x store?) combination is its own straight-line HANDLER, and every
// handler ends with its own indirect branch to the next one (threaded code, no central dispatch
// loop): ~8 SASS per node --
//     IADD  ad = lane + B          LDS  t = [ad]          (operand from the warp's value file)
//     IADD  pc += 16               LDS.128 next = [pc]    (the instruction after the next one)
//     SHL / LDC / SHF              (jump-table entry of the next handler, c[0x2])
//     FADD  acc = acc + t          BRX
// An instruction carries the handler index of its SUCCESSOR (.x), so the jump-table load does not
// wait for the instruction fetch.  MTGP_TH_SETS = 1 (default): a handler fetches its successor into
// the same registers once it has consumed its own fields.  MTGP_TH_SETS = 2 (the first version): the
// fetch alternates between two register sets (a*, b*) and every handler exists twice -- a handler of
// set a loads the next instruction into set b and jumps to a handler of set b; twice the code, slower.
//
// The interpreter is one inline-PTX block (CUDA C has no computed goto; `brx.idx` is the PTX
// form).  Everything that is not + - * load store leaves the block with a status code, is
// computed by ordinary C code (mtgp_math.h -- the same functions as constant folding and the
// oracle) and re-enters.  All arithmetic is `.rn` (never contracted, no FTZ): the results are the
// bits of the reference's float32 operators (genetic_programming.py:24-31, 356-376).
//
// Instruction = uint4: .x = handler of the NEXT instruction (mtgp_th)
//                      .y = A: x operand -- byte offset into the lane's value file (slot * 128) or immediate bits
//                      .z = B: y operand -- byte offset or immediate bits
//                      .w = C: store byte offset (…_ST handlers, ST) or slow opcode (SLOW_*)
// A program is [header][instr 0]...[END]; the header's .x names instruction 0's handler.
#pragma once

#include "mtgp_program.cuh"

#ifndef MTGP_TH_FUSE
#define MTGP_TH_FUSE 0  // pair fusion (see MTGP_TH_FUSED): measured slower, off
#endif
#ifndef MTGP_TH_SETS
#define MTGP_TH_SETS 1  // register sets of the instruction fetch: 2 = every handler exists twice (a / b), 1 = one copy (half the code; C5 74.5 vs 76.1 ms, C3 Dopri5 68.6 vs 72.5 ms)
#endif
// handler lists: arithmetic / load forms exist with and without a fused store (..._ST)
#define MTGP_TH_ARITH(X) \
  X(ADD_AM) X(ADD_AI) X(ADD_MM) X(ADD_MI) X(ADD_AA) \
  X(MUL_AM) X(MUL_AI) X(MUL_MM) X(MUL_MI) X(MUL_AA) \
  X(SUB_AM) X(RSB_AM) X(RSB_AI) X(SUB_MM) X(RSB_MI) \
  X(LOAD_M) X(LOAD_I)
#define MTGP_TH_OTHER(X) \
  X(ST) X(LDY_M) X(LDY_I) X(LDY_A) X(SLOW_A) X(SLOW_M) X(SLOW_I) X(SIN_A) X(SIN_M) X(COS_A) X(COS_M) \
  X(DIV_A) X(DIV_M) X(DIV_I) X(END) MTGP_TH_OTHER_F(X)
// FUSED PAIRS (mtgp_th_fuse): two consecutive operator nodes of an accumulator chain behind ONE dispatch.
// The head is a generic  r = fma(x, P, Q)  over x in {accumulator, memory} and y in {immediate, memory}:
//   x + y = fma(x, 1, y)    x * y = fma(x, y, -0)    c - x = fma(x, -1, c)      -- each EXACTLY the IEEE result
// of the plain operator (one rounding of the same exact value; (+0) + (-0) = +0 keeps the product's zero sign);
// with an immediate y the pair (P, Q) is written at fusion time, with a memory y it is selected by a mask word
// (two LOP3).  The tail is the same form on the accumulator.  F_<head: x y>_<tail: y>, each with a fused store.
// MEASURED (C5 workload, one B200): fusing 29 % of the dispatches this way makes the kernel SLOWER -- with two
// register sets 77 -> 106 ms (the 32 extra handlers alone, never executed, cost 77 -> 85 ms), with one register set
// 74.5 -> 83.9 ms.  A fused handler serialises two operand fetches (data slot -> address -> LDS -> select -> fma) in
// front of its branch, whereas two plain handlers overlap theirs with the branch's own latency.
// Compiled out by default (MTGP_TH_FUSE = 0); the host tests build it in (tests/hostlib) to keep the arithmetic pinned.
#if MTGP_TH_FUSE
#define MTGP_TH_OTHER_F(X) X(FDATA)
#define MTGP_TH_FUSED(X) X(F_AI_I) X(F_AI_M) X(F_MI_I) X(F_MI_M) X(F_AM_I) X(F_AM_M) X(F_MM_I) X(F_MM_M)
#else
#define MTGP_TH_OTHER_F(X)
#define MTGP_TH_FUSED(X)
#endif

enum mtgp_th : uint32_t {
#define X(n) TH_##n, TH_##n##_ST,
  MTGP_TH_ARITH(X)
#undef X
#define X(n) TH_##n,
  MTGP_TH_OTHER(X)
#undef X
#define X(n) TH_##n, TH_##n##_ST,
  MTGP_TH_FUSED(X)
#undef X
  TH_COUNT
};

// slow opcodes (status codes of the PTX block; 0 = END)
enum mtgp_sop : uint32_t {
  S_END = 0, S_SIN, S_COS, S_EXP, S_LOG, S_SQRT, S_TANH, S_NEG, S_ABS,  // unary: the order of X_SIN .. X_ABS
  S_DIV, S_MAX, S_MIN, S_ADD, S_SUB, S_MUL
};

struct __align__(16) mtgp_tinstr {
  uint32_t next, a, b, c;
};

// Worst case: a binary slow operator whose result is stored = LDY + SLOW + ST.
#define MTGP_TH_MAX_EXPAND 3

// compact instruction -> up to 3 threaded instructions; .next holds the instruction's OWN handler
// until mtgp_th_link() rewrites it.  Returns the count.
// SLOT_BYTES: bytes between value-file slots (128 for one value per lane, 1024 for the pointwise kernel's 8);
// TRIG_INBLOCK: sin / cos have in-block handlers (the rollout kernels) or leave the block like the other slow operators.
// DIV_INBLOCK: x / y has in-block handlers.  Only the pointwise kernel uses them: in the one-value block a
// div.rn.f32 brings ptxas' division subroutine into the interpreter function, which then is no leaf any more
// and costs the rollout kernels registers around every call (measured: 77 -> 83 ms on the C5 workload).
#ifndef MTGP_TH_DIV_INBLOCK
#define MTGP_TH_DIV_INBLOCK 0
#endif
template <int SLOT_BYTES = 128, bool TRIG_INBLOCK = true, bool DIV_INBLOCK = (MTGP_TH_DIV_INBLOCK != 0)>
__host__ __device__ __forceinline__ int mtgp_th_expand(uint2 c, mtgp_tinstr* out) {
  const uint32_t op = c.x & 15u, st = (c.x >> 8) & 0xFFu;
  const uint32_t xk = (c.x >> 4) & 3u, yk = (c.x >> 6) & 3u, xs = ((c.x >> 16) & 0xFFu) * (uint32_t)SLOT_BYTES,
                 ys = (c.x >> 24) * (uint32_t)SLOT_BYTES;
  const uint32_t imm = c.y, nimm = c.y ^ 0x80000000u, sto = st * (uint32_t)SLOT_BYTES;
  const bool has_st = st != MTGP_NO_SLOT;
  const uint32_t s1 = has_st ? 1u : 0u;
  auto put = [&](int i, uint32_t h, uint32_t a, uint32_t b, uint32_t cc) {
    out[i].next = h; out[i].a = a; out[i].b = b; out[i].c = cc;
  };
  if (op == X_END) { put(0, TH_END, 0, 0, 0); return 1; }
  if (op == X_LOAD && xk != OK_ACC) {
    put(0, (xk == OK_MEM ? TH_LOAD_M : TH_LOAD_I) + s1, xk == OK_MEM ? xs : imm, 0, sto);
    return 1;
  }
  if (op == X_ADD || op == X_MUL) {  // commutative: canonical forms AM AI MM MI AA
    const uint32_t base = (op == X_ADD) ? TH_ADD_AM : TH_MUL_AM;
    const uint32_t step = TH_ADD_AI - TH_ADD_AM;
    int form = -1;
    uint32_t a = 0, b = 0;
    if (xk == OK_ACC && yk == OK_ACC) form = 4;
    else if (xk == OK_ACC && yk == OK_MEM) { form = 0; b = ys; }
    else if (xk == OK_MEM && yk == OK_ACC) { form = 0; b = xs; }
    else if (xk == OK_ACC && yk == OK_IMM) { form = 1; b = imm; }
    else if (xk == OK_IMM && yk == OK_ACC) { form = 1; b = imm; }
    else if (xk == OK_MEM && yk == OK_MEM) { form = 2; a = xs; b = ys; }
    else if (xk == OK_MEM && yk == OK_IMM) { form = 3; a = xs; b = imm; }
    else if (xk == OK_IMM && yk == OK_MEM) { form = 3; a = ys; b = imm; }
    if (form >= 0) { put(0, base + step * (uint32_t)form + s1, a, b, sto); return 1; }
  }
  if (op == X_SUB) {
    int h = -1;
    uint32_t a = 0, b = 0;
    if (xk == OK_ACC && yk == OK_MEM) { h = TH_SUB_AM; b = ys; }
    else if (xk == OK_ACC && yk == OK_IMM) { h = TH_ADD_AI; b = nimm; }            // x - c == x + (-c)
    else if (xk == OK_MEM && yk == OK_ACC) { h = TH_RSB_AM; b = xs; }               // M - acc
    else if (xk == OK_IMM && yk == OK_ACC) { h = TH_RSB_AI; b = imm; }              // c - acc
    else if (xk == OK_MEM && yk == OK_MEM) { h = TH_SUB_MM; a = xs; b = ys; }
    else if (xk == OK_MEM && yk == OK_IMM) { h = TH_ADD_MI; a = xs; b = nimm; }
    else if (xk == OK_IMM && yk == OK_MEM) { h = TH_RSB_MI; a = ys; b = imm; }      // c - M
    if (h >= 0) { put(0, (uint32_t)h + s1, a, b, sto); return 1; }
  }
  // everything else: [LDY] SLOW [ST]
  uint32_t sop;
  bool binary = true;
  switch (op) {
    case X_ADD: sop = S_ADD; break;
    case X_SUB: sop = S_SUB; break;
    case X_MUL: sop = S_MUL; break;
    case X_DIV: sop = S_DIV; break;
    case X_MAX: sop = S_MAX; break;
    case X_MIN: sop = S_MIN; break;
    case X_LOAD: sop = S_ADD; binary = false; break;  // acc = acc: never emitted by the lowering pass
    default: sop = S_SIN + (op - X_SIN); binary = false; break;  // SIN COS EXP LOG SQRT TANH NEG ABS
  }
  int n = 0;
  if (TRIG_INBLOCK && (op == X_SIN || op == X_COS) && xk != OK_IMM) {  // in-block handlers (fast path inside the PTX block)
    const uint32_t h = (op == X_SIN) ? (xk == OK_ACC ? TH_SIN_A : TH_SIN_M) : (xk == OK_ACC ? TH_COS_A : TH_COS_M);
    put(n++, h, xk == OK_MEM ? xs : 0u, 0, 0);
    if (has_st) put(n++, TH_ST, 0, 0, sto);
    return n;
  }
  if (op == X_LOAD) { put(n++, TH_LDY_I, 0, 0x80000000u, 0); }  // x + (-0.0) == x
  else if (binary) put(n++, yk == OK_ACC ? TH_LDY_A : (yk == OK_MEM ? TH_LDY_M : TH_LDY_I), 0, yk == OK_MEM ? ys : imm, 0);
  if (DIV_INBLOCK && op == X_DIV)  // x / y with y loaded by the LDY before it: in-block (IEEE div.rn)
    put(n++, xk == OK_ACC ? TH_DIV_A : (xk == OK_MEM ? TH_DIV_M : TH_DIV_I), xk == OK_MEM ? xs : imm, 0, 0);
  else
    put(n++, xk == OK_ACC ? TH_SLOW_A : (xk == OK_MEM ? TH_SLOW_M : TH_SLOW_I), xk == OK_MEM ? xs : imm, 0, sop);
  if (has_st) put(n++, TH_ST, 0, 0, sto);
  return n;
}

// Dispatch.  Every handler ends with its own `brx.idx <successor's handler>` (threaded code: one taken branch
// per node, the jump-table load issued at the top of the handler, overlapping the operand load).  ptxas gives
// every brx.idx its OWN copy of the jump table in constant bank 2 -- 92 copies of 2 x 47 entries = 17 KB, which
// thrash the indexed-constant cache so badly that every dispatch waits for a miss (measured: issue utilisation
// 25 %, 287 ms on the C5 workload).  The copies are byte-identical (the table holds function-relative
// offsets), so the build redirects every jump-table LDC to the first copy of its table
// (multitreegp_b200/sass_patch.py, verified by re-disassembly): two 188-byte tables stay resident.
// MTGP_TH_CENTRAL selects the fallback that needs no patch: one shared brx.idx per register set, reached
// by a `bra` from every handler (two taken branches per node, table load after the operand wait).
// (Also measured, slower than both: own short tables for the 8 / 12 most frequent handlers.)

// header + n instructions (prog[0] = header, prog[1..n] = instructions with their OWN handler in .next):
// rewrite .next to the successor's handler.
// A fused head (mtgp_th_fuse) occupies two slots -- its own and a data slot -- and names the handler of the slot after both.
__host__ __device__ __forceinline__ void mtgp_th_link(mtgp_tinstr* prog, int n) {
  prog[0].a = 0; prog[0].b = 0; prog[0].c = 0;
  for (int j = 0; j < n; j++) {
#if MTGP_TH_FUSE
    const bool fused = j > 0 && prog[j].next >= TH_F_AI_I && prog[j].next < TH_COUNT;
#else
    const bool fused = false;
#endif
    prog[j].next = prog[j + (fused ? 2 : 1)].next;
  }
  prog[n].next = TH_END;
}

// Peephole over prog[1 .. n-1] (instructions with their OWN handler in .next, before mtgp_th_link): a store-free
// + / * / c-x node followed by an accumulator-form + / * / c-x node becomes one fused instruction:
//   head slot: .a = x offset (M heads), .b = P | y offset, .c = Q | op mask (0 = add, ~0 = mul)
//   data slot: .a = P | y offset, .b = Q | op mask, .c = store offset
__host__ __device__ __forceinline__ void mtgp_th_fuse(mtgp_tinstr* prog, int n) {
#if MTGP_TH_FUSE
  const uint32_t ONE = 0x3F800000u, MONE = 0xBF800000u, NZ = 0x80000000u;
  for (int j = 1; j + 1 < n; j++) {
    const uint32_t h = prog[j].next, t = prog[j + 1].next;
    if (t >= TH_ST) continue;
    const uint32_t tt = t & ~1u, st = t & 1u;
    int head, tail;
    uint32_t hb = prog[j].b, hc = 0, ta, tb;
    switch (h) {  // (even handler = no store)
      case TH_ADD_AI: head = 0; hc = hb; hb = ONE; break;
      case TH_MUL_AI: head = 0; hc = NZ; break;
      case TH_RSB_AI: head = 0; hc = hb; hb = MONE; break;
      case TH_ADD_MI: head = 1; hc = hb; hb = ONE; break;
      case TH_MUL_MI: head = 1; hc = NZ; break;
      case TH_RSB_MI: head = 1; hc = hb; hb = MONE; break;
      case TH_ADD_AM: head = 2; hc = 0u; break;
      case TH_MUL_AM: head = 2; hc = 0xFFFFFFFFu; break;
      case TH_ADD_MM: head = 3; hc = 0u; break;
      case TH_MUL_MM: head = 3; hc = 0xFFFFFFFFu; break;
      default: continue;
    }
    const uint32_t yb = prog[j + 1].b;
    switch (tt) {
      case TH_ADD_AI: tail = 0; ta = ONE; tb = yb; break;
      case TH_MUL_AI: tail = 0; ta = yb; tb = NZ; break;
      case TH_RSB_AI: tail = 0; ta = MONE; tb = yb; break;
      case TH_ADD_AM: tail = 1; ta = yb; tb = 0u; break;
      case TH_MUL_AM: tail = 1; ta = yb; tb = 0xFFFFFFFFu; break;
      default: continue;
    }
    prog[j].next = TH_F_AI_I + 2u * (uint32_t)(2 * head + tail) + st;
    prog[j].b = hb;
    prog[j].c = hc;
    prog[j + 1].next = TH_FDATA;
    prog[j + 1].a = ta;
    prog[j + 1].b = tb;  // .c keeps the store offset
    j++;
  }
#endif
}

// slow operators: ordinary C on the contract's functions
static __host__ __device__ __forceinline__ float mtgp_th_slow_op(uint32_t sop, float x, float y) {
  if (sop == S_SIN) return mtgp_sinf(x);
  if (sop == S_COS) return mtgp_cosf(x);
  switch (sop) {
    case S_DIV: return x / y;
    case S_MAX: return mtgp_maximumf(x, y);
    case S_MIN: return mtgp_minimumf(x, y);
    case S_EXP: return mtgp_expf(x);
    case S_LOG: return mtgp_logf(x);
    case S_SQRT: return mtgp_sqrtf(x);
    case S_TANH: return mtgp_tanhf(x);
    case S_NEG: return -x;
    case S_ABS: return mtgp_absf(x);
    case S_ADD: return x + y;
    case S_SUB: return x - y;
    default: return x * y;  // S_MUL
  }
}

#if defined(__CUDA_ARCH__)
// ---- the PTX block ---------------------------------------------------------------------------------
// operands: %0 acc  %1 pc  %2 nh (in: handler of [pc]; out: handler of the instruction after a slow exit)
//           %3 status  %4 xv  %5 yv  %6 lane  %7 parity (register set the instruction at pc belongs to: 0 = a, 1 = b)
// Macro parameters C / N: the register set of the current / next instruction (tokens a, b).
#define TH_PARNUM_a "0"
#define TH_PARNUM_b "1"
#if MTGP_TH_SETS == 1
// one register set: the successor's handler index is saved before the fetch overwrites it (ptxas folds the copy into
// the jump-table load).  In the arithmetic handlers the fetch is the LAST thing before the branch (TH_NEXT_LATE): issued
// earlier, ptxas loads into other registers and copies four words back; the branch latency covers the load either way.
#define TH_NEXT(N) " mov.u32 nx," #N "0;\n add.u32 %1,%1,16;\n ld.shared.v4.u32 {" #N "0," #N "1," #N "2," #N "3},[%1];\n"
#ifndef MTGP_TH_FETCH_LATE
#define MTGP_TH_FETCH_LATE 1
#endif
#if MTGP_TH_FETCH_LATE
#define TH_NEXT_EARLY(N) ""
#define TH_NEXT_LATE(N) TH_NEXT(N)
#else
#define TH_NEXT_EARLY(N) TH_NEXT(N)
#define TH_NEXT_LATE(N) ""
#endif
#else
#define TH_NEXT(N) " add.u32 %1,%1,16;\n ld.shared.v4.u32 {" #N "0," #N "1," #N "2," #N "3},[%1];\n"
#define TH_NEXT_EARLY(N) TH_NEXT(N)
#define TH_NEXT_LATE(N) ""
#endif
#ifdef MTGP_TH_CENTRAL
#define TH_GO(C, N) " bra DISP_" #C ";\n"               // shared dispatch of this register set
#else
#if MTGP_TH_SETS == 1
#define TH_GO(C, N) " brx.idx nx, TBLa;\n"
#else
#define TH_GO(C, N) " brx.idx " #C "0, TBL" #N ";\n"    // own dispatch (its jump-table copy is merged by sass_patch.py)
#endif
#endif
#define TH_XM(C) " add.u32 ad,%6," #C "1;\n ld.shared.f32 t0,[ad];\n"
#define TH_YM(C) " add.u32 ae,%6," #C "2;\n ld.shared.f32 t1,[ae];\n"
#define TH_XI(C) " mov.b32 t0," #C "1;\n"
#define TH_YI(C) " mov.b32 t1," #C "2;\n"
#define TH_NONE(C) ""
#define TH_XMYM(C) TH_XM(C) TH_YM(C)
#define TH_XMYI(C) TH_XM(C) TH_YI(C)
// a fused store: the address is formed before the successor is fetched (the fetch may overwrite the fields)
#define TH_STORE(C) " add.u32 ad,%6," #C "3;\n st.shared.f32 [ad],%0;\n"  // address and store in one place (handlers that fetch afterwards)
#define TH_STORE_A(C) " add.u32 sa,%6," #C "3;\n"
#define TH_STORE_D " st.shared.f32 [sa],%0;\n"
#define TH_NOSTORE_A(C) ""
#define TH_NOSTORE_D ""
#define TH_OP(PRE, OP, X, Y, C, N, ST) PRE(C) ST##_A(C) TH_NEXT_EARLY(N) " " OP " %0," X "," Y ";\n" ST##_D TH_NEXT_LATE(N) TH_GO(C, N)

#define TH_B_ADD_AM(C, N, ST) TH_OP(TH_YM, "add.rn.f32", "%0", "t1", C, N, ST)
#define TH_B_ADD_AI(C, N, ST) TH_OP(TH_YI, "add.rn.f32", "%0", "t1", C, N, ST)
#define TH_B_ADD_MM(C, N, ST) TH_OP(TH_XMYM, "add.rn.f32", "t0", "t1", C, N, ST)
#define TH_B_ADD_MI(C, N, ST) TH_OP(TH_XMYI, "add.rn.f32", "t0", "t1", C, N, ST)
#define TH_B_ADD_AA(C, N, ST) TH_OP(TH_NONE, "add.rn.f32", "%0", "%0", C, N, ST)
#define TH_B_MUL_AM(C, N, ST) TH_OP(TH_YM, "mul.rn.f32", "%0", "t1", C, N, ST)
#define TH_B_MUL_AI(C, N, ST) TH_OP(TH_YI, "mul.rn.f32", "%0", "t1", C, N, ST)
#define TH_B_MUL_MM(C, N, ST) TH_OP(TH_XMYM, "mul.rn.f32", "t0", "t1", C, N, ST)
#define TH_B_MUL_MI(C, N, ST) TH_OP(TH_XMYI, "mul.rn.f32", "t0", "t1", C, N, ST)
#define TH_B_MUL_AA(C, N, ST) TH_OP(TH_NONE, "mul.rn.f32", "%0", "%0", C, N, ST)
#define TH_B_SUB_AM(C, N, ST) TH_OP(TH_YM, "sub.rn.f32", "%0", "t1", C, N, ST)
#define TH_B_RSB_AM(C, N, ST) TH_OP(TH_YM, "sub.rn.f32", "t1", "%0", C, N, ST)
#define TH_B_RSB_AI(C, N, ST) TH_OP(TH_YI, "sub.rn.f32", "t1", "%0", C, N, ST)
#define TH_B_SUB_MM(C, N, ST) TH_OP(TH_XMYM, "sub.rn.f32", "t0", "t1", C, N, ST)
#define TH_B_RSB_MI(C, N, ST) TH_OP(TH_XMYI, "sub.rn.f32", "t1", "t0", C, N, ST)
#define TH_B_LOAD_M(C, N, ST) TH_XM(C) ST##_A(C) TH_NEXT_EARLY(N) " mov.f32 %0,t0;\n" ST##_D TH_NEXT_LATE(N) TH_GO(C, N)
#define TH_B_LOAD_I(C, N, ST) TH_XI(C) ST##_A(C) TH_NEXT_EARLY(N) " mov.f32 %0,t0;\n" ST##_D TH_NEXT_LATE(N) TH_GO(C, N)

#define TH_B_ST(C, N) TH_STORE(C) TH_NEXT(N) TH_GO(C, N)
#define TH_B_LDY_M(C, N) " add.u32 ae,%6," #C "2;\n ld.shared.f32 %5,[ae];\n" TH_NEXT(N) TH_GO(C, N)
#define TH_B_LDY_I(C, N) " mov.b32 %5," #C "2;\n" TH_NEXT(N) TH_GO(C, N)
#define TH_B_LDY_A(C, N) " mov.f32 %5,%0;\n" TH_NEXT(N) TH_GO(C, N)
// leaving the block before the next instruction is fetched: pc / nh / parity name it
#define TH_LEAVE(C, N) " mov.u32 %2," #C "0;\n add.u32 %1,%1,16;\n mov.u32 %7," TH_PARNUM_##N ";\n bra DONE;\n"
#define TH_B_SLOW_A(C, N) " mov.f32 %4,%0;\n mov.u32 %3," #C "3;\n" TH_LEAVE(C, N)
#define TH_B_SLOW_M(C, N) " add.u32 ad,%6," #C "1;\n ld.shared.f32 %4,[ad];\n mov.u32 %3," #C "3;\n" TH_LEAVE(C, N)
#define TH_B_SLOW_I(C, N) " mov.b32 %4," #C "1;\n mov.u32 %3," #C "3;\n" TH_LEAVE(C, N)
#define TH_B_END(C, N) " mov.u32 %3,0;\n bra DONE;\n"
// x / y, y from the LDY before it (IEEE division: div.rn.f32) -- see MTGP_TH_DIV_INBLOCK
#if MTGP_TH_DIV_INBLOCK
#define TH_B_DIV_A(C, N) TH_NEXT(N) " div.rn.f32 %0,%0,%5;\n" TH_GO(C, N)
#define TH_B_DIV_M(C, N) TH_XM(C) TH_NEXT(N) " div.rn.f32 %0,t0,%5;\n" TH_GO(C, N)
#define TH_B_DIV_I(C, N) TH_XI(C) TH_NEXT(N) " div.rn.f32 %0,t0,%5;\n" TH_GO(C, N)
#else  // never emitted by the expansion; the labels exist because the jump table covers the whole enumeration
#define TH_B_DIV_A(C, N) " mov.u32 %3,0xFFFF;\n bra DONE;\n"
#define TH_B_DIV_M(C, N) " mov.u32 %3,0xFFFF;\n bra DONE;\n"
#define TH_B_DIV_I(C, N) " mov.u32 %3,0xFFFF;\n bra DONE;\n"
#endif
#define TH_B_FDATA(C, N) " mov.u32 %3,0xFFFF;\n bra DONE;\n"  // a data slot is never dispatched
// fused pairs: the data slot goes to the OTHER register set, the instruction after it to this handler's own
// set once the head has consumed its fields, and execution continues in the same set (no extra registers:
// the interpreter's register budget is the rollout kernels' -- a larger one makes the caller spill around every call)
#define TH_F_DATA(N) " ld.shared.v4.u32 {" #N "0," #N "1," #N "2," #N "3},[%1+16];\n"
#define TH_F_NEXT(C) " mov.u32 nx," #C "0;\n add.u32 %1,%1,32;\n ld.shared.v4.u32 {" #C "0," #C "1," #C "2," #C "3},[%1];\n"
#define TH_F_SEL(M, K) " lop3.b32 fp," M "," K ",0x3F800000,0xE2;\n lop3.b32 fq," M "," K ",0x80000000,0xB8;\n"  // K ? (m, -0) : (1, m)
#define TH_FH_PRE_AI(C) ""
#define TH_FH_PRE_MI(C) " add.u32 ad,%6," #C "1;\n ld.shared.b32 m0,[ad];\n"
#define TH_FH_PRE_AM(C) " add.u32 ae,%6," #C "2;\n ld.shared.b32 m1,[ae];\n"
#define TH_FH_PRE_MM(C) TH_FH_PRE_MI(C) TH_FH_PRE_AM(C)
#define TH_FH_OP_AI(C) " mov.b32 fp," #C "2;\n mov.b32 fq," #C "3;\n fma.rn.f32 %0,%0,fp,fq;\n"
#define TH_FH_OP_MI(C) " mov.b32 fp," #C "2;\n mov.b32 fq," #C "3;\n fma.rn.f32 %0,m0,fp,fq;\n"
#define TH_FH_OP_AM(C) " mov.b32 fk," #C "3;\n" TH_F_SEL("m1", "fk") " fma.rn.f32 %0,%0,fp,fq;\n"
#define TH_FH_OP_MM(C) " mov.b32 fk," #C "3;\n" TH_F_SEL("m1", "fk") " fma.rn.f32 %0,m0,fp,fq;\n"
#define TH_FT_I(N) " mov.b32 fp," #N "1;\n mov.b32 fq," #N "2;\n fma.rn.f32 %0,%0,fp,fq;\n"
#define TH_FT_M(N) " add.u32 ae,%6," #N "1;\n ld.shared.b32 m1,[ae];\n mov.b32 fk," #N "2;\n" TH_F_SEL("m1", "fk") " fma.rn.f32 %0,%0,fp,fq;\n"
#define TH_F_STORE(N) " add.u32 ad,%6," #N "3;\n st.shared.f32 [ad],%0;\n"
#define TH_F_OP(H, T, C, N, ST) \
  TH_F_DATA(N) TH_FH_PRE_##H(C) TH_FH_OP_##H(C) TH_F_NEXT(C) TH_FT_##T(N) ST(N) " brx.idx nx, TBL" #C ";\n"
#define TH_F_ST_TH_STORE(N) TH_F_STORE(N)
#define TH_F_ST_TH_NOSTORE(N) ""
#if MTGP_TH_SETS == 1  // one register set: the data slot goes to the (otherwise unused) b registers
#define TH_F_DS(N) b
#else
#define TH_F_DS(N) N
#endif
#define TH_B_F_AI_I(C, N, ST) TH_F_OP(AI, I, C, TH_F_DS(N), TH_F_ST_##ST)
#define TH_B_F_AI_M(C, N, ST) TH_F_OP(AI, M, C, TH_F_DS(N), TH_F_ST_##ST)
#define TH_B_F_MI_I(C, N, ST) TH_F_OP(MI, I, C, TH_F_DS(N), TH_F_ST_##ST)
#define TH_B_F_MI_M(C, N, ST) TH_F_OP(MI, M, C, TH_F_DS(N), TH_F_ST_##ST)
#define TH_B_F_AM_I(C, N, ST) TH_F_OP(AM, I, C, TH_F_DS(N), TH_F_ST_##ST)
#define TH_B_F_AM_M(C, N, ST) TH_F_OP(AM, M, C, TH_F_DS(N), TH_F_ST_##ST)
#define TH_B_F_MM_I(C, N, ST) TH_F_OP(MM, I, C, TH_F_DS(N), TH_F_ST_##ST)
#define TH_B_F_MM_M(C, N, ST) TH_F_OP(MM, M, C, TH_F_DS(N), TH_F_ST_##ST)
// sin / cos: the operand goes to t0, cf = 0 (sin) / 1 (cos: sin(x + pi/2), one quadrant on), then the common body
#define TH_B_SIN_A(C, N) " mov.f32 t0,%0;\n mov.u32 cf,0;\n bra TRIG_" #C ";\n"
#define TH_B_SIN_M(C, N) TH_XM(C) " mov.u32 cf,0;\n bra TRIG_" #C ";\n"
#define TH_B_COS_A(C, N) " mov.f32 t0,%0;\n mov.u32 cf,1;\n bra TRIG_" #C ";\n"
#define TH_B_COS_M(C, N) TH_XM(C) " mov.u32 cf,1;\n bra TRIG_" #C ";\n"
// mtgp_sinf / mtgp_cosf (include/mtgp_math.h) for |x| < 1e5, operation for operation: Cody-Waite reduction
// (mul, add magic, sub, three fma), both kernels, quadrant select, sign.  Any lane outside the fast range
// (or NaN) sends the whole warp to the C code (Payne-Hanek) with status S_SIN / S_COS.
#define TH_TRIG_BODY(C, N)                                                                   \
  "TRIG_" #C ":\n"                                                                           \
  " abs.f32 t2,t0;\n setp.geu.f32 pq,t2,0f47C35000;\n vote.sync.any.pred pq,pq,0xffffffff;\n" \
  " @pq bra TRIGSLOW_" #C ";\n"                                                              \
  TH_NEXT(N)                                                                                 \
  " mul.rn.f32 t2,t0,0f3F22F983;\n add.rn.f32 t2,t2,0f4B400000;\n sub.rn.f32 t3,t2,0f4B400000;\n" \
  " neg.f32 t3,t3;\n fma.rn.f32 t4,t3,0f3FC90FDB,t0;\n fma.rn.f32 t4,t3,0fB33BBD2E,t4;\n"    \
  " fma.rn.f32 t4,t3,0fA6F72CED,t4;\n"                                                       \
  " mov.b32 qi,t2;\n add.u32 qi,qi,cf;\n"                                                    \
  " mul.rn.f32 t5,t4,t4;\n"                                                                  \
  " fma.rn.f32 t6,t5,0f3636C3EF,0fB950097E;\n fma.rn.f32 t6,t6,t5,0f3C088887;\n"             \
  " fma.rn.f32 t6,t6,t5,0fBE2AAAAB;\n mul.rn.f32 t3,t4,t5;\n fma.rn.f32 t6,t3,t6,t4;\n"      \
  " fma.rn.f32 t7,t5,0fB4928012,0f37D00AA9;\n fma.rn.f32 t7,t7,t5,0fBAB60B5F;\n"             \
  " fma.rn.f32 t7,t7,t5,0f3D2AAAAB;\n mul.rn.f32 t3,t5,t5;\n fma.rn.f32 t2,t5,0fBF000000,0f3F800000;\n" \
  " fma.rn.f32 t7,t3,t7,t2;\n"                                                               \
  " and.b32 ad,qi,1;\n setp.ne.u32 pq,ad,0;\n selp.f32 t6,t7,t6,pq;\n"                       \
  " shl.b32 ad,qi,30;\n and.b32 ad,ad,0x80000000;\n mov.b32 ae,t6;\n xor.b32 ae,ae,ad;\n mov.b32 %0,ae;\n" \
  TH_GO(C, N)                                                                               \
  "TRIGSLOW_" #C ":\n"                                                                       \
  " mov.f32 %4,t0;\n add.u32 %3,cf,1;\n" TH_LEAVE(C, N)

// label lists: the full table of one register set
#define TH_LABELS_ARITH_a(n) "H_" #n "_a, H_" #n "_ST_a, "
#define TH_LABELS_ARITH_b(n) "H_" #n "_b, H_" #n "_ST_b, "
#define TH_LABELS_OTHER_a(n) "H_" #n "_a, "
#define TH_LABELS_OTHER_b(n) "H_" #n "_b, "
#if MTGP_TH_SETS == 1
#define TH_CODE_ARITH(n) "H_" #n "_a:\n" TH_B_##n(a, a, TH_NOSTORE) "H_" #n "_ST_a:\n" TH_B_##n(a, a, TH_STORE)
#define TH_CODE_OTHER(n) "H_" #n "_a:\n" TH_B_##n(a, a)
#else
#define TH_CODE_ARITH(n)                                                                       \
  "H_" #n "_a:\n" TH_B_##n(a, b, TH_NOSTORE) "H_" #n "_b:\n" TH_B_##n(b, a, TH_NOSTORE)        \
  "H_" #n "_ST_a:\n" TH_B_##n(a, b, TH_STORE) "H_" #n "_ST_b:\n" TH_B_##n(b, a, TH_STORE)
#define TH_CODE_OTHER(n) "H_" #n "_a:\n" TH_B_##n(a, b) "H_" #n "_b:\n" TH_B_##n(b, a)
#endif

// Runs from the instruction at pc (handler nh, register set `parity`) until END (returns 0) or a slow
// operator (returns its opcode with xv / yv set; pc / nh / parity then name the instruction after it).
__device__ __forceinline__ uint32_t mtgp_th_block(float& acc, uint32_t& pc, uint32_t& nh, uint32_t& parity, float& xv,
                                                  float& yv, uint32_t lane) {
  uint32_t status;
  asm volatile(
      "{\n"
      " .reg .u32 a0,a1,a2,a3,b0,b1,b2,b3,ad,ae;\n"
      " .reg .f32 t0,t1,t2,t3,t4,t5,t6,t7;\n"
      " .reg .u32 qi,cf,nx,sa;\n"
      " .reg .b32 m0,m1,fp,fq,fk;\n"
      " .reg .pred pq;\n"
      "TBLa: .branchtargets " MTGP_TH_ARITH(TH_LABELS_ARITH_a) MTGP_TH_OTHER(TH_LABELS_OTHER_a) MTGP_TH_FUSED(TH_LABELS_ARITH_a) "DONE;\n"
#if MTGP_TH_SETS == 1
      " ld.shared.v4.u32 {a0,a1,a2,a3},[%1];\n"
      " brx.idx %2, TBLa;\n"
      MTGP_TH_ARITH(TH_CODE_ARITH)
      MTGP_TH_OTHER(TH_CODE_OTHER)
      MTGP_TH_FUSED(TH_CODE_ARITH)
      TH_TRIG_BODY(a, a)
#else
      "TBLb: .branchtargets " MTGP_TH_ARITH(TH_LABELS_ARITH_b) MTGP_TH_OTHER(TH_LABELS_OTHER_b) MTGP_TH_FUSED(TH_LABELS_ARITH_b) "DONE;\n"
      " setp.ne.u32 pq,%7,0;\n"
      " @pq bra ENTRY_b;\n"
      " ld.shared.v4.u32 {a0,a1,a2,a3},[%1];\n"
      " brx.idx %2, TBLa;\n"
      "ENTRY_b:\n"
      " ld.shared.v4.u32 {b0,b1,b2,b3},[%1];\n"
      " brx.idx %2, TBLb;\n"
#ifdef MTGP_TH_CENTRAL
      "DISP_a:\n brx.idx a0, TBLb;\n"
      "DISP_b:\n brx.idx b0, TBLa;\n"
#endif
      MTGP_TH_ARITH(TH_CODE_ARITH)
      MTGP_TH_OTHER(TH_CODE_OTHER)
      MTGP_TH_FUSED(TH_CODE_ARITH)
      TH_TRIG_BODY(a, b) TH_TRIG_BODY(b, a)
#endif
      "DONE:\n"
      "}\n"
      : "+f"(acc), "+r"(pc), "+r"(nh), "=r"(status), "=f"(xv), "+f"(yv), "+r"(lane), "+r"(parity)
      :
      : "memory");
  return status;
}
#endif

// One tree evaluation.  entry: shared address of the program's header; lane: the lane's base of the value file.
#ifndef MTGP_TH_INLINE
#define MTGP_TH_INLINE __noinline__
#endif
static __host__ __device__ MTGP_TH_INLINE float mtgp_run_threaded(mtgp_saddr entry, mtgp_saddr lane) {
#if defined(__CUDA_ARCH__)
  uint32_t nh, pc = entry + 16u, parity = 0u;  // instruction 0 goes to register set a
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nh) : "r"(entry));
  float acc = 0.0f, xv = 0.0f, yv = 0.0f;
  for (;;) {
    uint32_t sop = mtgp_th_block(acc, pc, nh, parity, xv, yv, lane);
    if (sop == S_END) return acc;
    asm volatile("" : "+r"(sop));  // (keeps the END test first: otherwise it is folded into the operator switch)
    acc = mtgp_th_slow_op(sop, xv, yv);
  }
#else
  // host build (tests/hostlib): the same semantics, one switch per instruction
  const mtgp_tinstr* prog = reinterpret_cast<const mtgp_tinstr*>(entry);
  const unsigned char* vb = reinterpret_cast<const unsigned char*>(lane);
  auto M = [&](uint32_t off) { float v; memcpy(&v, vb + off, 4); return v; };
  auto I = [](uint32_t bits) { return mtgp_u2f(bits); };
  float acc = 0.0f, yv = 0.0f;
  uint32_t h = prog[0].next;
  for (int j = 1;; j++) {
    const mtgp_tinstr w = prog[j];
    const uint32_t hh = h;
    h = w.next;
    bool st = false;
    switch (hh) {
#define CASE(n, EXPR) case TH_##n: acc = (EXPR); break; case TH_##n##_ST: acc = (EXPR); st = true; break;
      CASE(ADD_AM, acc + M(w.b)) CASE(ADD_AI, acc + I(w.b)) CASE(ADD_MM, M(w.a) + M(w.b)) CASE(ADD_MI, M(w.a) + I(w.b))
      CASE(ADD_AA, acc + acc)
      CASE(MUL_AM, acc * M(w.b)) CASE(MUL_AI, acc * I(w.b)) CASE(MUL_MM, M(w.a) * M(w.b)) CASE(MUL_MI, M(w.a) * I(w.b))
      CASE(MUL_AA, acc * acc)
      CASE(SUB_AM, acc - M(w.b)) CASE(RSB_AM, M(w.b) - acc) CASE(RSB_AI, I(w.b) - acc) CASE(SUB_MM, M(w.a) - M(w.b))
      CASE(RSB_MI, I(w.b) - M(w.a))
      CASE(LOAD_M, M(w.a)) CASE(LOAD_I, I(w.a))
#undef CASE
      case TH_ST: st = true; break;
      case TH_LDY_M: yv = M(w.b); break;
      case TH_LDY_I: yv = I(w.b); break;
      case TH_LDY_A: yv = acc; break;
      case TH_SIN_A: acc = mtgp_sinf(acc); break;
      case TH_SIN_M: acc = mtgp_sinf(M(w.a)); break;
      case TH_COS_A: acc = mtgp_cosf(acc); break;
      case TH_COS_M: acc = mtgp_cosf(M(w.a)); break;
      case TH_DIV_A: acc = acc / yv; break;
      case TH_DIV_M: acc = M(w.a) / yv; break;
      case TH_DIV_I: acc = I(w.a) / yv; break;
      case TH_SLOW_A: acc = mtgp_th_slow_op(w.c, acc, yv); break;
      case TH_SLOW_M: acc = mtgp_th_slow_op(w.c, M(w.a), yv); break;
      case TH_SLOW_I: acc = mtgp_th_slow_op(w.c, I(w.a), yv); break;
      default: {
#if !MTGP_TH_FUSE
        return acc;  // TH_END
#else
        if (hh < TH_F_AI_I || hh >= TH_COUNT) return acc;  // TH_END
        // fused pair: generic fma forms (see MTGP_TH_FUSED); the data slot follows
        const mtgp_tinstr d = prog[++j];
        const uint32_t f = hh - TH_F_AI_I, head = f >> 2, tail = (f >> 1) & 1u;
        auto pq = [&](uint32_t mask, float m, float& P, float& Q) {
          P = mask ? m : 1.0f;
          Q = mask ? -0.0f : m;
        };
        float P, Q;
        const float x = (head == 1 || head == 3) ? M(w.a) : acc;
        if (head <= 1) { P = I(w.b); Q = I(w.c); } else pq(w.c, M(w.b), P, Q);
        acc = fmaf(x, P, Q);
        if (tail == 0) { P = I(d.a); Q = I(d.b); } else pq(d.b, M(d.a), P, Q);
        acc = fmaf(acc, P, Q);
        if (f & 1u) memcpy(const_cast<unsigned char*>(vb) + d.c, &acc, 4);
        break;
#endif
      }
    }
    if (st) memcpy(const_cast<unsigned char*>(vb) + w.c, &acc, 4);
  }
#endif
}
