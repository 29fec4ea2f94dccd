// mtgp_lower.cuh -- lowering of one level-1 bytecode tree to an accumulator-machine program.
// Host+device so that the same code can be unit-tested on the CPU (tests/hostlib) -- the product
// only ever runs it inside k_lower (mtgp_encode.cu).
#pragma once

#include "mtgp_threaded.cuh"

#define MTGP_LHD __host__ __device__ __forceinline__

MTGP_LHD int mtgp_ffs32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v);
#else
  return __builtin_ffs((int)v);
#endif
}
MTGP_LHD int arity_of(int opcode) {
  return opcode < MTGP_OP_ADD ? 0 : (opcode <= MTGP_OP_MIN ? 2 : 1);
}
MTGP_LHD int wrap_clamp(int i, int n) {
  if (i < 0) i += n;  // NumPy-style negative index
  return (i < 0 ? 0 : (i > n - 1 ? n - 1 : i));  // XLA gather clamps
}

enum { DK_IMM = 0, DK_VAR = 1, DK_TMP = 2, DK_ACC = 3 };

struct SlotAlloc {
  uint32_t used[8];
  int high;
  MTGP_LHD void init() {
    for (int i = 0; i < 8; i++) used[i] = 0;
    high = 0;
  }
  MTGP_LHD int alloc() {
    for (int w = 0; w < 8; w++) {
      uint32_t freebits = ~used[w];
      if (freebits) {
        int b = mtgp_ffs32(freebits) - 1;
        used[w] |= 1u << b;
        int s = w * 32 + b;
        high = max(high, s + 1);
        return s;
      }
    }
    return -1;
  }
  MTGP_LHD void release(int s) { used[s >> 5] &= ~(1u << (s & 31)); }
};


// Lowers one tree (tmp_cap = temporaries the caller provisions per warp: a program may not need more).
// op/ra/rb/rc: the tree's level-1 rows; out: its program (>= mtgp_lstride(N)
// instructions).  Returns the meta word (MTGP_META_*) and the canonical-layout flag through *canonical_out.
template <int NMAX>
MTGP_LHD uint32_t mtgp_lower_tree(const uint8_t* op, const int16_t* ra, const int16_t* rb, const float* rc, int N,
                                  const mtgp_node_table& tab, int tmp_base, uint2* out, bool* canonical_out,
                                  int tmp_cap = MTGP_MAX_SLOTS) {
  uint8_t kind[NMAX];
  uint8_t ref[NMAX];
  uint32_t payload[NMAX];
  uint16_t size[NMAX];

  // ---- canonical-layout check --------------------------------------------
  bool canonical = true;
  int n_live = 0, first_live = N;
  for (int i = 0; i < N; i++) {
    const int id = op[i];
    if (id != 0) {
      n_live++;
      if (first_live == N) first_live = i;
    }
  }
  if (n_live == 0 || first_live != N - n_live) canonical = false;
  for (int i = 0; i < N && canonical; i++) {
    const int ar = arity_of(tab.opcode[op[i]]);
    const int a = ra[i], b = rb[i];
    if (i < first_live) {
      if (a != -1 || b != -1 || mtgp_f2u(rc[i]) != 0u) canonical = false;
      size[i] = 0;
      continue;
    }
    if (op[i] == 0) { canonical = false; break; }
    int sz = 1;
    if (ar >= 1) {
      if (a != i - 1 || a < first_live) { canonical = false; break; }
      sz += size[a];
      if (ar == 2) {
        const int eb = i - 1 - size[a];
        if (b != eb || eb < first_live) { canonical = false; break; }
        sz += size[eb];
      } else if (b != -1) canonical = false;
    } else if (a != -1 || b != -1) canonical = false;
    size[i] = (uint16_t)sz;
  }
  if (canonical && size[N - 1] != n_live) canonical = false;
  *canonical_out = canonical;

  // ---- pass 1: which rows does the root depend on, and how often ------------------------------
  for (int i = 0; i < N; i++) ref[i] = 0;
  ref[N - 1] = 1;
  for (int i = N - 1; i >= 0; i--) {
    if (!ref[i]) continue;
    const int ar = arity_of(tab.opcode[op[i]]);
    if (ar >= 1) {
      const int ia = wrap_clamp(ra[i], N);
      if (ia < i && ref[ia] < 255) ref[ia]++;
    }
    if (ar == 2) {
      const int ib = wrap_clamp(rb[i], N);
      if (ib < i && ref[ib] < 255) ref[ib]++;
    }
  }

  // ---- pass 2: emit -----------------------------------------------------------------------
  SlotAlloc slots;
  slots.init();
  int n = 0, acc_row = -1, last = -1;
  bool overflow = false;
  for (int i = 0; i < N; i++) {
    if (!ref[i]) continue;
    const int opcode = tab.opcode[op[i]];
    const int ar = arity_of(opcode);
    if (ar == 0) {
      if (opcode == MTGP_OP_VAR) {
        kind[i] = DK_VAR;
        payload[i] = (uint32_t)tab.arg[op[i]];
      } else {
        kind[i] = DK_IMM;
        payload[i] = (opcode == MTGP_OP_COEF) ? mtgp_f2u(rc[i]) : 0u;
      }
      continue;
    }
    // operand rows; a reference to row >= i reads that row's INITIAL column 3 (never produced by
    // the reference's operators, but well defined by body_fun: the row has not been written yet)
    int rx = wrap_clamp(ra[i], N), ry = (ar == 2) ? wrap_clamp(rb[i], N) : -1;
    int xk, yk = DK_IMM;
    uint32_t xp, yp = 0;
    if (rx >= i) { xk = DK_IMM; xp = mtgp_f2u(rc[rx]); rx = -1; } else { xk = kind[rx]; xp = payload[rx]; }
    if (ar == 2) {
      if (ry >= i) { yk = DK_IMM; yp = mtgp_f2u(rc[ry]); ry = -1; } else { yk = kind[ry]; yp = payload[ry]; }
    }
    // constant folding with the interpreter's own functions
    if (xk == DK_IMM && (ar == 1 || yk == DK_IMM)) {
      const float v = (ar == 1) ? mtgp_apply_unary(opcode, mtgp_u2f(xp))
                                : mtgp_apply_binary(opcode, mtgp_u2f(xp), mtgp_u2f(yp));
      if (rx >= 0 && ref[rx] != 255) ref[rx]--;
      if (ry >= 0 && ref[ry] != 255) ref[ry]--;
      kind[i] = DK_IMM;
      payload[i] = mtgp_f2u(v);
      continue;
    }
    // does the accumulator hold a value that outlives this instruction?
    const bool x_acc = (rx >= 0 && rx == acc_row && kind[rx] == DK_ACC);
    const bool y_acc = (!x_acc && ry >= 0 && ry == acc_row && kind[ry] == DK_ACC);
    if (acc_row >= 0 && kind[acc_row] == DK_ACC) {
      const int remaining = ref[acc_row] - ((x_acc || y_acc) ? 1 : 0);
      if (remaining > 0) {
        const int s = slots.alloc();
        if (s < 0 || s >= tmp_cap || tmp_base + s >= MTGP_MAX_SLOTS) { overflow = true; break; }  // tmp_cap: slots the kernels provision
        out[last].x = (out[last].x & ~0xFF00u) | ((uint32_t)(tmp_base + s) << 8);
        kind[acc_row] = DK_TMP;
        payload[acc_row] = (uint32_t)(tmp_base + s);
        // a second use of the same row by THIS instruction reads the spilled copy
        if (x_acc && ry == rx) { yk = DK_TMP; yp = payload[acc_row]; }
      }
    }
    if (x_acc) xk = DK_ACC; else if (rx >= 0) { xk = kind[rx]; xp = payload[rx]; }
    if (y_acc) yk = DK_ACC; else if (ry >= 0 && !(x_acc && ry == rx)) { yk = kind[ry]; yp = payload[ry]; }
    // consume the operands
    // (a saturated count is sticky: such a row keeps its slot)
    if (rx >= 0 && ref[rx] != 255 && --ref[rx] == 0 && kind[rx] == DK_TMP) slots.release((int)payload[rx] - tmp_base);
    if (ry >= 0 && ref[ry] != 255 && --ref[ry] == 0 && kind[ry] == DK_TMP) slots.release((int)payload[ry] - tmp_base);

    // one instruction: opcode + operand kinds; the (single) immediate rides in the instruction
    uint32_t xop;
    switch (opcode) {
      case MTGP_OP_ADD: xop = X_ADD; break;
      case MTGP_OP_SUB: xop = X_SUB; break;
      case MTGP_OP_MUL: xop = X_MUL; break;
      case MTGP_OP_DIV: xop = X_DIV; break;
      case MTGP_OP_MAX: xop = X_MAX; break;
      case MTGP_OP_MIN: xop = X_MIN; break;
      case MTGP_OP_SIN: xop = X_SIN; break;
      case MTGP_OP_COS: xop = X_COS; break;
      case MTGP_OP_EXP: xop = X_EXP; break;
      case MTGP_OP_LOG: xop = X_LOG; break;
      case MTGP_OP_SQRT: xop = X_SQRT; break;
      case MTGP_OP_TANH: xop = X_TANH; break;
      case MTGP_OP_NEG: xop = X_NEG; break;
      case MTGP_OP_ABS: xop = X_ABS; break;
      default: xop = X_MUL; break;  // MTGP_OP_SQUARE = x * x
    }
    if (opcode == MTGP_OP_SQUARE) { yk = xk; yp = xp; }
    const uint32_t xkind = (xk == DK_ACC) ? OK_ACC : (xk == DK_IMM ? OK_IMM : OK_MEM);
    const uint32_t ykind = (ar == 1 && opcode != MTGP_OP_SQUARE) ? OK_ACC
                         : ((yk == DK_ACC) ? OK_ACC : (yk == DK_IMM ? OK_IMM : OK_MEM));
    const uint32_t xs = (xkind == OK_MEM) ? xp : 0u, ys = (ykind == OK_MEM) ? yp : 0u;
    const float imm = (xkind == OK_IMM) ? mtgp_u2f(xp) : ((ykind == OK_IMM) ? mtgp_u2f(yp) : 0.0f);
    out[n] = mtgp_make_instr(xop, xkind, ykind, MTGP_NO_SLOT, xs, ys, imm);
    last = n++;
    acc_row = i;
    kind[i] = DK_ACC;
  }
  if (!overflow) {
    const int r = N - 1;
    if (kind[r] == DK_IMM) out[n++] = mtgp_make_instr(X_LOAD, OK_IMM, OK_ACC, MTGP_NO_SLOT, 0, 0, mtgp_u2f(payload[r]));
    else if (kind[r] != DK_ACC) out[n++] = mtgp_make_instr(X_LOAD, OK_MEM, OK_ACC, MTGP_NO_SLOT, payload[r], 0, 0.0f);
  } else {
    // cannot happen for N within the limits checked on the host; produce NaN rather than garbage
    n = 0;
    out[n++] = mtgp_make_instr(X_LOAD, OK_IMM, OK_ACC, MTGP_NO_SLOT, 0, 0, mtgp_nanf());
  }
  out[n] = mtgp_make_instr(X_END, 0, 0, MTGP_NO_SLOT, 0, 0, 0.0f);
  if (((n + 1) & 1) != 0) out[n + 1] = mtgp_make_instr(X_END, 0, 0, MTGP_NO_SLOT, 0, 0, 0.0f);  // 16-byte granules
  // length of the threaded-code form (header + expanded instructions + END)
  int n_th = 1;
  for (int i = 0; i <= n; i++) {
    mtgp_tinstr tmp[MTGP_TH_MAX_EXPAND];
    n_th += mtgp_th_expand(out[i], tmp);
  }
  return MTGP_META_MAKE(n, n_th, slots.high);
}

