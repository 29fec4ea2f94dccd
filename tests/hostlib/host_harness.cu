// host_harness.cu -- TEST-ONLY: compiles the product's lowering pass (mtgp_lower.cuh) and
// interpreter (mtgp_program.cuh) for the HOST so that the program logic can be checked against the
// oracle in a container without a GPU.  Never linked into libmtgp_b200.so, never used by the
// product: the product runs this code on the device only.
#include <string.h>

#include <vector>

#include "../../multitreegp_b200/csrc/mtgp_lower.cuh"

static int f2i_trunc(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 2147483647;
  if (v <= -2147483648.0f) return -2147483647 - 1;
  return (int)v;
}

extern "C" __attribute__((visibility("default")))
// Encodes (same rules as k_encode), lowers and interprets ONE tree on n_pts data vectors.
// tree [N,4]; data [n_pts, n_data]; out [n_pts]; returns meta (n_instr | n_tmp << 16), flags in *flags_out.
int hostlib_eval_tree(const float* tree, int N, const mtgp_node_table* tab, const float* data, int n_pts,
                      int n_data, float* out, int* flags_out, unsigned long long* prog_out, int prog_cap,
                      float* out_wide) {
  std::vector<uint8_t> op(N);
  std::vector<int16_t> a(N), b(N);
  std::vector<float> c(N);
  int lossy = 0;
  for (int i = 0; i < N; i++) {
    float f = tree[4 * i], fa = tree[4 * i + 1], fb = tree[4 * i + 2];
    int fi = f2i_trunc(f);
    int fc = fi < 0 ? 0 : (fi > tab->n_ids - 1 ? tab->n_ids - 1 : fi);
    if ((float)fi != f || fc != fi) lossy = 1;
    if (fc == 1 && f != 1.0f) fc = 0;
    int ai = f2i_trunc(fa), bi = f2i_trunc(fb);
    int ac = ai < -32768 ? -32768 : (ai > 32767 ? 32767 : ai);
    int bc = bi < -32768 ? -32768 : (bi > 32767 ? 32767 : bi);
    if ((float)ai != fa || ac != ai || (float)bi != fb || bc != bi) lossy = 1;
    op[i] = (uint8_t)fc; a[i] = (int16_t)ac; b[i] = (int16_t)bc; c[i] = tree[4 * i + 3];
  }
  const int lstride = (N + 2 + 1) & ~1;
  std::vector<uint2> prog(lstride);
  bool canonical = false;
  const int tmp_base = n_data;
  uint32_t meta = mtgp_lower_tree<MTGP_MAX_NODES>(op.data(), a.data(), b.data(), c.data(), N, *tab, tmp_base,
                                                  prog.data(), &canonical);
  *flags_out = (canonical ? 1 : 0) | (lossy ? 2 : 0);
  const int n_tmp = (int)MTGP_META_TMP(meta);
  std::vector<float> vf((size_t)(tmp_base + n_tmp + 2) * 32, 0.0f);
  // the rollout kernels' threaded-code form of the same program (mtgp_threaded.cuh): expansion and linking are
  // the product's code; the host build interprets it with the C restatement of the PTX handlers
  const int n_instr = (int)MTGP_META_N(meta) + 1;
  std::vector<mtgp_tinstr> th(1 + MTGP_TH_MAX_EXPAND * n_instr);
  int m = 1;
  for (int i = 0; i < n_instr; i++) m += mtgp_th_expand(prog[i], th.data() + m);
  mtgp_th_fuse(th.data(), m - 1);  // the rollout kernels' pair fusion (generic fma forms)
  mtgp_th_link(th.data(), m - 1);
  if (m != (int)MTGP_META_TH(meta)) return -1;
  for (int p = 0; p < n_pts; p++) {
    for (int j = 0; j < n_data; j++) vf[j * 32] = data[(size_t)p * n_data + j];
    out[p] = mtgp_interpret<1>(prog.data(), vf.data()).v[0];
    if (out_wide)
      out_wide[p] = mtgp_run_threaded(mtgp_saddr_of(th.data()), mtgp_saddr_of(vf.data()));
  }
  if (prog_out)
    for (int i = 0; i < lstride && i < prog_cap; i++) memcpy(&prog_out[i], &prog[i], 8);
  return (int)meta;
}
