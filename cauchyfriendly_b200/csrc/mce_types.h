// mce_types.h -- device data model of the B200 MCE term-propagation path.
//
// Vocabulary follows the reference: a *term* has m hyperplane rows A (m x d), weights p (q = p before L1
// normalisation), a shift b, and a G-table keyed by encoded sign vectors (bit l set <=> sign -1,
// cell_enumeration.hpp:58).  Only the half of every table with bit m-1 clear is stored (HALF_STORAGE,
// cauchy_types.hpp:28,34); keys are kept sorted (BINSEARCH_STORAGE, cauchy_types.hpp:22-24).
#ifndef MCE_TYPES_H_
#define MCE_TYPES_H_

#include "mce_math.h"

namespace mce {

constexpr int MAXD = 8;      // state dimension supported by the kernels' register/local arrays
constexpr int MAXM = 32;     // hyperplane rows incl. the virtual measurement row (reference cap 31, est:231-235)
constexpr int MAXPN = 4;     // process-noise columns after pre-coalignment
constexpr int NSHAPE = MAXM + 1;

// Thresholds: part of the parity contract (cauchy_constants.hpp:29-73).
constexpr double COALIGN_EPS = 1e-8;          // COALIGN_TP_EPS == COALIGN_MU_EPS
constexpr double MU_EPS = 1e-10;
constexpr double REDUCTION_EPS = 1e-8;
constexpr double TERM_APPROXIMATION_EPS = 1e-15;
constexpr double PLU_EPS = 1e-15;
constexpr double COND_EPS = 1e12;
constexpr int DCE_STORAGE_MULT = 4;

// Everything a step needs that is the same for all terms.  Passed to kernels by value.
struct StepParams {
  int d, with_tp, npn, has_bu, skip_post_mu, max_shape;
  int tr_order[12];
  double Phi[MAXD * MAXD];
  double GammaT[MAXPN * MAXD];   // rows L1-normalised and pre-coaligned on the host (cauchy_util.hpp:128)
  double beta[MAXPN];
  double H[MAXD];
  double bu[MAXD];               // B @ u (cauchy_term.hpp:451-455)
  double msmt, gamma, gscale;
  double root_point[MAXD];
  double b_pert[MAXM];
};

// One generation of parent terms (the survivors of the previous step).  A parent is addressed by its
// group id `gid` (groups of one shape are contiguous); `alive` lists the surviving gids in canonical
// order (shape ascending, then root index) -- the order the NUM_CPUS=1 reference walks them.
#if defined(__CUDA_ARCH__)
#define MCE_POPC(x) __popc(x)
#define MCE_FFS(x) (__ffs((int)(x)) - 1)
#define MCE_NOINLINE __noinline__
#define MCE_NOUNROLL _Pragma("unroll 1")
#else
#define MCE_NOUNROLL
#define MCE_POPC(x) __builtin_popcount(x)
#define MCE_FFS(x) (__builtin_ffs((int)(x)) - 1)
#define MCE_NOINLINE
#endif

// rank of `key` among the set bits of bm (pf = exclusive prefix popcounts per word), or -1 when absent
MCE_HD int bitmap_rank(const unsigned* bm, const unsigned short* pf, unsigned key) {
  const unsigned w = bm[key >> 5], bit = 1u << (key & 31);
  if (!(w & bit)) return -1;
  return (int)pf[key >> 5] + MCE_POPC(w & (bit - 1u));
}

struct GenView {
  int n_groups, n_alive;
  int gid_begin[NSHAPE + 1];          // gids of shape m: [gid_begin[m], gid_begin[m+1])
  long long A_base[NSHAPE], p_base[NSHAPE], tab_base[NSHAPE];
  int tab_stride[NSHAPE];
  unsigned char* g_m;                 // [n_groups]
  int* cells;                         // [n_groups]
  int* alive;                         // [n_alive]
  double *A, *p, *b;
  unsigned* keys;
  cplx* G;
  // rank structure of every table (KBuildRank): bitmap over the 2^m possible keys + exclusive prefix popcounts per word;
  // the rank of a key is its position in the sorted table, so a lookup is two loads and a popcount (no binary search)
  long long rk_base[NSHAPE];
  unsigned* rbm;
  unsigned short* rpf;
  double* gmax;                       // [n_groups] max over the table's cells of |re| + |im| (>= |G|), written by KBuildRank
};
MCE_HD int gen_m(const GenView& g, int gid) { return g.g_m[gid]; }
MCE_HD int rank_words(int m) { return m >= 5 ? (1 << (m - 5)) : 1; }
MCE_HD long long gen_rk_off(const GenView& g, int gid, int m) { return g.rk_base[m] + (long long)(gid - g.gid_begin[m]) * rank_words(m); }
MCE_HD double* gen_A(const GenView& g, int gid, int m, int d) { return g.A + g.A_base[m] + (long long)(gid - g.gid_begin[m]) * m * d; }
MCE_HD double* gen_p(const GenView& g, int gid, int m) { return g.p + g.p_base[m] + (long long)(gid - g.gid_begin[m]) * m; }
MCE_HD double* gen_b(const GenView& g, int gid, int d) { return g.b + (long long)gid * d; }
MCE_HD unsigned* gen_keys(const GenView& g, int gid, int m) { return g.keys + g.tab_base[m] + (long long)(gid - g.gid_begin[m]) * g.tab_stride[m]; }
MCE_HD cplx* gen_G(const GenView& g, int gid, int m) { return g.G + g.tab_base[m] + (long long)(gid - g.gid_begin[m]) * g.tab_stride[m]; }

// Per-parent scratch of one step, indexed by alive rank r.
struct ParentWs {
  double *A, *p, *b;          // time-propagated + Gamma-coaligned hyperplanes, [r][max_shape*d], [max_shape], [d] (TP steps)
  unsigned char* m_tp;        // [r]
  unsigned* sgnmask;          // [r] half-normalised sign(A H) mask: B_mu = B ^ sgnmask (cauchy_term.hpp:217-250)
  unsigned* bxor;             // [r] in-place re-orientations of the parent's B memory (flattening.hpp:433-441)
  unsigned* tpB;              // [r][tpB_stride] B^{k|k-1} after DCE-TP, sorted (TP steps)
  int* tpB_cells;             // [r]
  int tpB_stride;
};

struct SlotMeta {
  unsigned char newm;         // shape after MU coalignment; 0 = slot unused
  unsigned char pbc, z, flags;  // flags: bit0 is_new_child, bit1 has coalignment maps
  unsigned hflag;             // Horthog_flag
  unsigned enc_lhp;
  unsigned csneg;             // bit l set <=> cs_map[l] == -1
  int parent;                 // alive rank of the parent
  int pad_;
  double c_val, d_val;
};

// Child-term slots written by the MU kernel: parent r of old shape m owns MT(m)+1 slots
// (slot 0 = the parent as "old term", slot 1+t = child integrated over row t), MT = m + npn.
struct SlotView {
  int par_begin[NSHAPE + 1];          // alive ranks of old shape m: [par_begin[m], par_begin[m+1])
  long long slot_begin[NSHAPE + 1];   // global slot ids of the region of old shape m
  int MT[NSHAPE];
  long long A_off[NSHAPE], pq_off[NSHAPE];
  long long n_slots;
  double *A, *p, *q, *b;
  SlotMeta* meta;
  unsigned char* cmap;                // [slot][MAXM]
  cplx* g;                            // [slot] g value (moment evaluation, cauchy_term.hpp:312)
  double* y;                          // [slot][2d]  yei = (-sum p s a, b)
};

// Terms after MU coalignment regrouped by their new shape, in the canonical order
// (all old terms, then all children in generation order; cauchy_estimator.hpp:756-774).
struct TermView {
  int n[NSHAPE], n_old[NSHAPE];
  long long t_begin[NSHAPE + 1];      // global term index of the first term of shape m
  long long A_base[NSHAPE], pq_base[NSHAPE];
  double *A, *p, *q, *b;
  SlotMeta* meta;
  unsigned char* cmap;                // [gt][MAXM]
};
MCE_HD double* term_A(const TermView& t, int m, int i, int d) { return t.A + t.A_base[m] + (long long)i * m * d; }
MCE_HD double* term_p(const TermView& t, int m, int i) { return t.p + t.pq_base[m] + (long long)i * m; }
MCE_HD double* term_q(const TermView& t, int m, int i) { return t.q + t.pq_base[m] + (long long)i * m; }
MCE_HD double* term_b(const TermView& t, int m, int i, int d) { return t.b + (t.t_begin[m] + i) * d; }

MCE_HD int cell_count_central_half(int hyp, int dim) {   // cell_enumeration.hpp:34-44, halved
  if (hyp < dim) return 1 << (hyp - 1 < 0 ? 0 : hyp - 1);
  unsigned long long fc = 0;
  for (int i = 0; i < dim; i++) {
    unsigned long long res = 1; int n = hyp - 1, k = i;
    if (k > n - k) k = n - k;
    for (int j = 0; j < k; ++j) { res *= (unsigned long long)(n - j); res /= (unsigned long long)(j + 1); }
    fc += res;
  }
  return (int)fc;
}
MCE_HD int cell_count_general(int hyp, int dim) {        // cell_enumeration.hpp:46-56
  if (hyp < dim) return 1 << hyp;
  unsigned long long fc = 0;
  for (int i = 0; i < dim + 1; i++) {
    unsigned long long res = 1; int n = hyp, k = i;
    if (k > n - k) k = n - k;
    for (int j = 0; j < k; ++j) { res *= (unsigned long long)(n - j); res /= (unsigned long long)(j + 1); }
    fc += res;
  }
  return (int)fc;
}

// Binary search of a sorted u32 key array (gtable.hpp:283-301).
MCE_HD int key_search(const unsigned* keys, int n, unsigned target) {
  int low = 0, high = n - 1;
  while (low <= high) {
    int mid = (low + high) / 2;
    unsigned mk = keys[mid];
    if (mk == target) return mid;
    else if (mk > target) high = mid - 1;
    else low = mid + 1;
  }
  return -1;
}

// G_p lookup with half storage: the opposite cell's value is the complex conjugate (eval_gs.hpp:94-153).
// same lookup through the table's rank structure (GenView::rbm / rpf)
MCE_HD cplx g_lookup_rank(int enc_l, int phc, const unsigned* bm, const unsigned short* pf, const cplx* pG) {
  const int top = 1 << (phc - 1), rev = (1 << phc) - 1;
  const bool cj = (enc_l & top) != 0;
  const int r = bitmap_rank(bm, pf, (unsigned)(cj ? (rev ^ enc_l) : enc_l));
  if (r < 0) return make_cplx(0, 0);
  return cj ? cconj(pG[r]) : pG[r];
}
MCE_HD cplx g_lookup(int enc_l, int phc, const unsigned* pkeys, const cplx* pG, int pcells) {
  const int two_to_phc_minus1 = 1 << (phc - 1), rev_phc_mask = (1 << phc) - 1;
  if (enc_l & two_to_phc_minus1) {
    int idx = key_search(pkeys, pcells, (unsigned)(rev_phc_mask ^ enc_l));
    if (idx < 0) return make_cplx(0, 0);
    return cconj(pG[idx]);
  }
  int idx = key_search(pkeys, pcells, (unsigned)enc_l);
  if (idx < 0) return make_cplx(0, 0);
  return pG[idx];
}

// Parent sign-vector keys lambda+ / lambda- of a cell (flattening.hpp:156-218, cauchy_term.hpp:361-385):
// `signs` holds bit l = sign of the term's (post-coalignment) row l.
MCE_HD void parent_keys(unsigned signs, int m, int phc, int z, bool is_child, const unsigned char* cmap, unsigned csneg,
                        int* enc_lp, int* enc_lm) {
  const int phc_mask = (1 << phc) - 1;
  if (!is_child) { *enc_lp = (int)(signs & (unsigned)phc_mask); *enc_lm = *enc_lp; return; }
  int lp = 0, lm = 0, k = 0, l = 0;
  while (k < phc) {
    if (k == z) { lm |= (1 << k); k++; if (k == phc) break; }
    unsigned bit;
    if (cmap == nullptr) bit = (signs >> l) & 1u;
    else bit = ((signs >> cmap[l]) & 1u) ^ ((csneg >> l) & 1u);
    if (bit) { lp |= (1 << k); lm |= (1 << k); }
    k++; l++;
  }
  (void)m;
  *enc_lp = lp; *enc_lm = lm;
}

}  // namespace mce
#endif
