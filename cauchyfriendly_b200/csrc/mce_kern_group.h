// mce_kern_group.h -- kernels K2 (DCE-TP), K7 (child B-table) and K8 (G-table build / add / term
// approximation / root re-election), one CTA per parent (K2) or per reduction group (K7+K8).
// Reference loops replaced: cell_enumeration.hpp:676-855 (make_time_prop_btable), 206-371
// (make_new_child_btable), 584-625 (update_btable); flattening.hpp:77-255 (make_gtable), 259-315
// (add_gtables), 318-568 (make_gtables group loop).
//
// Sign vectors are 32-bit keys; every table lives in shared memory while a group is processed:
//   Bk  current B-table of the group (cells of the root's arrangement)
//   Pk  the evaluated term's parent keys (sorted) for the two G_p lookups per cell
//   acc root G-table accumulator, Gm member G-table
// Only the root's table is ever written to HBM (members are accumulated on chip, flattening.hpp:491-550).
#ifndef MCE_KERN_GROUP_H_
#define MCE_KERN_GROUP_H_

#include "mce_exec.h"
#include "mce_types.h"
#include "mce_kern_prop.h"

namespace mce {

MCE_HD int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }
MCE_HD unsigned hash_u32(unsigned k) { k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16; return k; }

// Block-wide bitonic sort of arr[0..n2) (n2 a power of two; pad with the maximum value).
template <class Ctx, class T>
MCE_KERNEL_FN void block_sort(Ctx& c, T* arr, int n2) {
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1)
      c.par([&](int tid) {
        for (int i = tid; i < n2; i += c.nthreads()) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const T a = arr[i], b = arr[ixj];
            const bool up = (i & k) == 0;
            if (up ? (a > b) : (a < b)) { arr[i] = b; arr[ixj] = a; }
          }
        }
      });
}

// update_btable's sigma (cell_enumeration.hpp:586-599): per row, the first component where both rows are
// >= eps in magnitude decides whether the two (equal up to sign) hyperplanes point the same way.
MCE_HD unsigned orient_bit(const double* Ai, const double* Aj, int k, int d) {
  int l = 0;
  while (l < d && ((fabs(Ai[k * d + l]) < REDUCTION_EPS) || (fabs(Aj[k * d + l]) < REDUCTION_EPS))) l++;
  if (l >= d) return 0;      // the reference would run off the row here
  return (Ai[k * d + l] * Aj[k * d + l]) < 0 ? (1u << k) : 0u;
}

// ---------------------------------------------------------------------------------------------
// K2: B^{k|k-1} of a time-propagated parent (DCE-TP), one CTA per parent.
// ---------------------------------------------------------------------------------------------
MCE_HD unsigned long long binom_u64(int n, int k) {
  if (k < 0 || k > n) return 0;
  unsigned long long res = 1;
  if (k > n - k) k = n - k;
  for (int i = 0; i < k; ++i) { res *= (unsigned long long)(n - i); res /= (unsigned long long)(i + 1); }
  return res;
}
MCE_HD void unrank_combo(long long idx, int n, int k, int* combo) {
  int x = 0;
  for (int p = 0; p < k; p++)
    for (;; x++) {
      const long long cnt = (long long)binom_u64(n - x - 1, k - p - 1);
      if (idx < cnt) { combo[p] = x; x++; break; }
      idx -= cnt;
    }
}
// PLU / solve_trf / cond('1') of cauchy_linalg.hpp:1180-1390 on a d x d system (d <= MAXD).
MCE_HD int plu_small(double* A, int* P, int n, double tol) {
  for (int j = 0; j < n; ++j) {
    double pivot = tol; int pivot_ind = -1;
    for (int i = j; i < n; ++i) if (fabs(A[i * n + j]) > fabs(pivot)) { pivot = A[i * n + j]; pivot_ind = i; }
    if (pivot_ind == -1) return 1;
    if (pivot_ind != j) for (int q = 0; q < n; q++) { const double t = A[j * n + q]; A[j * n + q] = A[pivot_ind * n + q]; A[pivot_ind * n + q] = t; }
    P[j] = pivot_ind;
    for (int k = j + 1; k < n; ++k) {
      A[k * n + j] /= A[j * n + j];
      const double temp = A[k * n + j];
      for (int q = j + 1; q < n; q++) A[k * n + q] -= temp * A[j * n + q];
    }
  }
  return 0;
}
MCE_HD void fwd_back_solve(const double* LU, double* b, int n) {
  for (int i = 0; i < n; i++) { double sol = b[i]; for (int j = 0; j < i; j++) sol -= LU[i * n + j] * b[j]; b[i] = sol; }
  for (int i = n - 1; i >= 0; i--) { double sol = b[i]; for (int j = n - 1; j > i; j--) sol -= LU[i * n + j] * b[j]; b[i] = sol / LU[i * n + i]; }
}
MCE_HD void perm_transpose(const int* P, int* P_T, int n) {
  int Preg[MAXD];
  for (int i = 0; i < n; i++) Preg[i] = i;
  for (int i = 0; i < n; i++) { const int t = Preg[i]; Preg[i] = Preg[P[i]]; Preg[P[i]] = t; }
  for (int i = 0; i < n; i++) P_T[Preg[i]] = i;
}
// Returns false when the vertex is rejected (singular or cond_1 > COND_EPS); otherwise the vertex solution.
MCE_HD bool solve_vertex(double* Ac, const double* bc, double* vertex, int n) {
  int P[MAXD], P_T[MAXD];
  double norm_val = -1;
  for (int i = 0; i < n; i++) { double v = 0; for (int j = 0; j < n; j++) v += fabs(Ac[j * n + i]); if (v > norm_val) norm_val = v; }
  if (plu_small(Ac, P, n, PLU_EPS)) return false;       // cond() returns DBL_MAX > COND_EPS
  perm_transpose(P, P_T, n);
  // 1-norm of the explicit inverse: column sums of A^{-1} = row sums of `work` before reflect_array
  double inv_norm = -1;
  for (int i = 0; i < n; i++) {
    double w[MAXD];
    for (int j = 0; j < n; j++) w[j] = 0;
    w[P_T[i]] = 1;
    fwd_back_solve(Ac, w, n);
    double v = 0;
    for (int j = 0; j < n; j++) v += fabs(w[j]);
    if (v > inv_norm) inv_norm = v;
  }
  if (norm_val * inv_norm > COND_EPS) return false;
  for (int i = 0; i < n; i++) vertex[P_T[i]] = bc[i];
  fwd_back_solve(Ac, vertex, n);
  return true;
}

// Strided variants for matrices kept in shared memory, one matrix per thread: element (i, j) lives at A[(i*n + j) * st],
// so that consecutive threads touch consecutive words (no bank conflicts, no local-memory traffic).
MCE_HD int plu_small_s(double* A, int st, int* P, int n, double tol) {
  for (int j = 0; j < n; ++j) {
    double pivot = tol; int pivot_ind = -1;
    for (int i = j; i < n; ++i) { const double v = A[(i * n + j) * st]; if (fabs(v) > fabs(pivot)) { pivot = v; pivot_ind = i; } }
    if (pivot_ind == -1) return 1;
    if (pivot_ind != j) for (int q = 0; q < n; q++) { const double t = A[(j * n + q) * st]; A[(j * n + q) * st] = A[(pivot_ind * n + q) * st]; A[(pivot_ind * n + q) * st] = t; }
    P[j] = pivot_ind;
    const double piv = A[(j * n + j) * st];
    for (int k = j + 1; k < n; ++k) {
      const double temp = A[(k * n + j) * st] / piv;
      A[(k * n + j) * st] = temp;
      for (int q = j + 1; q < n; q++) A[(k * n + q) * st] -= temp * A[(j * n + q) * st];
    }
  }
  return 0;
}
MCE_HD void fwd_back_solve_s(const double* LU, int st, double* b, int n) {
  for (int i = 0; i < n; i++) { double sol = b[i]; for (int j = 0; j < i; j++) sol -= LU[(i * n + j) * st] * b[j]; b[i] = sol; }
  for (int i = n - 1; i >= 0; i--) { double sol = b[i]; for (int j = n - 1; j > i; j--) sol -= LU[(i * n + j) * st] * b[j]; b[i] = sol / LU[(i * n + i) * st]; }
}
MCE_HD bool solve_vertex_s(double* Ac, int st, const double* bc, double* vertex, int n) {
  int P[MAXD], P_T[MAXD];
  double norm_val = -1;
  for (int i = 0; i < n; i++) { double v = 0; for (int j = 0; j < n; j++) v += fabs(Ac[(j * n + i) * st]); if (v > norm_val) norm_val = v; }
  if (plu_small_s(Ac, st, P, n, PLU_EPS)) return false;
  perm_transpose(P, P_T, n);
  double inv_norm = -1;
  for (int i = 0; i < n; i++) {
    double w[MAXD];
    for (int j = 0; j < n; j++) w[j] = 0;
    w[P_T[i]] = 1;
    fwd_back_solve_s(Ac, st, w, n);
    double v = 0;
    for (int j = 0; j < n; j++) v += fabs(w[j]);
    if (v > inv_norm) inv_norm = v;
  }
  if (norm_val * inv_norm > COND_EPS) return false;
  for (int i = 0; i < n; i++) vertex[P_T[i]] = bc[i];
  fwd_back_solve_s(Ac, st, vertex, n);
  return true;
}

struct KTpDce {
  StepParams sp; GenView gen; ParentWs ws; int vis_cap /*pow2*/, acc_cap /*pow2*/; int* diag;
  int r0 = 0;
  static MCE_HD size_t smem_bytes(int vis_cap, int acc_cap, int nthreads) {
    return sizeof(double) * MAXM * MAXD + sizeof(unsigned) * ((size_t)vis_cap + 2 * (size_t)acc_cap + 2 * (size_t)nthreads + 8);
  }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int r = r0 + c.block(), d = sp.d, gid = gen.alive[r], phc = gen_m(gen, gid), m = ws.m_tp[r], pcells = gen.cells[gid];
    const unsigned* pkeys = gen_keys(gen, gid, phc);
    unsigned* out = ws.tpB + (long long)r * ws.tpB_stride;
    if (m == phc) {                       // Gamma fully coaligned: B is unchanged (est:680-685)
      c.par([&](int tid) { for (int i = tid; i < pcells; i += c.nthreads()) out[i] = pkeys[i]; if (tid == 0) ws.tpB_cells[r] = pcells; });
      return;
    }
    if (m < d) {                          // ce:681-687: trivial keys, the cell count keeps its previous value
      c.par([&](int tid) { for (int i = tid; i < pcells; i += c.nthreads()) out[i] = (unsigned)i; if (tid == 0) ws.tpB_cells[r] = pcells; });
      return;
    }
    if (m == d) {                         // ce:454-459: no combinations exist for m <= d -> empty table (serial-path behaviour)
      c.par([&](int tid) { if (tid == 0) ws.tpB_cells[r] = 0; });
      return;
    }
    double* sA = (double*)c.smem();
    unsigned* vis = (unsigned*)(sA + MAXM * MAXD);
    unsigned* acc = vis + vis_cap;        // accepted sign vectors
    unsigned* outk = acc + acc_cap;       // emitted half
    unsigned* niv = outk + acc_cap;       // per in-flight combo: signs of the rows not in the vertex
    unsigned* cmask = niv + c.nthreads(); // per in-flight combo: bitmask of the combo rows (0 = rejected)
    int* cnt = (int*)(cmask + c.nthreads());   // [0] accepted count, [1] emitted count
    const double* Ag = ws.A + (long long)r * sp.max_shape * d;
    c.par([&](int tid) {
      for (int i = tid; i < m * d; i += c.nthreads()) sA[i] = Ag[i];
      for (int i = tid; i < vis_cap; i += c.nthreads()) vis[i] = 0xffffffffu;
      if (tid < 2) cnt[tid] = 0;
    });
    const long long ncombo = (long long)binom_u64(m, d);
    const int two_to_d = 1 << d;
    const unsigned phc_mask = (1u << phc) - 1u, top_phc = 1u << (phc - 1);
    for (long long base = 0; base < ncombo; base += c.nthreads()) {
      c.par([&](int tid) {                 // vertex of d hyperplanes with the perturbed offsets (ce:741-776)
        cmask[tid] = 0;
        const long long ci = base + tid;
        if (ci >= ncombo) return;
        int combo[MAXD]; double Ac[MAXD * MAXD], bc[MAXD], vertex[MAXD];
        unrank_combo(ci, m, d, combo);
        unsigned cm = 0;
        for (int j = 0; j < d; j++) {
          for (int l = 0; l < d; l++) Ac[j * d + l] = sA[combo[j] * d + l];
          bc[j] = sp.b_pert[combo[j]]; cm |= (1u << combo[j]);
        }
        if (!solve_vertex(Ac, bc, vertex, d)) return;
        unsigned s = 0;
        for (int ac = 0; ac < m; ac++) {
          if ((cm >> ac) & 1u) continue;
          if ((dot_lr(sA + ac * d, vertex, d) - sp.b_pert[ac]) < 0) s |= (1u << ac);
        }
        niv[tid] = s; cmask[tid] = cm;
      });
      c.par([&](int tid) {                 // encircle every vertex: 2^d sign patterns on the combo rows (ce:778-812)
        const long long items = (long long)c.nthreads() * two_to_d;
        for (long long it = tid; it < items; it += c.nthreads()) {
          const int slot = (int)(it / two_to_d), pat = (int)(it % two_to_d);
          const unsigned cm = cmask[slot];
          if (!cm) continue;
          unsigned sv = niv[slot]; int bit = 0;
          for (int row = 0; row < m; row++) if ((cm >> row) & 1u) { if ((pat >> bit) & 1) sv |= (1u << row); bit++; }
          // visited set: first inserter continues
          unsigned h = hash_u32(sv) & (unsigned)(vis_cap - 1); bool fresh = false; int probes = 0;
          for (;;) {
            const unsigned prev = c.atomic_cas(&vis[h], 0xffffffffu, sv);
            if (prev == 0xffffffffu) { fresh = true; break; }
            if (prev == sv) break;
            h = (h + 1) & (unsigned)(vis_cap - 1);
            if (++probes >= vis_cap) { c.atomic_add(diag + 1, 1); break; }
          }
          if (!fresh) continue;
          unsigned psv = sv & phc_mask;
          if (psv & top_phc) psv ^= phc_mask;
          if (key_search(pkeys, pcells, psv) < 0) continue;      // Check 1: restriction must be a parent cell
          const int o = c.atomic_add(cnt, 1);
          if (o < acc_cap) acc[o] = sv; else c.atomic_add(diag + 1, 1);
        }
      });
    }
    int nacc = c.uniform(cnt[0]);
    if (nacc > acc_cap) nacc = acc_cap;
    const int n2 = next_pow2(nacc > 1 ? nacc : 1);
    c.par([&](int tid) { for (int i = nacc + tid; i < n2; i += c.nthreads()) acc[i] = 0xffffffffu; });
    block_sort(c, acc, n2);
    const unsigned rev_m = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u), top_m = 1u << (m - 1);
    c.par([&](int tid) {                   // Check 2: keep cells whose opposite is also a cell; store the bit m-1 clear half (ce:816-852)
      for (int i = tid; i < nacc; i += c.nthreads()) {
        const unsigned b = acc[i];
        if (b & top_m) continue;
        if (key_search(acc, nacc, b ^ rev_m) >= 0) { const int o = c.atomic_add(cnt + 1, 1); outk[o] = b; }
      }
    });
    const int nout = c.uniform(cnt[1]);
    const int o2 = next_pow2(nout > 1 ? nout : 1);
    c.par([&](int tid) { for (int i = nout + tid; i < o2; i += c.nthreads()) outk[i] = 0xffffffffu; });
    block_sort(c, outk, o2);
    c.par([&](int tid) { for (int i = tid; i < nout; i += c.nthreads()) out[i] = outk[i]; if (tid == 0) ws.tpB_cells[r] = nout; });
  }
};

// ---------------------------------------------------------------------------------------------
// K7 + K8: one CTA per reduction group.
// ---------------------------------------------------------------------------------------------
struct GroupSm {          // block-uniform state, lives at the start of shared memory
  int nB, flag, cnt, owner, sigma, n_unhandled;
  // the term currently being evaluated
  int t_m, t_phc, t_pcells, t_z, t_is_child, t_has_cmap;
  unsigned t_hflag, t_enc_lhp, t_csneg;
  double t_c, t_d, t_psq;
  const cplx* t_pG;
  double q[MAXM];
  unsigned char cmap[MAXM];
};

struct KGTable {
  StepParams sp; GenView prev; GenView next; ParentWs ws; TermView tv;
  int m;                    // shape processed by this launch
  int g0;                   // first group (index within the shape) handled by block 0
  const int* order;         // [n] term indices of shape m sorted by (root, index)
  const int* grp_start;     // [n_groups + 1]
  int HC2;                  // capacity (power of two) of the shared-memory key arrays
  unsigned char* alive_flag;  // [next.n_groups]
  int* diag;
  static MCE_HD size_t smem_bytes(int HC2) {
    return ((sizeof(GroupSm) + 15) & ~(size_t)15) + (size_t)HC2 * (3 * sizeof(unsigned) + 2 * sizeof(cplx) + sizeof(unsigned long long));
  }

  // B_mu of a parent: B^{k|k-1} ^ sign(A H) mask ^ accumulated in-place re-orientations (term:245-250, flat:433-441)
  template <class Ctx> MCE_KERNEL_FN void load_parent_B(Ctx& c, int r, unsigned* dst, int* n_out) const {
    const int gid = prev.alive[r], phc = gen_m(prev, gid);
    const unsigned* src = sp.with_tp ? ws.tpB + (long long)r * ws.tpB_stride : gen_keys(prev, gid, phc);
    const int n = sp.with_tp ? ws.tpB_cells[r] : prev.cells[gid];
    const unsigned mask = ws.sgnmask[r] ^ ws.bxor[r];
    c.par([&](int tid) { for (int i = tid; i < n; i += c.nthreads()) dst[i] = src[i] ^ mask; });
    *n_out = n;
  }
  template <class Ctx> MCE_KERNEL_FN void sort_keys(Ctx& c, unsigned* arr, int n) const {
    const int n2 = next_pow2(n > 1 ? n : 1);
    c.par([&](int tid) { for (int i = n + tid; i < n2; i += c.nthreads()) arr[i] = 0xffffffffu; });
    block_sort(c, arr, n2);
  }

  // Stage the block-uniform description of term `ti` (+ its parent's sorted keys into Pk).
  template <class Ctx> MCE_KERNEL_FN void stage_term(Ctx& c, GroupSm* sm, unsigned* Pk, int ti) const {
    const long long gt = tv.t_begin[m] + ti;
    c.par([&](int tid) {
      const SlotMeta& me = tv.meta[gt];
      const int gidp = prev.alive[me.parent], phc = gen_m(prev, gidp), pc = prev.cells[gidp];
      if (tid == 0) {
        sm->t_m = m; sm->t_phc = phc; sm->t_pcells = pc; sm->t_z = me.z; sm->t_is_child = me.flags & 1; sm->t_has_cmap = (me.flags >> 1) & 1;
        sm->t_hflag = me.hflag; sm->t_enc_lhp = me.enc_lhp; sm->t_csneg = me.csneg; sm->t_c = me.c_val; sm->t_d = me.d_val;
        sm->t_pG = gen_G(prev, gidp, phc);
        const double* p = term_p(tv, m, ti);
        double s = 0; for (int i = 0; i < m; i++) s += p[i];       // sum_vec, flat:106-107
        sm->t_psq = s * s;
        sm->flag = 0;
      }
      if (tid < m) sm->q[tid] = term_q(tv, m, ti)[tid];
      if (tid < MAXM) sm->cmap[tid] = tv.cmap[gt * MAXM + tid];
      const unsigned* pk = gen_keys(prev, gidp, phc);
      for (int i = tid; i < pc; i += c.nthreads()) Pk[i] = pk[i];
    });
  }
  // G of one cell of the staged term (flat:129-227); also raises sm->flag when the cell is not negligible (flat:242-247).
  MCE_HD cplx eval_cell(GroupSm* sm, const unsigned* Pk, unsigned key) const {
    const int mm = sm->t_m;
    double ygi = 0;
    for (int k = 0; k < mm; k++) if (!((sm->t_hflag >> k) & 1u)) ygi += ((key >> k) & 1u) ? -sm->q[k] : sm->q[k];
    int lp, lm;
    parent_keys(key, mm, sm->t_phc, sm->t_z, sm->t_is_child != 0, sm->t_has_cmap ? sm->cmap : nullptr, sm->t_csneg, &lp, &lm);
    const cplx gp = g_lookup(lp ^ (int)sm->t_enc_lhp, sm->t_phc, Pk, sm->t_pG, sm->t_pcells);
    const cplx gm = g_lookup(lm ^ (int)sm->t_enc_lhp, sm->t_phc, Pk, sm->t_pG, sm->t_pcells);
    cplx g = csub(cdiv(gp, make_cplx(ygi + sm->t_d, sm->t_c)), cdiv(gm, make_cplx(ygi - sm->t_d, sm->t_c)));
    g = cscale(g, sp.gscale);
    if ((sm->t_psq * cabs_(g)) > TERM_APPROXIMATION_EPS) sm->flag = 1;
    return g;
  }
  // sigma between the rows of two terms of this shape (sets sm->sigma)
  template <class Ctx> MCE_KERNEL_FN void orient(Ctx& c, GroupSm* sm, int ti, int tj) const {
    const double* Ai = term_A(tv, m, ti, sp.d); const double* Aj = term_A(tv, m, tj, sp.d);
    c.par([&](int tid) { if (tid == 0) sm->sigma = 0; });
    c.par([&](int tid) { if (tid < m) { unsigned b = orient_bit(Ai, Aj, tid, sp.d); if (b) c.atomic_or((unsigned*)&sm->sigma, b); } });
  }

  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int d = sp.d;
    unsigned char* base = c.smem();
    GroupSm* sm = (GroupSm*)base;
    unsigned* Bk = (unsigned*)(base + ((sizeof(GroupSm) + 15) & ~(size_t)15));
    unsigned* Pk = Bk + HC2;
    unsigned* Tk = Pk + HC2;
    cplx* acc = (cplx*)(Tk + HC2);
    cplx* Gm = acc + HC2;
    unsigned long long* Sk = (unsigned long long*)(Gm + HC2);
    const int gi = g0 + c.block();
    const int start = grp_start[gi], ncomb = grp_start[gi + 1] - start;
    const int* members = order + start;
    const int gid_out = next.gid_begin[m] + gi;
    const unsigned rev_m = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u), top_m = 1u << (m - 1);

    // ---- B-table of the root (K7) ----
    const int root = members[0];
    const SlotMeta meR = tv.meta[tv.t_begin[m] + root];
    int nB = 0;
    if (meR.flags & 1) {
      if (m <= d) {                          // elementary table, ce:214-227
        nB = 1 << (m - 1);
        c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) Bk[i] = (unsigned)i; });
      } else {
        int ncp = 0;
        load_parent_B(c, meR.parent, Pk, &ncp);
        sort_keys(c, Pk, ncp);
        const int pbc = meR.pbc, z = meR.z;
        const unsigned mask_z = 1u << z, hbit = 1u << (pbc - 1), rev_pbc = (pbc >= 32) ? 0xffffffffu : ((1u << pbc) - 1u), mask_low = (1u << z) - 1u;
        c.par([&](int tid) { if (tid == 0) sm->cnt = 0; });
        c.par([&](int tid) {                 // pairs (b, b ^ 2^z) both present -> two child sign vectors (ce:261-320)
          for (int i = tid; i < ncp; i += c.nthreads()) {
            const unsigned b = Pk[i];
            unsigned bq = b ^ mask_z;
            if (bq & hbit) bq ^= rev_pbc;
            if (!(b < bq) || key_search(Pk, ncp, bq) < 0) continue;
            const unsigned z_bit = (b & mask_z) >> z;
            const unsigned csv1 = ((b >> (z + 1)) << z) | (b & mask_low) | (z_bit << (pbc - 1)), csv2 = csv1 ^ hbit;
            const int o = c.atomic_add(&sm->cnt, 2);
            Tk[o] = (csv1 & hbit) ? csv1 ^ rev_pbc : csv1;
            Tk[o + 1] = (csv2 & hbit) ? csv2 ^ rev_pbc : csv2;
          }
        });
        int nuc = c.uniform(sm->cnt);
        if (m < pbc) {                       // coalignment: keep the first bit of every class, drop duplicates (ce:323-368)
          const unsigned char* cm = tv.cmap + (tv.t_begin[m] + root) * MAXM;
          c.par([&](int tid) {
            unsigned seen = 0, sel[MAXM]; int nsel = 0;
            for (int j = 0; j < pbc; j++) { const unsigned ci = cm[j]; if (!((seen >> ci) & 1u)) { seen |= (1u << ci); sel[nsel++] = 1u << j; } }
            for (int i = tid; i < nuc; i += c.nthreads()) {
              const unsigned b = Tk[i]; unsigned bc = 0;
              for (int l = 0; l < nsel; l++) if (b & sel[l]) bc |= (1u << l);
              Tk[i] = bc;
            }
          });
          sort_keys(c, Tk, nuc);
          c.par([&](int tid) {
            if (tid != 0) return;
            int o = 0;
            for (int i = 0; i < nuc; i++) if (i == 0 || Tk[i] != Tk[i - 1]) Bk[o++] = Tk[i];
            sm->cnt = o;
          });
          nB = c.uniform(sm->cnt);
        } else {
          sort_keys(c, Tk, nuc);
          c.par([&](int tid) { for (int i = tid; i < nuc; i += c.nthreads()) Bk[i] = Tk[i]; });
          nB = nuc;
        }
      }
      c.par([&](int tid) { if (tid == 0) sm->owner = -1; });
    } else {
      load_parent_B(c, meR.parent, Bk, &nB);
      sort_keys(c, Bk, nB);
      c.par([&](int tid) { if (tid == 0) sm->owner = meR.parent; });
    }

    // ---- root table, with re-election when the candidate is negligible (flat:399-489) ----
    int k = 0, cur = root, accepted = 0;
    for (;;) {
      stage_term(c, sm, Pk, cur);
      c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) acc[i] = eval_cell(sm, Pk, Bk[i]); });
      if (c.uniform(sm->flag)) { accepted = 1; break; }
      const int lfr = cur;
      if (++k >= ncomb) break;
      cur = members[k];
      const SlotMeta meK = tv.meta[tv.t_begin[m] + cur];
      if (meK.flags & 1) {                   // new child: re-orient the group's table in place (flat:433-441)
        orient(c, sm, lfr, cur);
        unsigned sigma = (unsigned)c.uniform(sm->sigma);
        if (sigma & top_m) sigma ^= rev_m;
        if (sigma) {
          c.par([&](int tid) {
            for (int i = tid; i < nB; i += c.nthreads()) Bk[i] ^= sigma;
            if (tid == 0 && sm->owner >= 0) c.atomic_xor(ws.bxor + sm->owner, sigma);   // the table is a parent's B memory, shared with its children
          });
        }
      } else {                               // old term: its own table becomes the group's table (flat:443-473)
        load_parent_B(c, meK.parent, Bk, &nB);
        sort_keys(c, Bk, nB);
        c.par([&](int tid) { if (tid == 0) sm->owner = meK.parent; });
      }
    }
    if (!accepted) {
      c.par([&](int tid) { if (tid == 0) { alive_flag[gid_out] = 0; next.cells[gid_out] = 0; next.g_m[gid_out] = (unsigned char)m; } });
      return;
    }
    const int rsel = cur;

    // ---- remaining members: build their table and add it to the root's (flat:491-550) ----
    for (++k; k < ncomb; ++k) {
      const int t = members[k];
      const SlotMeta meT = tv.meta[tv.t_begin[m] + t];
      int own_cells = nB;
      if (!(meT.flags & 1)) {
        const int gidp = prev.alive[meT.parent];
        own_cells = sp.with_tp ? ws.tpB_cells[meT.parent] : prev.cells[gidp];
      }
      orient(c, sm, rsel, t);
      const unsigned sigma_raw = (unsigned)c.uniform(sm->sigma);
      if ((meT.flags & 1) || own_cells != nB) {
        // table = root's table re-oriented (update_btable, ce:584-625): cell i of the member is cell i of the root
        const unsigned sigma_n = (sigma_raw & top_m) ? (sigma_raw ^ rev_m) : sigma_raw;
        if (!(meT.flags & 1)) c.par([&](int tid) { if (tid == 0) c.atomic_add(diag, 1); });   // flat:516-539 also rewrites the parent's B memory: not modelled
        stage_term(c, sm, Pk, t);
        c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) Gm[i] = eval_cell(sm, Pk, Bk[i] ^ sigma_n); });
        if (c.uniform(sm->flag)) {
          const bool cj = (sigma_raw & top_m) != 0;
          c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) acc[i] = cadd(acc[i], cj ? cconj(Gm[i]) : Gm[i]); });
        }
      } else {
        // old term with its own (equal-sized) table: add by key lookup (add_gtables, flat:291-314)
        int nT = 0;
        load_parent_B(c, meT.parent, Tk, &nT);
        sort_keys(c, Tk, nT);
        stage_term(c, sm, Pk, t);
        c.par([&](int tid) { for (int i = tid; i < nT; i += c.nthreads()) Gm[i] = eval_cell(sm, Pk, Tk[i]); });
        if (c.uniform(sm->flag)) {
          c.par([&](int tid) {
            for (int i = tid; i < nB; i += c.nthreads()) {
              unsigned kq = Bk[i] ^ sigma_raw; bool cj = false;
              if (kq & top_m) { cj = true; kq ^= rev_m; }
              const int jj = key_search(Tk, nT, kq);
              if (jj >= 0) acc[i] = cadd(acc[i], cj ? cconj(Gm[jj]) : Gm[jj]);
            }
          });
        }
      }
    }

    // ---- write the surviving term: table sorted by key (flat:251-252), A, p, b (become_parent, term:748) ----
    const int n2 = next_pow2(nB > 1 ? nB : 1);
    c.par([&](int tid) {
      for (int i = tid; i < n2; i += c.nthreads()) Sk[i] = i < nB ? (((unsigned long long)Bk[i] << 32) | (unsigned)i) : ~0ull;
    });
    block_sort(c, Sk, n2);
    unsigned* ko = gen_keys(next, gid_out, m); cplx* Go = gen_G(next, gid_out, m);
    const double* As = term_A(tv, m, rsel, d); const double* ps = term_p(tv, m, rsel); const double* bs = term_b(tv, m, rsel, d);
    double* Ao = gen_A(next, gid_out, m, d); double* po = gen_p(next, gid_out, m); double* bo = gen_b(next, gid_out, d);
    c.par([&](int tid) {
      for (int i = tid; i < nB; i += c.nthreads()) { const unsigned long long s = Sk[i]; ko[i] = (unsigned)(s >> 32); Go[i] = acc[(unsigned)s]; }
      for (int i = tid; i < m * d; i += c.nthreads()) Ao[i] = As[i];
      for (int i = tid; i < m; i += c.nthreads()) po[i] = ps[i];
      for (int i = tid; i < d; i += c.nthreads()) bo[i] = bs[i];
      if (tid == 0) {
        alive_flag[gid_out] = 1; next.cells[gid_out] = nB; next.g_m[gid_out] = (unsigned char)m;
        c.atomic_add_u64((unsigned long long*)(diag + 2), (unsigned long long)nB);   // total cells of the new generation
      }
    });
  }
};

// Survivor list of the new generation: alive[rank] = gid for every group whose flag is set (ascending gid).
struct KAliveCompact {      // rank[] = exclusive scan of the flags
  int n; const unsigned char* alive_flag; const int* rank; int* alive;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int g = c.block() * c.nthreads() + tid;
      if (g >= n) return;
      if (alive_flag[g]) alive[rank[g]] = g;
    });
  }
};
struct KFlagsToInt {
  int n; const unsigned char* f; int* out;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) { const int g = c.block() * c.nthreads() + tid; if (g < n) out[g] = f[g] ? 1 : 0; });
  }
};

}  // namespace mce
#endif
