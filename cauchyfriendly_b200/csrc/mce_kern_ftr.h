// mce_kern_ftr.h -- kernels K5/K6: fast term reduction (global deduplication) and reduction groups.
// Reference loops replaced: term_reduction.hpp:30-80 (per-axis ordered maps), 89-276 (fast_term_reduction),
// cauchy_util.hpp:349-408 (ForwardFlagArray).
//
// The reference elects roots greedily in index order: i ascending, every still-unreduced j > i inside
// i's epsilon-box that passes the p / A checks gets F[j] = i.  Equivalently: j's root is the smallest
// matching i < j that is itself a root; j is a root when there is none.  That is a "lexicographically
// first" fixed point, evaluated here by rounds: a term is decided once every matching lower-index term
// below its would-be root is decided.  Results are identical to the serial loop (parity: bit-exact F).
#ifndef MCE_KERN_FTR_H_
#define MCE_KERN_FTR_H_

#include "mce_exec.h"
#include "mce_types.h"

namespace mce {

// Order-preserving image of an fp64 value (sort key for the primary search axis).
MCE_HD unsigned long long f64_sort_key(double x) {
  union { double d; unsigned long long u; } v; v.d = x;
  return (v.u & 0x8000000000000000ull) ? ~v.u : (v.u | 0x8000000000000000ull);
}

// match(i -> j): would root i absorb term j?  (term_reduction.hpp:108-157 window semantics: on every axis
// fl(b_i - eps) < b_j <= fl(b_i + eps); then |p_j - p_i|_inf <= eps (tr:223-236) and the A check exactly
// as written in tr:238-263, which compares the first m scalars of A, d times each -- quirk A.9(i).)
MCE_HD bool ftr_match(const double* bi, const double* bj, const double* pi, const double* pj, const double* Ai, const double* Aj,
                      int m, int d, const int* order) {
  const double ep = REDUCTION_EPS;
  for (int a = 0; a < d; a++) {
    const int ax = order[a];
    const double lo = bi[ax] - ep, hi = bi[ax] + ep, v = bj[ax];
    if (!(lo < v && v <= hi)) return false;
  }
  for (int k = 0; k < m; k++) if (fabs(pj[k] - pi[k]) > ep) return false;
  for (int k = 0; k < m; k++) {
    const double ar = Ai[k], ac = Aj[k];
    const bool pos = fabs(ar - ac) < ep, neg = fabs(ar + ac) < ep;
    if (!(pos || neg)) return false;
  }
  return true;
}

// sort keys of the primary axis for one shape
struct KFtrKeys {
  TermView tv; int m, d, axis0; unsigned long long* keys; int* idx; int* F;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int i = c.block() * c.nthreads() + tid;
      if (i >= tv.n[m]) return;
      keys[i] = f64_sort_key(term_b(tv, m, i, d)[axis0]);
      idx[i] = i; F[i] = -1;
    });
  }
};

// wide[i]: the root's own axis-0 window holds at least two points (the `(gti - lti) > 2` gate, tr:121)
// window bounds of term i on the sorted axis: wide flag + (sampled) window population
MCE_HD int ftr_wide_term(const TermView& tv, int m, int d, int axis0, const unsigned long long* skeys, int i, int* win) {
  const int n = tv.n[m];
  const double qp = term_b(tv, m, i, d)[axis0];
  const unsigned long long klo = f64_sort_key(qp - REDUCTION_EPS), khi = f64_sort_key(qp + REDUCTION_EPS);
  // lti = last index with value <= lo ; gti = first index with value > hi   (tr:89-106)
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (skeys[mid] <= klo) lo = mid + 1; else hi = mid; }
  const int lti = lo - 1;
  lo = 0; hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (skeys[mid] <= khi) lo = mid + 1; else hi = mid; }
  const int gti = lo;
  *win = gti - lti - 1;
  return (gti - lti) > 2;
}
struct KFtrWide {
  TermView tv; int m, d, axis0; const unsigned long long* skeys; unsigned char* wide; unsigned long long* density /* sum of window sizes */;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int i = c.block() * c.nthreads() + tid;
      if (i >= tv.n[m]) return;
      int win;
      wide[i] = (unsigned char)ftr_wide_term(tv, m, d, axis0, skeys, i, &win);
      if ((i & 15) == 0) c.atomic_add_u64(density, (unsigned long long)win);   // 1/16 sample of the window sizes
    });
  }
};

// One resolution round; `n_unknown` counts terms still undecided after the round.
// One term (sorted position `pos`) of one resolution round: returns 1 when the term stays undecided.
template <class Ctx>
MCE_KERNEL_FN int ftr_round_pos(Ctx& c, const TermView& tv, const StepParams& sp, int m, const unsigned long long* skeys, const int* sidx,
                                const unsigned char* wide, int* F, int pos) {
  const int n = tv.n[m], d = sp.d;
  const int j = sidx[pos];
  if (c.load_relaxed(F + j) != -1) return 0;
  const double* bj = term_b(tv, m, j, d); const double* pj = term_p(tv, m, j); const double* Aj = term_A(tv, m, j, d);
  const int ax0 = sp.tr_order[0];
  // conservative scan bounds around b_j on the primary axis; ftr_match applies the exact interval test
  const double slack = 4.0 * REDUCTION_EPS + 8.0 * fabs(bj[ax0]) * 2.3e-16;
  const unsigned long long klo = f64_sort_key(bj[ax0] - slack), khi = f64_sort_key(bj[ax0] + slack);
  // Ascending scan of the window.  Only the lowest-index matching candidate matters (a root there decides j, an undecided
  // term there postpones j), so candidates at or above the best index so far are skipped without touching their data;
  // the sort is stable, hence a cluster of coincident terms is visited in index order and costs one match per term.
  int min_root = 0x7fffffff, min_unknown = 0x7fffffff;
  int qs = pos;
  while (qs > 0 && skeys[qs - 1] >= klo) qs--;
  for (int q = qs; q < n; q++) {
    if (q == pos) continue;
    if (skeys[q] > khi) break;
    const int i = sidx[q];
    if (i < j && i < min_root && i < min_unknown && wide[i]) {
      const int Fi = c.load_relaxed(F + i);
      if ((Fi == i || Fi == -1) && ftr_match(term_b(tv, m, i, d), bj, term_p(tv, m, i), pj, term_A(tv, m, i, d), Aj, m, d, sp.tr_order)) {
        if (Fi == i) min_root = i; else min_unknown = i;
      }
    }
  }
  if (min_unknown < min_root) return 1;   // an undecided lower term could still claim j
  F[j] = (min_root != 0x7fffffff) ? min_root : j;
  return 0;
}
struct KFtrRound {
  TermView tv; StepParams sp; int m; const unsigned long long* skeys; const int* sidx; const unsigned char* wide; int* F; int* n_unknown;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int pos = c.block() * c.nthreads() + tid;
      if (pos >= tv.n[m]) return;
      if (ftr_round_pos(c, tv, sp, m, skeys, sidx, wide, F, pos)) c.atomic_add(n_unknown, 1);
    });
  }
};

// Tiled variant of KFtrRound: a block owns FTR_TB consecutive positions of the sorted primary axis.  The union of their
// epsilon-windows is staged chunk by chunk into shared memory (index, decision state, b, p and the first m scalars of A of
// every candidate), so each candidate is fetched from HBM once per block instead of once per (term, candidate) pair.
constexpr int FTR_TB = 128, FTR_C = 128;
struct KFtrRoundTiled {
  TermView tv; StepParams sp; int m; const unsigned long long* skeys; const int* sidx; const unsigned char* wide; int* F; int* n_unknown;
  static MCE_HD size_t smem_bytes(int m, int d) {
    return (size_t)(FTR_TB + FTR_C) * (sizeof(double) * (d + 2 * m) + sizeof(unsigned long long) + 2 * sizeof(int) + 4) + 64;
  }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int n = tv.n[m], d = sp.d, ax0 = sp.tr_order[0], W = d + 2 * m;
    const int p0 = c.block() * FTR_TB;
    unsigned char* base = c.smem();
    double* tdat = (double*)base;                              // [FTR_TB][W]  own terms: b, p, A[0..m)
    double* cdat = tdat + FTR_TB * W;                          // [FTR_C][W]   candidates
    unsigned long long* ckey = (unsigned long long*)(cdat + FTR_C * W);
    int* cidx = (int*)(ckey + FTR_C);
    int* cF = cidx + FTR_C;
    int* tmin = cF + FTR_C;                                    // [FTR_TB][2] running (min_root, min_unknown)
    int* ctl = tmin + 2 * FTR_TB;                              // [0] any undecided term in the tile, [1] lo_pos, [2] hi_pos
    unsigned char* cw = (unsigned char*)(ctl + 4);             // [FTR_C] wide flags
    const int ntile = (n - p0) < FTR_TB ? (n - p0) : FTR_TB;
    c.par([&](int tid) { if (tid == 0) ctl[0] = 0; });
    c.par([&](int tid) {
      for (int t = tid; t < ntile; t += c.nthreads()) {
        const int j = sidx[p0 + t];
        if (c.load_relaxed(F + j) == -1) ctl[0] = 1;
        tmin[2 * t] = 0x7fffffff; tmin[2 * t + 1] = 0x7fffffff;
        double* row = tdat + t * W;
        const double* bj = term_b(tv, m, j, d); const double* pj = term_p(tv, m, j); const double* Aj = term_A(tv, m, j, d);
        for (int k = 0; k < d; k++) row[k] = bj[k];
        for (int k = 0; k < m; k++) { row[d + k] = pj[k]; row[d + m + k] = Aj[k]; }
      }
      if (tid == 0) {            // union of the tile's windows on the sorted axis (conservative slack; the exact interval test is in ftr_match)
        const double vlo = term_b(tv, m, sidx[p0], d)[ax0], vhi = term_b(tv, m, sidx[p0 + ntile - 1], d)[ax0];
        const unsigned long long klo = f64_sort_key(vlo - (4.0 * REDUCTION_EPS + 8.0 * fabs(vlo) * 2.3e-16));
        const unsigned long long khi = f64_sort_key(vhi + (4.0 * REDUCTION_EPS + 8.0 * fabs(vhi) * 2.3e-16));
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (skeys[mid] < klo) lo = mid + 1; else hi = mid; }
        ctl[1] = lo;
        lo = 0; hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (skeys[mid] <= khi) lo = mid + 1; else hi = mid; }
        ctl[2] = lo;
      }
    });
    if (!ctl[0]) return;         // every term of the tile is decided (ctl[0] is not written again)
    const int lo_pos = ctl[1], hi_pos = ctl[2];
    for (int cs = lo_pos; cs < hi_pos; cs += FTR_C) {
      const int cn = (hi_pos - cs) < FTR_C ? (hi_pos - cs) : FTR_C;
      c.par([&](int tid) {        // stage a chunk of candidates
        for (int q = tid; q < cn; q += c.nthreads()) {
          const int i = sidx[cs + q];
          cidx[q] = i; ckey[q] = skeys[cs + q]; cF[q] = c.load_relaxed(F + i); cw[q] = wide[i];
          double* row = cdat + q * W;
          const double* bi = term_b(tv, m, i, d); const double* pi = term_p(tv, m, i); const double* Ai = term_A(tv, m, i, d);
          for (int k = 0; k < d; k++) row[k] = bi[k];
          for (int k = 0; k < m; k++) { row[d + k] = pi[k]; row[d + m + k] = Ai[k]; }
        }
      });
      c.par([&](int tid) {        // every undecided term of the tile against the chunk
        for (int t = tid; t < ntile; t += c.nthreads()) {
          const int j = sidx[p0 + t];
          if (c.load_relaxed(F + j) != -1) continue;
          const double* rj = tdat + t * W;
          const double slack = 4.0 * REDUCTION_EPS + 8.0 * fabs(rj[ax0]) * 2.3e-16;
          const unsigned long long klo = f64_sort_key(rj[ax0] - slack), khi = f64_sort_key(rj[ax0] + slack);
          int min_root = tmin[2 * t], min_unknown = tmin[2 * t + 1];
          for (int q = 0; q < cn; q++) {
            const int i = cidx[q];
            if (!(i < j && i < min_root && i < min_unknown) || ckey[q] < klo || ckey[q] > khi || !cw[q]) continue;   // only the lowest matching index matters
            const int Fi = cF[q];
            if (Fi != i && Fi != -1) continue;
            const double* ri = cdat + q * W;
            if (!ftr_match(ri, rj, ri + d, rj + d, ri + d + m, rj + d + m, m, d, sp.tr_order)) continue;
            if (Fi == i) min_root = i; else min_unknown = i;
          }
          tmin[2 * t] = min_root; tmin[2 * t + 1] = min_unknown;
        }
      });
    }
    c.par([&](int tid) {
      for (int t = tid; t < ntile; t += c.nthreads()) {
        const int j = sidx[p0 + t];
        if (c.load_relaxed(F + j) != -1) continue;
        const int min_root = tmin[2 * t], min_unknown = tmin[2 * t + 1];
        if (min_unknown < min_root) { c.atomic_add(n_unknown, 1); continue; }   // an undecided lower term could still claim j
        F[j] = (min_root != 0x7fffffff) ? min_root : j;
      }
    });
  }
};

// After sorting term indices by (F, index): group heads and sizes. order[] holds term indices sorted by root.
struct KGroupHeads {
  int n; const int* F; const int* order; int* is_head;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int k = c.block() * c.nthreads() + tid;
      if (k >= n) return;
      is_head[k] = (F[order[k]] == order[k]) ? 1 : 0;
    });
  }
};
struct KGroupFill {     // head_rank = exclusive scan of is_head; grp_start[rank] = position of the head in order[]
  int n; const int* is_head; const int* head_rank; int* grp_start; const int* n_groups /* device: KCountRoots */;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int k = c.block() * c.nthreads() + tid;
      if (k >= n) return;
      if (is_head[k]) grp_start[head_rank[k]] = k;
      if (k == 0) grp_start[n_groups[0]] = n;
    });
  }
};
struct KCountRoots {    // out[0] = number of roots, out[1] = number of roots that are old terms (index < n_old)
  int n, n_old; const int* F; int* out;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    int* sm = (int*)c.smem();
    c.par([&](int tid) { if (tid < 2) sm[tid] = 0; });
    c.par([&](int tid) {
      const int j = c.block() * c.nthreads() + tid;
      if (j < n && F[j] == j) { c.atomic_add(&sm[0], 1); if (j < n_old) c.atomic_add(&sm[1], 1); }
    });
    c.par([&](int tid) { if (tid < 2 && sm[tid]) c.atomic_add(out + tid, sm[tid]); });
  }
};
struct KShapeBounds {   // bounds[m] = number of surviving groups with gid < gid_begin[m] (rank = exclusive scan of the alive flags)
  GenView gen; int ngr; const int* flags; const int* rank; int* bounds;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      if (tid > NSHAPE) return;
      const int g = gen.gid_begin[tid];
      bounds[tid] = g >= ngr ? rank[ngr - 1] + flags[ngr - 1] : rank[g];
    });
  }
};
struct KRootKeys {      // sort key for grouping: (root index << 32) | term index
  int n; const int* F; unsigned long long* keys; int* vals;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int j = c.block() * c.nthreads() + tid;
      if (j >= n) return;
      keys[j] = ((unsigned long long)(unsigned)F[j] << 32) | (unsigned)j;
      vals[j] = j;
    });
  }
};

// ---------------------------------------------------------------------------------------------
// Whole term reduction of one SMALL shape (n <= FTR_SMALL_N) in one CTA: sort of the primary axis, window flags, resolution
// rounds until the fixed point, grouping sort, group heads / starts and root counts.  Replaces ~25 launches and every host
// round trip of the general path for the early steps of a window, which are launch-latency bound.  Same results: the sorts
// order by (key, index), which is what the stable radix sort of the general path produces.
// ---------------------------------------------------------------------------------------------
constexpr int FTR_SMALL_N = 4096;
struct KFtrSmall {
  TermView tv; StepParams sp; int shapes[NSHAPE];        // block b handles shape shapes[b]
  unsigned long long* k1_all; int* i1_all; unsigned char* wide_all; int* F_all; int* order_all; int* gstart_all; int* cr /*[NSHAPE][2]*/;
  static MCE_HD size_t smem_bytes() { return (sizeof(unsigned long long) + sizeof(int)) * FTR_SMALL_N + 64; }
  template <class Ctx> MCE_KERNEL_FN void sort_pairs(Ctx& c, unsigned long long* key, int* val, int n2) const {   // bitonic, ascending by (key, val)
    for (int k = 2; k <= n2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1)
        c.par([&](int tid) {
          for (int i = tid; i < n2; i += c.nthreads()) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const unsigned long long ka = key[i], kb = key[ixj]; const int va = val[i], vb = val[ixj];
              const bool gt = ka > kb || (ka == kb && va > vb);
              const bool up = (i & k) == 0;
              if (up ? gt : !gt) { key[i] = kb; key[ixj] = ka; val[i] = vb; val[ixj] = va; }
            }
          }
        });
  }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int m = shapes[c.block()], n = tv.n[m], d = sp.d;
    const long long tb = tv.t_begin[m];
    unsigned long long* skeys = k1_all + tb; int* sidx = i1_all + tb; unsigned char* wide = wide_all + tb;
    int* F = F_all + tb; int* order = order_all + tb; int* gstart = gstart_all + tb + m;
    unsigned long long* key = (unsigned long long*)c.smem();
    int* val = (int*)(key + FTR_SMALL_N);
    int* ctl = val + FTR_SMALL_N;                            // [0] undecided terms of the round, [1] roots, [2] old-term roots
    int n2 = 1; while (n2 < n) n2 <<= 1;
    // ---- primary-axis sort ----
    c.par([&](int tid) {
      for (int i = tid; i < n2; i += c.nthreads()) {
        if (i < n) { key[i] = f64_sort_key(term_b(tv, m, i, d)[sp.tr_order[0]]); val[i] = i; F[i] = -1; }
        else { key[i] = ~0ull; val[i] = 0x7fffffff; }
      }
    });
    sort_pairs(c, key, val, n2);
    c.par([&](int tid) { for (int i = tid; i < n; i += c.nthreads()) { skeys[i] = key[i]; sidx[i] = val[i]; } });
    c.par([&](int tid) { for (int i = tid; i < n; i += c.nthreads()) { int win; wide[i] = (unsigned char)ftr_wide_term(tv, m, d, sp.tr_order[0], key, i, &win); } });
    // ---- resolution rounds ----
    for (int round = 0; round <= n + 1; round++) {
      c.par([&](int tid) { if (tid == 0) ctl[0] = 0; });
      c.par([&](int tid) {
        int unk = 0;
        for (int pos = tid; pos < n; pos += c.nthreads()) unk += ftr_round_pos(c, tv, sp, m, key, val, wide, F, pos);
        if (unk) c.atomic_add(&ctl[0], unk);
      });
      if (c.uniform(ctl[0]) == 0) break;
    }
    // ---- groups: terms sorted by (root, index); heads, starts, root counts ----
    c.par([&](int tid) {
      for (int i = tid; i < n2; i += c.nthreads()) {
        if (i < n) { key[i] = ((unsigned long long)(unsigned)F[i] << 32) | (unsigned)i; val[i] = i; }
        else { key[i] = ~0ull; val[i] = 0x7fffffff; }
      }
      if (tid == 0) { ctl[1] = 0; ctl[2] = 0; }
    });
    sort_pairs(c, key, val, n2);
    int* is_head = (int*)key;                                // the keys are dead once order[] is written: reuse as int scratch
    int* head_rank = is_head + FTR_SMALL_N;
    c.par([&](int tid) { for (int i = tid; i < n; i += c.nthreads()) order[i] = val[i]; });
    c.par([&](int tid) { for (int i = tid; i < n; i += c.nthreads()) { const int t = val[i]; is_head[i] = (F[t] == t) ? 1 : 0; } });
    const int chunk = (n + c.nthreads() - 1) / c.nthreads();
    c.par([&](int tid) {                                     // exclusive scan, chunk per thread (totals go to val[], dead by now)
      const int lo = tid * chunk, hi = lo + chunk < n ? lo + chunk : n;
      int acc = 0, old = 0;
      for (int i = lo; i < hi; i++) { head_rank[i] = acc; acc += is_head[i]; if (is_head[i] && order[i] < tv.n_old[m]) old++; }
      val[tid] = acc;
      if (old) c.atomic_add(&ctl[2], old);
    });
    c.par([&](int tid) {
      if (tid != 0) return;
      int acc = 0;
      for (int t = 0; t < c.nthreads(); t++) { const int v = val[t]; val[t] = acc; acc += v; }
      ctl[1] = acc;
    });
    c.par([&](int tid) {
      const int lo = tid * chunk, hi = lo + chunk < n ? lo + chunk : n;
      const int base = val[tid];
      for (int i = lo; i < hi; i++) if (is_head[i]) gstart[head_rank[i] + base] = i;
      if (tid == 0) { gstart[ctl[1]] = n; cr[2 * m] = ctl[1]; cr[2 * m + 1] = ctl[2]; }
    });
  }
};

}  // namespace mce
#endif
