// mce_kern_part.h -- kernels of the PARTITIONED multi-GPU step: every term lives on
// exactly one rank.  A rank propagates its own parents (time propagation, DCE-TP, measurement update), then
//   1. the post-coalignment terms of every new shape are routed to the rank that owns their reduction key
//      (b[TR_SEARCH_IDXS_ORDERING[0]], the axis term_reduction.hpp:30-80 sorts on) by range splitters that are snapped
//      to gaps wider than the reduction window, so that no FTR window (tr:89-157) straddles two ranks and the
//      deduplication stays global;
//   2. the owner fetches the tables of the parents its terms descend from (one record per parent, all-to-all);
//   3. FTR, reduction groups and the G-table kernel run unchanged on the owner, over the fetched ("imported") parents;
//   4. the survivors stay where they are: they are the next step's parents of that rank.
// Canonical order (the NUM_CPUS = 1 reference's) is carried by a 64-bit ordinal per term, not by position:
//   gidx = kind << 62 | global alive rank of the parent << 6 | slot of the parent (0 = old term, 1 + t = child t)
// -- old terms before children, then generation order (cauchy_estimator.hpp:756-774).  Owned terms are sorted by gidx, so
// "lowest index" root election, member order and survivor order are the reference's.
#ifndef MCE_KERN_PART_H_
#define MCE_KERN_PART_H_

#include <string.h>

#include "mce_kern_ftr.h"

namespace mce {

constexpr int PART_MAXW = 16;             // ranks of one estimator
constexpr int PART_TB = 32;               // terms / parents per CTA of the pack and unpack kernels
constexpr unsigned long long PART_CHILD_BIT = 1ull << 62;

MCE_HD double f64_from_sort_key(unsigned long long k) {
  union { double d; unsigned long long u; } v;
  v.u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return v.d;
}
// Two consecutive sorted axis values are "far apart": no reduction window (tr:108-157: |b_i - b_j| <= 1e-8 on every axis,
// scanned with the slack of ftr_round_pos) can hold both.
MCE_HD bool part_gap(unsigned long long klo, unsigned long long khi) {
  const double a = f64_from_sort_key(klo), b = f64_from_sort_key(khi);
  return (b - a) > 16.0 * REDUCTION_EPS + 64.0 * 2.3e-16 * (fabs(a) > fabs(b) ? fabs(a) : fabs(b));
}

// ---- term records: one fixed-size record of 8-byte words per term of shape m ----
//   word 0 gidx | 1..5 SlotMeta | 6..9 coalignment map | b[d] | p[m] | q[m] | A[m*d]
MCE_HD int part_rec_words(int m, int d) { return 10 + d + 2 * m + m * d; }
static_assert(sizeof(SlotMeta) == 40, "term records carry SlotMeta as five words");

// sort keys of the primary reduction axis for every term of one shape
struct KPartKeys {
  TermView tv; int m, d, axis0; unsigned long long* keys;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int i = c.block() * c.nthreads() + tid;
      if (i < tv.n[m]) keys[i] = f64_sort_key(term_b(tv, m, i, d)[axis0]);
    });
  }
};

// Range splitters of one shape from ALL ranks' sorted keys: splitter j starts at the (j+1)/W quantile and moves down to the
// nearest position whose left neighbour is a gap away; 0 when there is none (then the ranks below receive nothing).
struct KPartSplit {
  const unsigned long long* sk; int n, W; unsigned long long* split /*[W-1]*/;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    int* found = (int*)c.smem();                       // [nthreads + 1]
    const int j = c.block(), NT = c.nthreads();
    int q = (int)((long long)n * (j + 1) / W), best = -1;
    for (int hi = q; hi > 0 && best < 0; hi -= NT) {
      c.par([&](int tid) { const int i = hi - tid; found[tid] = (i >= 1 && part_gap(sk[i - 1], sk[i])) ? i : -1; });
      c.par([&](int tid) { if (tid == 0) { int b = -1; for (int t = 0; t < NT; t++) if (found[t] > b) b = found[t]; found[NT] = b; } });
      best = c.uniform(found[NT]);
    }
    c.par([&](int tid) { if (tid == 0) split[j] = (best > 0 && best < n) ? sk[best] : 0ull; });
  }
};

// 8-byte moves between differently typed fields (records are plain words)
MCE_HD unsigned long long part_ld64(const void* p) { unsigned long long v; memcpy(&v, p, 8); return v; }
MCE_HD void part_st64(void* p, unsigned long long v) { memcpy(p, &v, 8); }

struct PartSplitters { unsigned long long s[PART_MAXW]; };
MCE_HD int part_dest(const PartSplitters& sp, int W, unsigned long long key) {
  int dst = 0;
  for (int j = 0; j < W - 1; j++) dst += (sp.s[j] <= key) ? 1 : 0;
  return dst;
}

// destination rank of every term of one shape + per-destination counts
struct KPartDest {
  const unsigned long long* keys; int n, W; const unsigned long long* split; unsigned char* dest; int* cnt /*[W]*/;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    int* hist = (int*)c.smem();                        // [PART_MAXW]
    c.par([&](int tid) { if (tid < PART_MAXW) hist[tid] = 0; });
    c.par([&](int tid) {
      const int i = c.block() * c.nthreads() + tid;
      if (i >= n) return;
      PartSplitters s; for (int j = 0; j < W - 1; j++) s.s[j] = split[j];
      const int dst = part_dest(s, W, keys[i]);
      dest[i] = (unsigned char)dst;
      c.atomic_add(&hist[dst], 1);
    });
    c.par([&](int tid) { if (tid < W && hist[tid]) c.atomic_add(&cnt[tid], hist[tid]); });
  }
};

// Packs the terms of one shape into per-destination runs of records.  Positions inside a run are handed out by an atomic
// cursor: the order of arrival does not matter, the owner sorts by gidx.
struct KPartPack {
  StepParams sp; GenView gen; SlotView sl; TermView tv; int m, rank;
  const int* gpos;                        // global alive rank of every local parent
  const long long* slot_of_term;
  const unsigned char* dest; const long long* run_off /*[W] records*/; int* cursor /*[W]*/;
  unsigned long long* out;
  static MCE_HD size_t smem_bytes() { return sizeof(long long) * PART_TB; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    long long* pos = (long long*)c.smem();
    const int d = sp.d, n = tv.n[m], t0 = c.block() * PART_TB, RW = part_rec_words(m, d);
    c.par([&](int tid) {
      if (tid >= PART_TB || t0 + tid >= n) return;
      const int dst = dest[t0 + tid];
      pos[tid] = run_off[dst] + c.atomic_add(&cursor[dst], 1);
    });
    c.par([&](int tid) {
      const int cnt = n - t0 < PART_TB ? n - t0 : PART_TB;
      for (int e = tid; e < cnt * RW; e += c.nthreads()) {
        const int k = e / RW, w = e - k * RW, t = t0 + k;
        const long long gt = tv.t_begin[m] + t;
        unsigned long long v;
        if (w == 0) {
          const SlotMeta& me = tv.meta[gt];
          const long long slot = slot_of_term[gt];
          const int ms = slot_region(sl, slot);
          const int ts = (int)((slot - sl.slot_begin[ms]) % (sl.MT[ms] + 1));
          v = ((me.flags & 1) ? PART_CHILD_BIT : 0ull) | ((unsigned long long)(unsigned)gpos[me.parent] << 6) | (unsigned long long)ts;
        } else if (w < 6) {
          SlotMeta me = tv.meta[gt];
          me.pad_ = rank | (gen_m(gen, gen.alive[me.parent]) << 8);        // home rank and shape of the parent
          v = part_ld64((const unsigned char*)&me + 8 * (w - 1));
        } else if (w < 10) v = part_ld64(tv.cmap + gt * MAXM + 8 * (w - 6));
        else {
          const double* src; int o = w - 10;
          if (o < d) src = term_b(tv, m, t, d) + o;
          else if ((o -= d) < m) src = term_p(tv, m, t) + o;
          else if ((o -= m) < m) src = term_q(tv, m, t) + o;
          else src = term_A(tv, m, t, d) + (o - m);
          v = part_ld64(src);
        }
        out[pos[k] * RW + w] = v;
      }
    });
  }
};

// gidx of every received record (sort key of the owner's canonical order)
struct KPartRecKeys {
  const unsigned long long* recs; int n, RW; unsigned long long* keys; int* idx;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int i = c.block() * c.nthreads() + tid;
      if (i < n) { keys[i] = recs[(long long)i * RW]; idx[i] = i; }
    });
  }
};
// number of old terms among the sorted gidx of one shape
struct KPartCountOld {
  const unsigned long long* sk; int n; int* out;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      if (tid != 0) return;
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sk[mid] < PART_CHILD_BIT) lo = mid + 1; else hi = mid; }
      *out = lo;
    });
  }
};
// received records -> the owner's TermView, in gidx order
struct KPartUnpack {
  int d, m; TermView tv; const unsigned long long* recs; const int* order; unsigned long long* gidx /*[global term]*/;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int n = tv.n[m], t0 = c.block() * PART_TB, RW = part_rec_words(m, d);
    c.par([&](int tid) {
      const int cnt = n - t0 < PART_TB ? n - t0 : PART_TB;
      for (int e = tid; e < cnt * RW; e += c.nthreads()) {
        const int k = e / RW, w = e - k * RW, t = t0 + k;
        const long long gt = tv.t_begin[m] + t;
        const unsigned long long v = recs[(long long)order[t] * RW + w];
        if (w == 0) gidx[gt] = v;
        else if (w < 6) part_st64((unsigned char*)&tv.meta[gt] + 8 * (w - 1), v);
        else if (w < 10) part_st64(tv.cmap + gt * MAXM + 8 * (w - 6), v);
        else {
          double* dst; int o = w - 10;
          if (o < d) dst = term_b(tv, m, t, d) + o;
          else if ((o -= d) < m) dst = term_p(tv, m, t) + o;
          else if ((o -= m) < m) dst = term_q(tv, m, t) + o;
          else dst = term_A(tv, m, t, d) + (o - m);
          part_st64(dst, v);
        }
      }
    });
  }
};

// ---- parent import: which parents do my terms descend from? ----
// import key = parent shape << 40 | home rank << 32 | alive rank on the home rank
MCE_HD unsigned long long part_import_key(int phc, int home, int r) { return ((unsigned long long)phc << 40) | ((unsigned long long)home << 32) | (unsigned)r; }
struct KImportKeys {
  const SlotMeta* meta; long long n; unsigned long long* keys; int* idx;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const long long i = (long long)c.block() * c.nthreads() + tid;
      if (i >= n) return;
      const SlotMeta& me = meta[i];
      keys[i] = part_import_key((me.pad_ >> 8) & 0xff, me.pad_ & 0xff, me.parent); idx[i] = (int)i;
    });
  }
};
struct KUniqFlags {
  const unsigned long long* sk; long long n; int* flag;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const long long i = (long long)c.block() * c.nthreads() + tid;
      if (i < n) flag[i] = (i == 0 || sk[i] != sk[i - 1]) ? 1 : 0;
    });
  }
};
// every term learns the import index of its parent; the import list (sorted by shape, home rank, alive rank) is written
struct KImportAssign {
  const unsigned long long* sk; const int* sidx; const int* flag; const int* pos; long long n;
  SlotMeta* meta; unsigned long long* ilist; int* rlist; int* n_import;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const long long i = (long long)c.block() * c.nthreads() + tid;
      if (i >= n) return;
      const int imp = pos[i] + flag[i] - 1;
      meta[sidx[i]].parent = imp;
      if (flag[i]) { ilist[imp] = sk[i]; rlist[imp] = (int)(unsigned)(sk[i] & 0xffffffffull); }
      if (i == n - 1) *n_import = imp + 1;
    });
  }
};
// seg[s * W + h] = first import index of (parent shape s, home rank h); seg[NSHAPE * W] = number of imports
struct KImportSegs {
  const unsigned long long* ilist; const int* n_import; int W; int* seg;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int e = c.block() * c.nthreads() + tid, n = *n_import;
      if (e > NSHAPE * W) return;
      if (e == NSHAPE * W) { seg[e] = n; return; }
      const unsigned long long key = part_import_key(e / W, e % W, 0);
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ilist[mid] < key) lo = mid + 1; else hi = mid; }
      seg[e] = lo;
    });
  }
};

// ---- parent records (one per requested parent of shape phc), 16-byte aligned ----
//   G[S] | header (cells, tp cells, sgnmask, global alive rank; gmax; m_tp, pad) | keys[S] | rbm[rw] | tpB[tpS] | rpf[rw]
struct ParentRecLayout { int S, rw, tpS, o_hdr, o_keys, o_rbm, o_tpb, o_rpf, bytes; };
MCE_HD ParentRecLayout parent_rec_layout(int phc, int d, int tpS) {
  ParentRecLayout L; L.S = cell_count_central_half(phc, d); L.rw = rank_words(phc); L.tpS = tpS;
  L.o_hdr = 16 * L.S; L.o_keys = L.o_hdr + 32; L.o_rbm = L.o_keys + 4 * L.S; L.o_tpb = L.o_rbm + 4 * L.rw; L.o_rpf = L.o_tpb + 4 * tpS;
  L.bytes = (L.o_rpf + 2 * L.rw + 15) & ~15;
  return L;
}
struct ParentRecHdr { int cells, tp_cells; unsigned sgnmask; int gpos; double gmax; int m_tp, pad; };
static_assert(sizeof(ParentRecHdr) == 32, "parent record header");

struct KImportPack {                      // home rank: one CTA per requested parent
  GenView gen; ParentWs ws; int with_tp, phc, d; ParentRecLayout L; const int* req; const int* gpos; unsigned char* out;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int r = req[c.block()], gid = gen.alive[r], cells = gen.cells[gid];
    unsigned char* rec = out + (long long)c.block() * L.bytes;
    const cplx* G = gen_G(gen, gid, phc); const unsigned* keys = gen_keys(gen, gid, phc);
    const long long rk = gen_rk_off(gen, gid, phc);
    const int tpc = with_tp ? ws.tpB_cells[r] : 0;
    c.par([&](int tid) {
      cplx* oG = (cplx*)rec; unsigned* ok = (unsigned*)(rec + L.o_keys); unsigned* ob = (unsigned*)(rec + L.o_rbm);
      unsigned* ot = (unsigned*)(rec + L.o_tpb); unsigned short* of = (unsigned short*)(rec + L.o_rpf);
      for (int i = tid; i < cells; i += c.nthreads()) { oG[i] = G[i]; ok[i] = keys[i]; }
      for (int i = tid; i < L.rw; i += c.nthreads()) { ob[i] = gen.rbm[rk + i]; of[i] = gen.rpf[rk + i]; }
      if (with_tp) { const unsigned* tb = ws.tpB + (long long)r * ws.tpB_stride; for (int i = tid; i < tpc && i < L.tpS; i += c.nthreads()) ot[i] = tb[i]; }
      if (tid == 0) {
        ParentRecHdr h; h.cells = cells; h.tp_cells = tpc; h.sgnmask = ws.sgnmask[r]; h.gpos = gpos[r]; h.gmax = gen.gmax[gid];
        h.m_tp = with_tp ? ws.m_tp[r] : phc; h.pad = 0;
        *(ParentRecHdr*)(rec + L.o_hdr) = h;
      }
    });
  }
};
struct KImportUnpack {                    // owner: records of shape phc -> the import store (a GenView of its own) + ParentWs
  GenView imp; ParentWs iws; int with_tp, phc, d; ParentRecLayout L; const unsigned char* recs; int first /* import index of the first record */; int* igpos;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int i = first + c.block(), gid = i;      // import index == gid == alive rank in the import store
    const unsigned char* rec = recs + (long long)c.block() * L.bytes;
    const ParentRecHdr h = *(const ParentRecHdr*)(rec + L.o_hdr);
    cplx* G = gen_G(imp, gid, phc); unsigned* keys = gen_keys(imp, gid, phc);
    const long long rk = gen_rk_off(imp, gid, phc);
    c.par([&](int tid) {
      const cplx* iG = (const cplx*)rec; const unsigned* ik = (const unsigned*)(rec + L.o_keys); const unsigned* ib = (const unsigned*)(rec + L.o_rbm);
      const unsigned* it = (const unsigned*)(rec + L.o_tpb); const unsigned short* ifp = (const unsigned short*)(rec + L.o_rpf);
      for (int k = tid; k < h.cells; k += c.nthreads()) { G[k] = iG[k]; keys[k] = ik[k]; }
      for (int k = tid; k < L.rw; k += c.nthreads()) { imp.rbm[rk + k] = ib[k]; imp.rpf[rk + k] = ifp[k]; }
      if (with_tp) { unsigned* tb = iws.tpB + (long long)i * iws.tpB_stride; for (int k = tid; k < h.tp_cells && k < L.tpS; k += c.nthreads()) tb[k] = it[k]; }
      if (tid == 0) {
        imp.g_m[gid] = (unsigned char)phc; imp.cells[gid] = h.cells; imp.alive[i] = gid; imp.gmax[gid] = h.gmax;
        iws.sgnmask[i] = h.sgnmask; iws.bxor[i] = 0; igpos[i] = h.gpos;
        if (with_tp) iws.tpB_cells[i] = h.tp_cells;
      }
    });
  }
};

// ---- in-place re-orientation masks (flattening.hpp:433-441) across ranks: the group that holds a parent's old term is the
// only writer, every rank that imported the parent reads the mask in the second G-table phase ----
struct KBxorScatter {
  int n; const unsigned* bxor; const int* igpos; unsigned* global;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) { const int i = c.block() * c.nthreads() + tid; if (i < n && bxor[i]) global[igpos[i]] = bxor[i]; });
  }
};
struct KBxorGather {
  int n; unsigned* bxor; const int* igpos; const unsigned* global;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) { const int i = c.block() * c.nthreads() + tid; if (i < n) bxor[i] = global[igpos[i]]; });
  }
};

// ---- survivors: canonical sort key (gidx of the group's first member) and global alive rank ----
struct KSurvKeys {
  GenView next; TermView tv; int n_surv; const int* order_all; const int* gstart_all; const unsigned long long* gidx; unsigned long long* skey;
  int gstart_off[NSHAPE];
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int a = c.block() * c.nthreads() + tid;
      if (a >= n_surv) return;
      const int gid = next.alive[a], m = gen_m(next, gid), gi = gid - next.gid_begin[m];
      const int first = order_all[tv.t_begin[m] + gstart_all[gstart_off[m] + gi]];
      skey[a] = gidx[tv.t_begin[m] + first];
    });
  }
};
struct PartSurvLayout { long long off[PART_MAXW][NSHAPE]; int cnt[PART_MAXW][NSHAPE]; int shape_base[NSHAPE]; };
struct KSurvRank {                        // gpos[a] = survivors of lower shapes + survivors of the same shape with a smaller key, over all ranks
  GenView next; int n_surv, W; const unsigned long long* skey; const unsigned long long* all; int* gpos; PartSurvLayout L;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int a = c.block() * c.nthreads() + tid;
      if (a >= n_surv) return;
      const int m = gen_m(next, next.alive[a]);
      const unsigned long long key = skey[a];
      int rank = L.shape_base[m];
      for (int h = 0; h < W; h++) {
        const unsigned long long* lst = all + L.off[h][m];
        int lo = 0, hi = L.cnt[h][m];
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (lst[mid] < key) lo = mid + 1; else hi = mid; }
        rank += lo;
      }
      gpos[a] = rank;
    });
  }
};

// ---- ordered moments: every slot's (g, y) goes to its canonical position of the global slot list ----
struct KSlotScatter {
  SlotView sl; int d; const int* gpos; cplx* g_out; double* y_out;      // y_out == nullptr: only g
  double* re_out;                                                        // != nullptr: only Re g, packed (all the exact scan of Re fz needs: half the bytes to combine)
  long long gslot_begin[NSHAPE]; int gshape_base[NSHAPE];
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const long long s = (long long)c.block() * c.nthreads() + tid;
      if (s >= sl.n_slots) return;
      const int ms = slot_region(sl, s), per = sl.MT[ms] + 1;
      const long long ls = s - sl.slot_begin[ms];
      const int r = sl.par_begin[ms] + (int)(ls / per), t = (int)(ls % per);
      const long long gs = gslot_begin[ms] + (long long)(gpos[r] - gshape_base[ms]) * per + t;
      if (re_out) { re_out[gs] = sl.g[s].re; return; }
      g_out[gs] = sl.g[s];
      if (y_out) for (int i = 0; i < 2 * d; i++) y_out[gs * 2 * d + i] = sl.y[s * 2 * d + i];
    });
  }
};
struct KIota { int n; int* out; template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const { c.par([&](int tid) { const int i = c.block() * c.nthreads() + tid; if (i < n) out[i] = i; }); } };

}  // namespace mce
#endif
