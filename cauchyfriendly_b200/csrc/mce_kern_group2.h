// mce_kern_group2.h -- K7 + K8 for max_shape <= 16: the same algorithm as KGTable (mce_kern_group.h), with every
// sign-vector set held as a BITMAP over the 2^m possible keys in shared memory:
//   * membership tests of the DCE-MU pair search are single bit tests (replaces the B_mu hash of
//     cell_enumeration.hpp:237-259);
//   * duplicates of coaligned child keys vanish by construction (replaces B_coal_hash, ce:326-366);
//   * the rank of a key (prefix popcount) IS its position in the sorted table, so the parent G-table lookup
//     (binsearch, gtable.hpp:283-301: 46 % of the reference's CPU time) is two shared-memory reads + a popcount,
//     and the final table is emitted key-sorted without sorting (replaces the qsort of flattening.hpp:251-252).
// No sort and no hash probe is left in the kernel; what remains per cell is the fp64 arithmetic of flat:129-247
// (two complex divisions in libgcc's algorithm, one scale, |G| only while the term still looks negligible).
#ifndef MCE_KERN_GROUP2_H_
#define MCE_KERN_GROUP2_H_

#include "mce_kern_group.h"

namespace mce {

// Vertex of d hyperplanes with the reference's acceptance tests (cell_enumeration.hpp:741-776: PLU with partial pivoting, reject when
// singular or cond_1 > 1e12 from the explicit inverse) for a COMPILE-TIME dimension: every loop is unrolled and every array index is
// static, so the d x d system, the permutation and the right-hand sides live in registers (no shared-memory matrix, no local memory) and
// the d independent columns of the inverse give the scheduler instruction-level parallelism.  Same operations in the same order as
// solve_vertex / solve_vertex_s (mce_kern_group.h), which stay the readable restatement and serve the max_shape > 16 kernel.
#if defined(__CUDA_ARCH__)
#define MCE_UNROLL _Pragma("unroll")
#else
#define MCE_UNROLL
#endif
template <int D>
MCE_HD bool solve_vertex_r(const double* sA, const int* combo, const double* b_pert, double* vertex) {
  double M[D][D], bc[D];
  MCE_UNROLL for (int j = 0; j < D; j++) {
    const double* row = sA + combo[j] * D;
    MCE_UNROLL for (int l = 0; l < D; l++) M[j][l] = row[l];
    bc[j] = b_pert[combo[j]];
  }
  double norm_val = -1;
  MCE_UNROLL for (int i = 0; i < D; i++) { double v = 0; MCE_UNROLL for (int j = 0; j < D; j++) v += fabs(M[j][i]); if (v > norm_val) norm_val = v; }
  int P[D];
  MCE_UNROLL for (int j = 0; j < D; j++) {
    double pivot = PLU_EPS; int pi = -1;
    MCE_UNROLL for (int i = j; i < D; i++) { const double v = M[i][j]; if (fabs(v) > fabs(pivot)) { pivot = v; pi = i; } }
    if (pi == -1) return false;
    MCE_UNROLL for (int i = j + 1; i < D; i++)
      if (i == pi) { MCE_UNROLL for (int q = 0; q < D; q++) { const double t = M[j][q]; M[j][q] = M[i][q]; M[i][q] = t; } }
    P[j] = pi;
    const double piv = M[j][j];
    MCE_UNROLL for (int k = j + 1; k < D; k++) {
      const double temp = M[k][j] / piv;
      M[k][j] = temp;
      MCE_UNROLL for (int q = j + 1; q < D; q++) M[k][q] -= temp * M[j][q];
    }
  }
  int Preg[D], P_T[D];                               // perm_transpose with static indices
  MCE_UNROLL for (int i = 0; i < D; i++) Preg[i] = i;
  MCE_UNROLL for (int i = 0; i < D; i++) {
    MCE_UNROLL for (int k = 0; k < D; k++) if (k == P[i] && k != i) { const int t = Preg[i]; Preg[i] = Preg[k]; Preg[k] = t; }
  }
  MCE_UNROLL for (int i = 0; i < D; i++) { MCE_UNROLL for (int k = 0; k < D; k++) if (k == Preg[i]) P_T[k] = i; }
  double inv_norm = -1;
  MCE_NOUNROLL for (int c = 0; c <= D; c++) {        // columns of the inverse, then the right-hand side: one solve body
    if (c == D && norm_val * inv_norm > COND_EPS) return false;
    double w[D];
    MCE_UNROLL for (int k = 0; k < D; k++) {
      double v = 0;
      MCE_UNROLL for (int i = 0; i < D; i++) if (P_T[i] == k) v = (c == D) ? bc[i] : (i == c ? 1.0 : 0.0);
      w[k] = v;
    }
    MCE_UNROLL for (int i = 0; i < D; i++) { double sol = w[i]; MCE_UNROLL for (int j = 0; j < i; j++) sol -= M[i][j] * w[j]; w[i] = sol; }
    MCE_UNROLL for (int i = D - 1; i >= 0; i--) { double sol = w[i]; MCE_UNROLL for (int j = D - 1; j > i; j--) sol -= M[i][j] * w[j]; w[i] = sol / M[i][i]; }
    if (c < D) { double v = 0; MCE_UNROLL for (int j = 0; j < D; j++) v += fabs(w[j]); if (v > inv_norm) inv_norm = v; }
    else { MCE_UNROLL for (int j = 0; j < D; j++) vertex[j] = w[j]; }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// K2 for max_shape <= 16: DCE-TP with bitmaps.  The visited set F[2^m] of cell_enumeration.hpp:735 is a 2^m-bit
// bitmap (atomicOr tells the first visitor), "restriction is a parent cell" (ce:790-798) is a bit test in the
// parent's key bitmap, "the opposite was accepted too" (ce:816-852) a bit test in the accepted bitmap, and the
// surviving half is enumerated in ascending key order from that bitmap -- no hash table, no sort.
// ---------------------------------------------------------------------------------------------
template <int D>
struct KTpDce2T {
  // the vertex systems live in registers (solve_vertex_r: 2 D^2 registers for the matrix alone).  Measured at D = 7 on the LEO7 window (MU 10, 9.9 M vertices):
  // 2 / 3 / 4 / 5 / 6 CTAs per SM = 5.5 / 4.4 / 4.0 / 4.4 / 4.6 ms -- occupancy beats a spill-free matrix -- against 5.3 ms for the shared-memory solver
  static constexpr int kMaxThreads = 128, kMinBlocks = D <= 4 ? 8 : (D <= 7 ? 4 : 3);
  StepParams sp; GenView gen; ParentWs ws; int NW /* 2^max_shape / 32 */; int* diag;
  int r0 = 0;                           // first parent of this launch
  static MCE_HD size_t smem_bytes(int NW, int nthreads, int) {
    return sizeof(double) * (MAXM * MAXD) + sizeof(unsigned) * (3 * (size_t)NW + 2 * (size_t)nthreads + 8) + sizeof(unsigned short) * ((size_t)NW + 16 + 1024);
  }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    constexpr int d = D;
    const int r = r0 + c.block(), gid = gen.alive[r], phc = gen_m(gen, gid), m = ws.m_tp[r], pcells = gen.cells[gid];
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
    unsigned* bmVis = (unsigned*)(sA + MAXM * MAXD);     // visited sign vectors (m bits)
    unsigned* bmAcc = bmVis + NW;                        // accepted sign vectors
    unsigned* bmPar = bmAcc + NW;                        // parent keys (phc bits)
    unsigned* niv = bmPar + NW;
    unsigned* cmask = niv + c.nthreads();
    int* cnt = (int*)(cmask + c.nthreads());
    unsigned short* pf = (unsigned short*)(cnt + 8);
    const int nwM = m >= 5 ? (1 << (m - 5)) : 1, nwP = phc >= 5 ? (1 << (phc - 5)) : 1;
    const double* Ag = ws.A + (long long)r * sp.max_shape * d;
    c.par([&](int tid) {
      for (int i = tid; i < m * d; i += c.nthreads()) sA[i] = Ag[i];
      for (int i = tid; i < nwM; i += c.nthreads()) { bmVis[i] = 0; bmAcc[i] = 0; }
      for (int i = tid; i < nwP; i += c.nthreads()) bmPar[i] = 0;
    });
    c.par([&](int tid) { for (int i = tid; i < pcells; i += c.nthreads()) { const unsigned k = pkeys[i]; c.atomic_or(&bmPar[k >> 5], 1u << (k & 31)); } });
    const long long ncombo = (long long)binom_u64(m, d);
    const unsigned phc_mask = (1u << phc) - 1u, top_phc = 1u << (phc - 1);
    const int NT = c.nthreads();
    const int LB = d < 4 ? d : 4, per_slot = 1 << (d - LB);           // a thread walks 2^LB sign patterns of one vertex in Gray-code order
    for (long long base = 0; base < ncombo; base += NT) {
      c.par([&](int tid) {                 // vertex of d hyperplanes with the perturbed offsets (ce:741-776)
        cmask[tid] = 0;
        const long long ci = base + tid;
        if (ci >= ncombo) return;
        int combo[D]; double vertex[D];
        {                                  // unrank_combo with static indices
          long long idx = ci; int x = 0;
          MCE_UNROLL for (int i = 0; i < D; i++) {
            for (;; x++) { const long long cnt = (long long)binom_u64(m - 1 - x, D - 1 - i); if (idx < cnt) break; idx -= cnt; }
            combo[i] = x++;
          }
        }
        unsigned cm = 0;
        MCE_UNROLL for (int j = 0; j < D; j++) cm |= (1u << combo[j]);
        if (!solve_vertex_r<D>(sA, combo, sp.b_pert, vertex)) return;
        unsigned sgn = 0;
        for (int ac = 0; ac < m; ac++) {
          if ((cm >> ac) & 1u) continue;
          const double* row = sA + ac * D;
          double dot = 0;
          MCE_UNROLL for (int l = 0; l < D; l++) dot += row[l] * vertex[l];
          if ((dot - sp.b_pert[ac]) < 0) sgn |= (1u << ac);
        }
        niv[tid] = sgn; cmask[tid] = cm;
      });
      c.par([&](int tid) {                 // encircle every vertex: 2^d sign patterns on the combo rows (ce:778-812)
        const int items = NT * per_slot;
        MCE_NOUNROLL for (int it = tid; it < items; it += NT) {
          const int slot = it / per_slot, blk = it - slot * per_slot;
          const unsigned cm = cmask[slot];
          if (!cm) continue;
          unsigned long long rowpos = 0; unsigned rest = cm;            // 5-bit positions of the combo rows, ascending
          for (int b = 0; rest; b++) { const int row = MCE_FFS(rest); rest &= rest - 1u; rowpos |= (unsigned long long)row << (5 * b); }
          const unsigned g0 = (unsigned)blk << LB, gray0 = g0 ^ (g0 >> 1);
          unsigned sv = niv[slot];
          for (int b = 0; b < d; b++) if ((gray0 >> b) & 1u) sv |= 1u << (unsigned)((rowpos >> (5 * b)) & 31u);
          MCE_NOUNROLL for (int i = 0; i < (1 << LB); i++) {
            if (i) { const int b = MCE_FFS((unsigned)i); sv ^= 1u << (unsigned)((rowpos >> (5 * b)) & 31u); }   // Gray code: one row flips
            const unsigned vbit = 1u << (sv & 31);
            if (bmVis[sv >> 5] & vbit) continue;                       // cheap pre-test
            if (c.atomic_or(&bmVis[sv >> 5], vbit) & vbit) continue;   // somebody else was first
            unsigned psv = sv & phc_mask;
            if (psv & top_phc) psv ^= phc_mask;
            if (!((bmPar[psv >> 5] >> (psv & 31)) & 1u)) continue;     // Check 1: restriction must be a parent cell
            c.atomic_or(&bmAcc[sv >> 5], vbit);
          }
        }
      });
    }
    // Check 2 (ce:816-852): keep cells whose opposite was accepted too, store the half with bit m-1 clear.
    const unsigned rev_m = (1u << m) - 1u;
    const int nwH = m >= 6 ? (1 << (m - 6)) : 1;                     // words holding keys with bit m-1 clear
    c.par([&](int tid) {
      for (int w = tid; w < nwH; w += c.nthreads()) {
        unsigned bits = bmAcc[w], keep = 0;
        if (m < 6) bits &= (1u << (1 << (m - 1))) - 1u;              // tiny arrangements: lower half of the single word
        while (bits) {
          const int b = MCE_FFS(bits); bits &= bits - 1u;
          const unsigned key = (unsigned)(w * 32 + b), opp = key ^ rev_m;
          if ((bmAcc[opp >> 5] >> (opp & 31)) & 1u) keep |= 1u << b;
        }
        bmVis[w] = keep;                                             // bmVis is free now: reuse it for the surviving half
      }
    });
    // prefix popcounts + ordered enumeration
    if (nwH <= 64) {
      c.par([&](int tid) {
        if (tid >= nwH) return;
        int sacc = 0;
        for (int i = 0; i < tid; i++) sacc += MCE_POPC(bmVis[i]);
        pf[tid] = (unsigned short)sacc;
        if (tid == nwH - 1) cnt[0] = sacc + MCE_POPC(bmVis[tid]);
      });
    } else {
      c.par([&](int tid) {
        if (tid != 0) return;
        int sacc = 0;
        for (int i = 0; i < nwH; i++) { pf[i] = (unsigned short)sacc; sacc += MCE_POPC(bmVis[i]); }
        cnt[0] = sacc;
      });
    }
    c.par([&](int tid) {
      for (int w = tid; w < nwH; w += c.nthreads()) {
        unsigned bits = bmVis[w]; int o = pf[w];
        while (bits) { const int b = MCE_FFS(bits); bits &= bits - 1u; out[o++] = (unsigned)(w * 32 + b); }
      }
      if (tid == 0) ws.tpB_cells[r] = cnt[0];
    });
  }
};

constexpr int G2_CHUNK = 16;      // members staged per chunk
constexpr int G2_M = 16;          // this kernel serves max_shape <= 16
constexpr int G2_SLOTS = 3;       // staged q / orientation bits (the lean variant keeps three members in flight: pending, current, next)

// The "member is not negligible" flag is raised by whichever thread finds a qualifying cell and only ever read to skip a test:
// an intended benign race (every writer stores 1).  Building with -DMCE_RACECHECK turns both sides into atomics so that
// compute-sanitizer's racecheck reports everything else (tools/sanitize_pass.py).
#if defined(MCE_RACECHECK) && defined(__CUDA_ARCH__)
#define G2_FLAG_READ(f) atomicOr((f), 0)
#define G2_FLAG_SET(f) atomicExch((f), 1)
#else
#define G2_FLAG_READ(f) (*(volatile int*)(f))
#define G2_FLAG_SET(f) (*(f) = 1)
#endif

struct Group2Member {              // everything the kernel needs to know about one member, gathered in one parallel phase
  int ti, parent, gidp, phc, pc, own_cells;
  long long rk_off;                // parent's rank structure inside prev.rbm / prev.rpf
  const cplx* pG;                  // the parent's table and rank structure, resolved once per member (eval_cell runs per cell)
  const unsigned* bmP; const unsigned short* pfP;
  unsigned hflag, enc_lhp, csneg, mask, kflip;
  unsigned char z, is_child, has_cmap, pbc;
  unsigned char skip;              // certified negligible from the parent's largest |G| alone: no cell needs evaluating
  double c, d, psq;
  unsigned char ksrc[G2_M];
};

struct Group2Sm {
  int cnt, owner, pad0, pad1;
  unsigned sgbits[G2_SLOTS][G2_M];        // per-row orientation bits of a staged member (update_btable's sigma); slot = member index & 1
  int flag[G2_CHUNK];              // per member: some cell is not negligible
  double q[G2_SLOTS][G2_M];
  Group2Member mem[G2_CHUNK];
};

// x with its sign bit xor-ed by `neg` (0 or 1): -x is exact, so this equals `neg ? -x : x`
MCE_HD double flip_sign(double x, unsigned neg) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(x) ^ (int)(neg << 31), __double2loint(x));
#else
  union { double d; unsigned long long u; } v; v.d = x; v.u ^= (unsigned long long)neg << 63; return v.d;
#endif
}

// exclusive prefix popcounts of a bitmap; *total (may be null) receives the number of set bits
template <class Ctx> MCE_KERNEL_FN MCE_NOINLINE void bm_prefix_any(Ctx& c, const unsigned* bm, unsigned short* pf, int nw, int* total) {
  if (nw <= 64 && nw <= c.nthreads()) {               // one phase: word w sums the popcounts below it
    c.par([&](int tid) {
      if (tid >= nw) return;
      int s = 0;
      for (int i = 0; i < tid; i++) s += MCE_POPC(bm[i]);
      pf[tid] = (unsigned short)s;
      if (tid == nw - 1 && total) *total = s + MCE_POPC(bm[tid]);
    });
    return;
  }
  const int nt = c.nthreads(), chunk = (nw + nt - 1) / nt;
  c.par([&](int tid) {
    const int lo = tid * chunk, hi = lo + chunk < nw ? lo + chunk : nw;
    int s = 0;
    for (int i = lo; i < hi; i++) { pf[i] = (unsigned short)s; s += MCE_POPC(bm[i]); }
    if (lo < nw) pf[nw + tid] = (unsigned short)s;       // chunk totals live behind the prefix array
  });
  c.par([&](int tid) {
    if (tid != 0) return;
    int acc = 0;
    const int nchunks = (nw + chunk - 1) / chunk;
    for (int k = 0; k < nchunks; k++) { const int v = pf[nw + k]; pf[nw + k] = (unsigned short)acc; acc += v; }
    if (total) *total = acc;
  });
  c.par([&](int tid) {
    const int lo = tid * chunk, hi = lo + chunk < nw ? lo + chunk : nw;
    if (lo >= nw) return;
    const unsigned short base = pf[nw + tid];
    if (base) for (int i = lo; i < hi; i++) pf[i] = (unsigned short)(pf[i] + base);
  });
}


// Rank structure of every parent table (GenView::rbm / rpf): one CTA per surviving parent.
struct KBuildRank {
  GenView gen;
  static MCE_HD size_t smem_bytes(int nw_max, int nthreads) { return sizeof(unsigned) * (size_t)nw_max + sizeof(unsigned short) * ((size_t)nw_max + nthreads + 16) + 16 + sizeof(double) * nthreads; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int gid = gen.alive[c.block()], m = gen_m(gen, gid), nw = rank_words(m), cells = gen.cells[gid];
    const unsigned* keys = gen_keys(gen, gid, m);
    const cplx* G = gen_G(gen, gid, m);
    const long long off = gen_rk_off(gen, gid, m);
    unsigned* sbm = (unsigned*)c.smem();
    unsigned short* spf = (unsigned short*)(sbm + nw);
    double* smx = (double*)(((uintptr_t)(spf + nw + c.nthreads() + 16) + 15) & ~(uintptr_t)15);
    c.par([&](int tid) { for (int i = tid; i < nw; i += c.nthreads()) sbm[i] = 0; });
    c.par([&](int tid) {
      double mx = 0;              // largest |re| + |im| of the table: an upper bound of every |G| (see Group2Member::skip)
      for (int i = tid; i < cells; i += c.nthreads()) {
        const unsigned k = keys[i]; c.atomic_or(&sbm[k >> 5], 1u << (k & 31));
        const cplx g = G[i]; const double a = fabs(g.re) + fabs(g.im);
        if (!(a <= mx)) mx = a;   // NaN propagates: a NaN bound never certifies anything
      }
      smx[tid] = mx;
    });
    bm_prefix_any(c, sbm, spf, nw, (int*)nullptr);
    c.par([&](int tid) {
      for (int i = tid; i < nw; i += c.nthreads()) { gen.rbm[off + i] = sbm[i]; gen.rpf[off + i] = spf[i]; }
      if (tid == 0) { double mx = 0; for (int t = 0; t < c.nthreads(); t++) if (!(smx[t] <= mx)) mx = smx[t]; gen.gmax[gid] = mx; }
    });
  }
};

// Reduction groups with more than BIG_T members are split over several CTAs: one CTA elects the root (G2_BIG_ROOT), CTAs
// of BIG_PART members each store their members' addends (G2_BIG_PARTS), one CTA adds them in member order (G2_BIG_FINAL).
// A group of thousands of members (LTI problems: many children coincide) would otherwise serialise in one CTA.
constexpr int BIG_T = 192, BIG_PART = 48;
enum { G2_NORMAL = 0, G2_BIG_ROOT = 1, G2_BIG_PARTS = 2, G2_BIG_FINAL = 3 };
struct BigGroup { int gi, ncomb, nparts, part_base; long long rows_off, flags_off, keys_off; int hdr[4]; /* nB, root term, accepted, accepted candidate */ };
struct BigPart { int slot, part; };
struct BigArgs { BigGroup* groups; const BigPart* parts; cplx* rows; int* flags; unsigned* keys; int row_stride; };

// "member has no cell here" marker inside a stored addend row (adding +0.0 instead would not be exact for a -0.0 sum)
MCE_HD cplx skip_addend() {
  union { double d; unsigned long long u; } v; v.u = 0x7ff8dead0000beefULL;
  return make_cplx(v.d, 0.0);
}
MCE_HD bool is_skip_addend(const cplx& x) {
  union { double d; unsigned long long u; } v; v.d = x.re;
  return v.u == 0x7ff8dead0000beefULL;
}

// libgcc's full complex division, twice, out of line (the rare path of eval_cell)
static MCE_HDN MCE_NOINLINE void g2_cdiv_pair_full(cplx u1, cplx v1, cplx u2, cplx v2, cplx* r1, cplx* r2) { *r1 = cdiv(u1, v1); *r2 = cdiv(u2, v2); }

// LEAN: no second value table in shared memory (16 + 4 instead of 32 + 4 bytes per cell): tables of up to ~11 000 cells fit (d = 7 with 16 hyperplanes:
// 9 949), at the price of ~10 % more time on the tables both variants can hold (measured) -- the engine uses it where the
// standard variant does not fit.
template <int MODE, bool LEAN = false>
struct KGTable2T {
  static constexpr int kMaxThreads = 128, kMinBlocks = 8;   // 64 registers: 8 CTAs (32 warps) per SM
  StepParams sp; GenView prev; GenView next; ParentWs ws; TermView tv;
  int m, g0;
  const int* order; const int* grp_start;
  int HC;                       // capacity of the per-cell arrays (max cells of any table this step)
  int NW;                       // bitmap words: 2^max_shape / 32
  unsigned char* alive_flag; int* diag;
  int big_T;                    // G2_NORMAL: groups with more members are left to the split launches
  BigArgs big;
  int gid_shift = 0;            // slot of group gi in the new generation = gid_begin[m] + gi + gid_shift (sharded layouts pad each phase)
  static MCE_HD size_t smem_bytes(int HC, int NW) {
    return ((sizeof(Group2Sm) + 15) & ~(size_t)15) + (size_t)HC * (sizeof(unsigned) + (LEAN ? 1 : 2) * sizeof(cplx)) + (size_t)NW * 2 * (sizeof(unsigned) + sizeof(unsigned short)) + 2 * (16 + kMaxThreads) * sizeof(unsigned short) + 64;
  }

  template <class Ctx> MCE_KERNEL_FN void bm_prefix(Ctx& c, const unsigned* bm, unsigned short* pf, int nw, int* total) const { bm_prefix_any(c, bm, pf, nw, total); }

  // B_mu of a parent as (source keys, count, xor mask): B^{k|k-1} ^ sign(A H) ^ in-place re-orientations
  MCE_HD void parent_B_src(int r, const unsigned** src, int* n, unsigned* mask) const {
    const int gid = prev.alive[r], phc = gen_m(prev, gid);
    *src = sp.with_tp ? ws.tpB + (long long)r * ws.tpB_stride : gen_keys(prev, gid, phc);
    *n = sp.with_tp ? ws.tpB_cells[r] : prev.cells[gid];
    *mask = ws.sgnmask[r] ^ ws.bxor[r];
  }

  // One thread per member of the chunk: every dependent global load of the member's description happens here, in parallel.
  MCE_HD void load_member(Group2Member* e, int ti) const {
    const long long gt = tv.t_begin[m] + ti;
    const SlotMeta me = tv.meta[gt];
    const int gidp = prev.alive[me.parent], phc = gen_m(prev, gidp);
    e->ti = ti; e->parent = me.parent; e->gidp = gidp; e->phc = phc; e->pc = prev.cells[gidp];
    e->own_cells = sp.with_tp ? ws.tpB_cells[me.parent] : e->pc;
    e->rk_off = gen_rk_off(prev, gidp, phc);
    e->pG = gen_G(prev, gidp, phc); e->bmP = prev.rbm + e->rk_off; e->pfP = prev.rpf + e->rk_off;
    e->hflag = me.hflag; e->enc_lhp = me.enc_lhp; e->csneg = me.csneg; e->mask = ws.sgnmask[me.parent];   // bxor is read when needed (it can change)
    e->z = me.z; e->is_child = me.flags & 1; e->has_cmap = (me.flags >> 1) & 1; e->pbc = me.pbc;
    e->c = me.c_val; e->d = me.d_val;
    const double* p = term_p(tv, m, ti);
    double s = 0; for (int i = 0; i < m; i++) s += p[i];       // sum_vec, flat:106-107
    e->psq = s * s;
    // Negligibility certificate (flat:242-247 keeps a member only if some cell has (sum p)^2 |G| > 1e-15).  Every cell is
    // G = [G_p(l+)/(ygi + d + ic) - G_p(l-)/(ygi - d + ic)] * scale with |ygi +- d + ic| >= |c| and |G_p| <= gmax_p, hence
    // |G| <= 2 gmax_p scale / |c| up to rounding (a few 1e-16 relative per operation).  With a factor 2 to spare the computed
    // product of EVERY cell is below the threshold, so the member is negligible exactly as the cell-by-cell test would find.
    { const double bound = 4.0 * prev.gmax[gidp] * fabs(sp.gscale) / fabs(me.c_val); e->skip = (e->psq * bound <= TERM_APPROXIMATION_EPS) ? 1 : 0; }
    e->kflip = 0;
    if ((me.flags & 3) == 3) {         // coaligned new child: parent position k <- child row cmap[l], flipped by cs_map[l] (flat:194-217)
      const unsigned char* cm = tv.cmap + gt * MAXM;
      unsigned flip = 0; int k = 0, l = 0;
      while (k < phc) {
        if (k == me.z) { e->ksrc[k] = 255; k++; if (k == phc) break; }
        e->ksrc[k] = cm[l]; if ((me.csneg >> l) & 1u) flip |= (1u << k);
        k++; l++;
      }
      e->kflip = flip;
    }
  }

  // Stage member `e` into slot `sl`: q and -- when `ref` >= 0 -- the per-row orientation bits between term `ref` and this
  // member (ce:586-599).  Runs inside a phase (threads < m take part); the parent's rank structure is read from HBM/L1.
  MCE_HD void stage_qs(Group2Sm* sm, const Group2Member* e, int sl, int ref, int tid) const {
    if (tid < m) {
      sm->q[sl][tid] = ((e->hflag >> tid) & 1u) ? 0.0 : term_q(tv, m, e->ti)[tid];
      if (ref >= 0) sm->sgbits[sl][tid] = orient_bit(term_A(tv, m, ref, sp.d), term_A(tv, m, e->ti, sp.d), tid, sp.d);
    }
  }
  MCE_HD unsigned sigma_of(const Group2Sm* sm, int sl) const { unsigned s = 0; for (int i = 0; i < m; i++) s |= sm->sgbits[sl][i]; return s; }

  // G_p lookup through the rank structure (eval_gs.hpp:94-153 semantics: half storage, conjugate of the opposite cell, 0 when absent)
  MCE_HD cplx lookup(const Group2Member* e, const cplx* pG, const unsigned* bmP, const unsigned short* pfP, int enc_l) const {   // bmP / pfP: the parent's rank structure
    const int phc = e->phc, top = 1 << (phc - 1), rev = (1 << phc) - 1;
    const bool cj = (enc_l & top) != 0;
    const int r = bitmap_rank(bmP, pfP, (unsigned)(cj ? (rev ^ enc_l) : enc_l));
    if (r < 0) return make_cplx(0, 0);
    const cplx v = pG[r];
    return cj ? cconj(v) : v;
  }

  // G of one cell of the staged member, flattening.hpp:129-247
  MCE_HDN MCE_NOINLINE cplx eval_cell(const double* q, const Group2Member* e, int* flag, unsigned key) const {
    const int phc = e->phc;
    const unsigned* bmP = e->bmP; const unsigned short* pfP = e->pfP;
    // ygi = sum over non-H-orthogonal rows of q_k s_k, in row order (flat:137-154).  sm->q holds +0.0 for the H-orthogonal
    // rows (adding +0.0 never changes a running sum that started at +0.0), so the loop is branch-free; s_k flips the sign bit.
    double ygi = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 16; k++) { if (k >= m) break; ygi += flip_sign(q[k], (key >> k) & 1u); }   // this kernel serves max_shape <= 16
    int lp, lm;
    const int phc_mask = (1 << phc) - 1;
    if (!e->is_child) { lp = (int)(key & (unsigned)phc_mask); lm = lp; }
    else {
      const int z = e->z;
      if (!e->has_cmap) {              // insert a zero bit at position z, truncate to phc bits
        const unsigned low = key & ((1u << z) - 1u), high = (z < 31) ? ((key >> z) << (z + 1)) : 0u;
        lp = (int)((z < phc ? (low | high) : key) & (unsigned)phc_mask);
      } else {
        unsigned v = 0;
        for (int k = 0; k < phc; k++) { const unsigned s = e->ksrc[k]; if (s != 255u) v |= ((key >> s) & 1u) << k; }
        lp = (int)((v ^ e->kflip) & (unsigned)phc_mask);
        if (z < phc) lp &= ~(1 << z);
      }
      lm = (z < phc) ? (lp | (1 << z)) : lp;
    }
    const cplx* pG = e->pG;
    const cplx gp = lookup(e, pG, bmP, pfP, lp ^ (int)e->enc_lhp);
    const cplx gm = lookup(e, pG, bmP, pfP, lm ^ (int)e->enc_lhp);
    const cplx vp = make_cplx(ygi + e->d, e->c), vm = make_cplx(ygi - e->d, e->c);
    cplx rp, rm;
    if (!cdiv2_fast(gp, vp, gm, vm, &rp, &rm)) g2_cdiv_pair_full(gp, vp, gm, vm, &rp, &rm);   // rare: guards, absent cells (0 numerators)
    cplx g = csub(rp, rm);
    g = cscale(g, sp.gscale);
    if (!G2_FLAG_READ(flag)) {             // |G| only matters until one cell is found non-negligible (flat:242-247)
      // max(|re|,|im|) <= |G| <= |re|+|im| and rounding is monotone, so the two cheap bounds decide almost every cell
      // exactly as `psq * cabs(G) > eps` would; hypot runs only in between.
      const double ar = fabs(g.re), ai = fabs(g.im), mx = ar > ai ? ar : ai;
      if ((e->psq * mx) > TERM_APPROXIMATION_EPS) G2_FLAG_SET(flag);
      else if ((e->psq * (ar + ai)) > TERM_APPROXIMATION_EPS) { if ((e->psq * cabs_(g)) > TERM_APPROXIMATION_EPS) G2_FLAG_SET(flag); }
    }
    return g;
  }

  // Shared-memory carve-up and group description of one CTA.
  struct Ws {
    Group2Sm* sm; cplx* acc; cplx* Gm; unsigned* Bk; unsigned* bmP; unsigned* bmA; unsigned short* pfP; unsigned short* pfA;
    const int* members; int ncomb, gid_out, nwM, cbase; unsigned rev_m, top_m;
  };

  // ---- B-table of the root (K7) and the root's own table, with re-election when the candidate is negligible (flat:399-489).
  // Returns false when every member is negligible.  `primary` = this CTA owns the group's side effects.
  template <class Ctx> MCE_KERNEL_FN bool root_phase(Ctx& c, Ws& w, bool primary, int* nB_out, int* k_out, int* rsel_out) const {
    const int d = sp.d;
    Group2Sm* sm = w.sm; cplx* acc = w.acc; unsigned* Bk = w.Bk; unsigned* bmP = w.bmP; unsigned* bmA = w.bmA;
    unsigned short* pfP = w.pfP;
    const int ncomb = w.ncomb, nwM = w.nwM; const unsigned rev_m = w.rev_m, top_m = w.top_m;
    const int* members = w.members;
    auto load_chunk = [&](int tid) {
      if (tid < G2_CHUNK) { sm->flag[tid] = 0; if (w.cbase + tid < ncomb) load_member(&sm->mem[tid], members[w.cbase + tid]); }
    };
    w.cbase = 0;
    c.par([&](int tid) { load_chunk(tid); if (tid == 0) sm->owner = -1; });
    const Group2Member* eR = &sm->mem[0];
    int nB = 0;
    // A group rooted at a new child holds new children only (old terms sort first), so no candidate owns a parent's B memory and
    // root election has no side effect.  When every member is certified negligible (load_member) the group dies right here,
    // before its B-table is enumerated.
    if (eR->is_child && ncomb <= G2_CHUNK) {
      bool dead = true;
      for (int k = 0; k < ncomb; k++) if (!sm->mem[k].skip || !sm->mem[k].is_child) dead = false;
      if (dead) { *nB_out = 0; *k_out = ncomb; *rsel_out = -1; return false; }
    }
    if (eR->is_child) {
      if (m <= d) {
        nB = 1 << (m - 1);
        c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) Bk[i] = (unsigned)i; });
      } else {
        const unsigned* src; int ncp; unsigned mask;
        parent_B_src(eR->parent, &src, &ncp, &mask);
        const int pbc = eR->pbc, z = eR->z;
        const int nwPb = pbc >= 5 ? (1 << (pbc - 5)) : 1;
        unsigned* bmC = bmP;            // the child-key bitmap borrows bmP (not in use before the first stage_member)
        c.par([&](int tid) {
          MCE_NOUNROLL for (int i = tid; i < nwPb; i += c.nthreads()) bmA[i] = 0;
          MCE_NOUNROLL for (int i = tid; i < nwM; i += c.nthreads()) bmC[i] = 0;
        });
        c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < ncp; i += c.nthreads()) { const unsigned b = src[i] ^ mask; c.atomic_or(&bmA[b >> 5], 1u << (b & 31)); } });
        const unsigned mask_z = 1u << z, hbit = 1u << (pbc - 1), rev_pbc = (pbc >= 32) ? 0xffffffffu : ((1u << pbc) - 1u), mask_low = (1u << z) - 1u;
        const bool coal = m < pbc;
        const unsigned char* cm = tv.cmap + (tv.t_begin[m] + eR->ti) * MAXM;
        c.par([&](int tid) {            // pairs (b, b ^ 2^z) both present -> two child sign vectors (ce:261-320), coalesced on the fly (ce:323-366)
          unsigned sel[MAXM]; int nsel = 0;
          if (coal) { unsigned seen = 0; for (int j = 0; j < pbc; j++) { const unsigned ci = cm[j]; if (!((seen >> ci) & 1u)) { seen |= (1u << ci); sel[nsel++] = 1u << j; } } }
          MCE_NOUNROLL for (int i = tid; i < ncp; i += c.nthreads()) {
            const unsigned b = src[i] ^ mask;
            unsigned bq = b ^ mask_z;
            if (bq & hbit) bq ^= rev_pbc;
            if (!(b < bq) || !((bmA[bq >> 5] >> (bq & 31)) & 1u)) continue;
            const unsigned z_bit = (b & mask_z) >> z;
            const unsigned csv1 = ((b >> (z + 1)) << z) | (b & mask_low) | (z_bit << (pbc - 1)), csv2 = csv1 ^ hbit;
            unsigned k1 = (csv1 & hbit) ? csv1 ^ rev_pbc : csv1, k2 = (csv2 & hbit) ? csv2 ^ rev_pbc : csv2;
            if (coal) {
              unsigned c1 = 0, c2 = 0;
              for (int l = 0; l < nsel; l++) { if (k1 & sel[l]) c1 |= (1u << l); if (k2 & sel[l]) c2 |= (1u << l); }
              k1 = c1; k2 = c2;
            }
            c.atomic_or(&bmC[k1 >> 5], 1u << (k1 & 31));
            c.atomic_or(&bmC[k2 >> 5], 1u << (k2 & 31));
          }
        });
        bm_prefix(c, bmC, pfP, nwM, &sm->cnt);
        c.par([&](int tid) {            // enumerate the set bits: keys in ascending order
          MCE_NOUNROLL for (int wd = tid; wd < nwM; wd += c.nthreads()) {
            unsigned bits = bmC[wd]; int o = pfP[wd];
            while (bits) { const int b = MCE_FFS(bits); bits &= bits - 1u; Bk[o++] = (unsigned)(wd * 32 + b); }
          }
        });
        nB = sm->cnt;                    // written two barriers ago; not modified again
      }
    } else {
      const unsigned* src; unsigned mask;
      parent_B_src(eR->parent, &src, &nB, &mask);
      c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) Bk[i] = src[i] ^ mask; if (tid == 0) sm->owner = eR->parent; });
    }

    auto nothing = [](int) {};
    int k = 0, accepted = 0, lfr = -1;
    const Group2Member* e = eR;
    for (;;) {
      // candidate k (chunk-local index kk)
      if (k - w.cbase >= G2_CHUNK) { c.par(nothing); w.cbase = k; c.par([&](int tid) { load_chunk(tid); }); }   // the empty phase keeps slow readers of the old chunk ahead of its reload
      const int kk = k - w.cbase;
      e = &sm->mem[kk];
      if (k > 0 && !e->is_child) {       // old term: its own table becomes the group's table (flat:443-473)
        const unsigned* src; unsigned mask;
        parent_B_src(e->parent, &src, &nB, &mask);
        c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) Bk[i] = src[i] ^ mask; if (tid == 0) sm->owner = e->parent; });
      }
      { const int ref = (k > 0 && e->is_child) ? lfr : -1; c.par([&](int tid) { stage_qs(sm, e, 0, ref, tid); }); }
      unsigned sigma = 0;
      if (k > 0 && e->is_child) {        // new child: re-orient the group's table in place (flat:433-441)
        sigma = sigma_of(sm, 0);
        if (sigma & top_m) sigma ^= rev_m;
      }
      c.par([&](int tid) {
        if (e->skip) {                   // certified negligible (load_member): only the re-orientation of the table happens
          if (sigma) { MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) Bk[i] ^= sigma; }
        } else {
          MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) { const unsigned key = Bk[i] ^ sigma; Bk[i] = key; acc[i] = eval_cell(sm->q[0], e, &sm->flag[kk], key); }
        }
        if (sigma && tid == 0 && sm->owner >= 0 && primary) c.atomic_xor(ws.bxor + sm->owner, sigma);   // the table is a parent's B memory, shared with its children
      });
      if (sm->flag[kk]) { accepted = 1; break; }
      lfr = e->ti;
      if (++k >= ncomb) break;
    }
    *nB_out = nB; *k_out = k; *rsel_out = accepted ? e->ti : -1;
    return accepted != 0;
  }

  // ---- members [k_from, k_to) (flat:491-550): G table of the member, added cell by cell to the root's.  When `rows` is
  // given (split groups) the addends are stored instead -- row k, cell i = what member k adds to root cell i -- and summed
  // later in member order by final_phase.
  template <class Ctx> MCE_KERNEL_FN void members_phase(Ctx& c, Ws& w, int nB, int rsel, int k_from, int k_to, cplx* rows, int* rflags, int row_stride) const {
    Group2Sm* sm = w.sm; cplx* acc = w.acc; cplx* Gm = w.Gm; unsigned* Bk = w.Bk; unsigned* bmP = w.bmP; unsigned* bmA = w.bmA;
    unsigned short* pfP = w.pfP; unsigned short* pfA = w.pfA;
    const int ncomb = w.ncomb, nwM = w.nwM; const unsigned rev_m = w.rev_m, top_m = w.top_m;
    const int* members = w.members;
    auto load_chunk = [&](int tid) {
      if (tid < G2_CHUNK) { sm->flag[tid] = 0; if (w.cbase + tid < ncomb) load_member(&sm->mem[tid], members[w.cbase + tid]); }
    };
    int pend = 0, pend_k = 0; bool pend_cj = false;          // deferred "acc += Gm" of the previous member (runs inside the next phase)
    auto do_pending = [&](int tid) {
      if (!pend) return;
      if (rows) {
        cplx* row = rows + (long long)pend_k * row_stride;
        MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) row[i] = pend_cj ? cconj(Gm[i]) : Gm[i];
        if (tid == 0) rflags[pend_k] = 1;
      } else {
        MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) acc[i] = cadd(acc[i], pend_cj ? cconj(Gm[i]) : Gm[i]);
      }
    };
    // One barrier per member: the phase that evaluates member k also adds member k-1's table (same thread, same cells) and
    // stages q / orientation bits of member k+1 into the other slot.
    bool staged = false;                 // member k's q / sgbits already sit in slot k & 1
    for (int k = k_from; k < k_to; ++k) {
      if (k - w.cbase >= G2_CHUNK || k < w.cbase) { c.par([&](int tid) { do_pending(tid); }); pend = 0; w.cbase = k; c.par([&](int tid) { load_chunk(tid); }); staged = false; }
      const int kk = k - w.cbase, sl = k & 1;
      const Group2Member* et = &sm->mem[kk];
      if (et->skip) { staged = false; continue; }       // certified negligible (load_member): contributes nothing, costs nothing
      if (!staged) c.par([&](int tid) { stage_qs(sm, et, sl, rsel, tid); });
      const bool next_here = (k + 1 < k_to) && (kk + 1 < G2_CHUNK);       // member k+1 is in the loaded chunk: stage it during this phase
      const Group2Member* en = et + 1;
      const unsigned sigma_raw = sigma_of(sm, sl);
      const bool cj = (sigma_raw & top_m) != 0;
      if (et->is_child || et->own_cells != nB) {
        // table = root's table re-oriented (update_btable, ce:584-625): cell i of the member is cell i of the root
        const unsigned sigma_n = cj ? (sigma_raw ^ rev_m) : sigma_raw;
        const int pd = pend, pk = pend_k; const bool pcj = pend_cj;
        cplx* prow = rows ? rows + (long long)pk * row_stride : nullptr;
        c.par([&](int tid) {
          MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) {
            if (pd) { const cplx a = pcj ? cconj(Gm[i]) : Gm[i]; if (prow) prow[i] = a; else acc[i] = cadd(acc[i], a); }
            Gm[i] = eval_cell(sm->q[sl], et, &sm->flag[kk], Bk[i] ^ sigma_n);
          }
          if (pd && prow && tid == 0) rflags[pk] = 1;
          if (!et->is_child && tid == 0) c.atomic_add(diag, 1);     // flat:516-539 also rewrites the parent's B memory: not modelled
          if (next_here) stage_qs(sm, en, sl ^ 1, rsel, tid);
        });
        pend = 0;
        if (sm->flag[kk]) { pend = 1; pend_cj = cj; pend_k = k; }
      } else {
        // old term with its own table: cell i of its B_mu is position i of its (sorted) source table; add by key (flat:291-314)
        const unsigned* src; int nT; unsigned mask;
        parent_B_src(et->parent, &src, &nT, &mask);
        if (sp.with_tp) {               // on TP steps B_mu comes from the DCE-TP table, not from the G-table keys: rank over tpB
          c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nwM; i += c.nthreads()) bmA[i] = 0; });
          c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nT; i += c.nthreads()) { const unsigned b = src[i]; c.atomic_or(&bmA[b >> 5], 1u << (b & 31)); } });
          bm_prefix(c, bmA, pfA, nwM, (int*)nullptr);
        }
        const unsigned* bmT = sp.with_tp ? bmA : prev.rbm + et->rk_off; const unsigned short* pfT = sp.with_tp ? pfA : prev.rpf + et->rk_off;
        c.par([&](int tid) {
          do_pending(tid);
          MCE_NOUNROLL for (int i = tid; i < nT; i += c.nthreads()) Gm[i] = eval_cell(sm->q[sl], et, &sm->flag[kk], src[i] ^ mask);
          if (next_here) stage_qs(sm, en, sl ^ 1, rsel, tid);
        });
        pend = 0;
        if (sm->flag[kk])
          c.par([&](int tid) {
            cplx* row = rows ? rows + (long long)k * row_stride : nullptr;
            MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) {
              unsigned kq = Bk[i] ^ sigma_raw; bool cjj = false;
              if (kq & top_m) { cjj = true; kq ^= rev_m; }
              const int jj = bitmap_rank(bmT, pfT, kq ^ mask);      // position of the member's cell with key kq
              if (row) row[i] = jj >= 0 ? (cjj ? cconj(Gm[jj]) : Gm[jj]) : skip_addend();
              else if (jj >= 0) acc[i] = cadd(acc[i], cjj ? cconj(Gm[jj]) : Gm[jj]);
            }
            if (row && tid == 0) rflags[k] = 1;
          });
      }
      staged = next_here;
    }
    if (pend) { c.par([&](int tid) { do_pending(tid); }); pend = 0; }
  }

  // ---- members [k_from, k_to), LEAN variant: the same sums without the member's value table.
  //
  // A member's table is only added when one of its cells is not negligible (flat:242-247), so its values cannot go into the root's
  // sums before that is known -- but they need not be STORED either.  Every thread evaluates its first cell into a register; almost
  // every accepted member raises its flag right there.  The barrier that ends the phase publishes the flag; the member's other
  // cells are evaluated in the NEXT phase straight into the sums (same thread, same cells, so the order of the additions to a
  // cell is the member order), together with the first cell of the next member.  While no flag is up, a thread goes on evaluating
  // its further cells only to settle the negligibility test (the values are dropped; they are re-evaluated in the rare case that
  // a late cell raises the flag).  One barrier per member, no second value table in shared memory.
  struct MemberPlan {
    int on, k, kk, sl; bool cj, own; unsigned sigma_n, maskT; const unsigned* bmT; const Group2Member* e;
  };
  // What member P does to root cell i: the member's cell with key Bk[i] ^ sigma_n, conjugated when cj (update_btable, ce:584-625:
  // cell i of the member is cell i of the root).  An old term that keeps its own table (flat:291-314) only has the cells its
  // (sorted) source table holds: a bit test in that table's rank bitmap.
  MCE_HD bool member_has(const MemberPlan& P, unsigned key) const {
    if (!P.own) return true;
    const unsigned b = key ^ P.maskT;
    return ((P.bmT[b >> 5] >> (b & 31)) & 1u) != 0;
  }
  template <class Ctx> MCE_KERNEL_FN void members_phase_lean(Ctx& c, Ws& w, int nB, int rsel, int k_from, int k_to, cplx* rows, int* rflags, int row_stride) const {
    Group2Sm* sm = w.sm; cplx* acc = w.acc; unsigned* Bk = w.Bk; unsigned* bmA = w.bmA;
    const int ncomb = w.ncomb, nwM = w.nwM, NT = c.nthreads(); const unsigned rev_m = w.rev_m, top_m = w.top_m;
    const int* members = w.members;
    auto v0 = c.template priv<cplx>();   // the member's first cell of this thread, waiting for the flag
    auto load_chunk = [&](int tid) {
      if (tid < G2_CHUNK) { sm->flag[tid] = 0; if (w.cbase + tid < ncomb) load_member(&sm->mem[tid], members[w.cbase + tid]); }
    };
    MemberPlan pend; pend.on = 0;          // accepted member whose cells (but the first) are still to be evaluated
    auto do_rest = [&](int tid) {
      if (!pend.on) return;
      cplx* row = rows ? rows + (long long)pend.k * row_stride : nullptr;
      MCE_NOUNROLL for (int i = tid; i < nB; i += NT) {
        const unsigned key = Bk[i] ^ pend.sigma_n;
        if (!member_has(pend, key)) { if (row) row[i] = skip_addend(); continue; }
        cplx g = (i == tid) ? v0[tid] : eval_cell(sm->q[pend.sl], pend.e, &sm->flag[pend.kk], key);
        if (pend.cj) g = cconj(g);
        if (row) row[i] = g; else acc[i] = cadd(acc[i], g);
      }
      if (row && tid == 0) rflags[pend.k] = 1;
    };
    // One phase body serves every case (the rest of the pending member, then the first cells of member `cur` when cur.on): the code of the
    // two evaluation loops exists once -- the kernel is instruction-cache bound as soon as they are duplicated.
    bool staged = false;                 // the member's q / sgbits already sit in slot nsl
    bool own_ready = false;              // member k keeps its own table and the scratch bitmap / counter are prepared for it
    int last_sl = 0, nsl = 1;            // slot of the last evaluated member; slot the next member was staged into
    enum { AFTER_NONE = 0, AFTER_RELOAD = 1, AFTER_OWN = 2, AFTER_END = 3 };
    MemberPlan cur; cur.on = 0;
    const unsigned* src = nullptr; int nT = 0;
    int k = k_from;
    for (;;) {
      int after = AFTER_NONE;
      bool next_here = false;
      const Group2Member* et = nullptr; const Group2Member* en = nullptr;
      cur.on = 0;
      if (k >= k_to) { if (!pend.on) break; after = AFTER_END; }
      else if (k - w.cbase >= G2_CHUNK || k < w.cbase) after = AFTER_RELOAD;     // the phase below finishes the pending member (and is the barrier in front of the reload)
      else {
        const int kk = k - w.cbase;
        et = &sm->mem[kk];
        if (et->skip) { staged = false; ++k; continue; }       // certified negligible (load_member): contributes nothing, costs nothing
        const bool own = !(et->is_child || et->own_cells != nB);
        if (own && !own_ready) after = AFTER_OWN;                // the scratch bitmap may still serve the pending member: finish that one first
        else {
          const int sl = staged ? nsl : (last_sl + 1) % G2_SLOTS;
          if (!staged) c.par([&](int tid) { stage_qs(sm, et, sl, rsel, tid); });
          nsl = (sl + 1) % G2_SLOTS;
          next_here = (k + 1 < k_to) && (kk + 1 < G2_CHUNK);     // member k+1 is in the loaded chunk: stage it during this phase
          en = et + 1;
          const unsigned sigma_raw = sigma_of(sm, sl);
          cur.on = 1; cur.k = k; cur.kk = kk; cur.sl = sl; cur.e = et;
          cur.cj = (sigma_raw & top_m) != 0;
          cur.sigma_n = cur.cj ? (sigma_raw ^ rev_m) : sigma_raw;
          cur.own = own; cur.maskT = 0; cur.bmT = nullptr;
          if (own) { unsigned mask; parent_B_src(et->parent, &src, &nT, &mask); cur.bmT = sp.with_tp ? bmA : prev.rbm + et->rk_off; cur.maskT = mask; }
        }
      }
      c.par([&](int tid) {
        do_rest(tid);                      // consumes the previous member's v0 before this member's first cell replaces it
        if (!cur.on) return;
        int matched = 0;
        MCE_NOUNROLL for (int i = tid; i < nB; i += NT) {
          const unsigned key = Bk[i] ^ cur.sigma_n;
          if (!member_has(cur, key)) continue;
          matched++;
          if (i != tid && G2_FLAG_READ(&sm->flag[cur.kk])) { if (cur.own) continue; break; }   // accepted: the other cells wait for the next phase
          const cplx g = eval_cell(sm->q[cur.sl], et, &sm->flag[cur.kk], key);
          if (i == tid) v0[tid] = g;
        }
        if (cur.own && matched) c.atomic_add(&sm->cnt, matched);
        if (!et->is_child && !cur.own && tid == 0) c.atomic_add(diag, 1);     // flat:516-539 also rewrites the parent's B memory: not modelled
        if (next_here) stage_qs(sm, en, nsl, rsel, tid);
      });
      pend.on = 0;
      if (cur.on) {
        bool accepted = sm->flag[cur.kk] != 0;
        if (!accepted && cur.own && sm->cnt != nT) {
          // cells of the member's table without a counterpart in the root's: they take part in the negligibility test all the same
          c.par([&](int tid) {
            MCE_NOUNROLL for (int j = tid; j < nT; j += NT) { if (G2_FLAG_READ(&sm->flag[cur.kk])) break; (void)eval_cell(sm->q[cur.sl], et, &sm->flag[cur.kk], src[j] ^ cur.maskT); }
          });
          accepted = sm->flag[cur.kk] != 0;
        }
        if (accepted) pend = cur;
        last_sl = cur.sl; staged = next_here; own_ready = false;
        ++k;
      } else if (after == AFTER_RELOAD) {
        w.cbase = k; c.par([&](int tid) { load_chunk(tid); }); staged = false;
      } else if (after == AFTER_OWN) {
        // old term with its own table: cell i of its B_mu is position i of its (sorted) source table; root cell i takes the cell with the
        // same key, when the table has it (flat:291-314).  On TP steps B_mu comes from the DCE-TP table, not from the G-table keys: bitmap over tpB
        unsigned mask;
        parent_B_src(et->parent, &src, &nT, &mask);
        c.par([&](int tid) { if (tid == 0) sm->cnt = 0; if (sp.with_tp) { MCE_NOUNROLL for (int i = tid; i < nwM; i += NT) bmA[i] = 0; } });
        if (sp.with_tp) c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nT; i += NT) { const unsigned b = src[i]; c.atomic_or(&bmA[b >> 5], 1u << (b & 31)); } });
        own_ready = true;
      } else break;
    }
  }

  // ---- write the surviving term; rank of a key in the bitmap of the final keys = its sorted position (flat:251-252) ----
  template <class Ctx> MCE_KERNEL_FN void emit(Ctx& c, Ws& w, int nB, int rsel) const {
    const int d = sp.d, nwM = w.nwM, gid_out = w.gid_out;
    cplx* acc = w.acc; unsigned* Bk = w.Bk; unsigned* bmA = w.bmA; unsigned short* pfA = w.pfA;
    c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nwM; i += c.nthreads()) bmA[i] = 0; });
    c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) { const unsigned b = Bk[i]; c.atomic_or(&bmA[b >> 5], 1u << (b & 31)); } });
    bm_prefix(c, bmA, pfA, nwM, (int*)nullptr);
    unsigned* ko = gen_keys(next, gid_out, m); cplx* Go = gen_G(next, gid_out, m);
    const double* As = term_A(tv, m, rsel, d); const double* ps = term_p(tv, m, rsel); const double* bs = term_b(tv, m, rsel, d);
    double* Ao = gen_A(next, gid_out, m, d); double* po = gen_p(next, gid_out, m); double* bo = gen_b(next, gid_out, d);
    c.par([&](int tid) {
      MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) { const unsigned b = Bk[i]; const int pos = bitmap_rank(bmA, pfA, b); ko[pos] = b; Go[pos] = acc[i]; }
      for (int i = tid; i < m * d; i += c.nthreads()) Ao[i] = As[i];
      MCE_NOUNROLL for (int i = tid; i < m; i += c.nthreads()) po[i] = ps[i];
      MCE_NOUNROLL for (int i = tid; i < d; i += c.nthreads()) bo[i] = bs[i];
      if (tid == 0) {
        alive_flag[gid_out] = 1; next.cells[gid_out] = nB; next.g_m[gid_out] = (unsigned char)m;
        c.atomic_add_u64((unsigned long long*)(diag + 2), (unsigned long long)nB);
      }
    });
  }

  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    unsigned char* base = c.smem();
    Ws w;
    w.sm = (Group2Sm*)base;
    w.acc = (cplx*)(base + ((sizeof(Group2Sm) + 15) & ~(size_t)15));
    w.Gm = w.acc + HC;
    w.Bk = (unsigned*)(w.Gm + (LEAN ? 0 : HC));
    w.bmP = w.Bk + HC;             // parent-table rank structure of the staged member
    w.bmA = w.bmP + NW;            // scratch bitmap: parent B_mu / TP table / final keys
    w.pfP = (unsigned short*)(w.bmA + NW);
    w.pfA = w.pfP + NW + 16 + c.nthreads();   // prefix arrays carry chunk totals behind them
    int gi, part = 0; BigGroup* bg = nullptr;
    if (MODE == G2_NORMAL) gi = g0 + c.block();
    else if (MODE == G2_BIG_PARTS) { const BigPart bp = big.parts[c.block()]; bg = big.groups + bp.slot; part = bp.part; gi = bg->gi; }
    else { bg = big.groups + c.block(); gi = bg->gi; }
    const int start = grp_start[gi];
    w.ncomb = grp_start[gi + 1] - start;
    w.members = order + start;
    w.gid_out = next.gid_begin[m] + gi + gid_shift;
    w.rev_m = (1u << m) - 1u; w.top_m = 1u << (m - 1);
    w.nwM = m >= 5 ? (1 << (m - 5)) : 1;
    w.cbase = 0;
    if (MODE == G2_NORMAL) {
      if (w.ncomb > big_T) return;               // split groups are handled by the G2_BIG_* launches
      int nB, k, rsel;
      if (!root_phase(c, w, true, &nB, &k, &rsel)) {
        c.par([&](int tid) { if (tid == 0) { alive_flag[w.gid_out] = 0; next.cells[w.gid_out] = 0; next.g_m[w.gid_out] = (unsigned char)m; } });
        return;
      }
      if (LEAN) members_phase_lean(c, w, nB, rsel, k + 1, w.ncomb, (cplx*)nullptr, (int*)nullptr, 0);
      else members_phase(c, w, nB, rsel, k + 1, w.ncomb, (cplx*)nullptr, (int*)nullptr, 0);
      emit(c, w, nB, rsel);
      return;
    }
    cplx* rows = big.rows + bg->rows_off; int* rflags = big.flags + bg->flags_off; unsigned* keys = big.keys + bg->keys_off;
    if (MODE == G2_BIG_ROOT) {                   // root election; the group's keys and the root's own table go to scratch
      int nB, k, rsel;
      const bool ok = root_phase(c, w, true, &nB, &k, &rsel);
      c.par([&](int tid) {
        if (tid == 0) {
          bg->hdr[0] = nB; bg->hdr[1] = rsel; bg->hdr[2] = ok ? 1 : 0; bg->hdr[3] = k;
          if (!ok) { alive_flag[w.gid_out] = 0; next.cells[w.gid_out] = 0; next.g_m[w.gid_out] = (unsigned char)m; }
        }
        if (ok) { MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) { keys[i] = w.Bk[i]; rows[i] = w.acc[i]; } }
      });
      return;
    }
    const int nB = bg->hdr[0], rsel = bg->hdr[1], ok = bg->hdr[2], kacc = bg->hdr[3];
    if (!ok) return;
    if (MODE == G2_BIG_PARTS) {                  // BIG_PART members of one split group; their addends go to scratch rows
      const int lo = 1 + part * BIG_PART, hi = lo + BIG_PART < w.ncomb ? lo + BIG_PART : w.ncomb;
      const int from = lo > kacc + 1 ? lo : kacc + 1;
      if (from >= hi) return;
      c.par([&](int tid) { MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) w.Bk[i] = keys[i]; });
      w.cbase = from + 1;                        // forces the first iteration to load its chunk
      if (LEAN) members_phase_lean(c, w, nB, rsel, from, hi, rows, rflags, big.row_stride);
      else members_phase(c, w, nB, rsel, from, hi, rows, rflags, big.row_stride);
      return;
    }
    // G2_BIG_FINAL: root table + every stored addend, in member order (the order fixes the floating-point sums)
    c.par([&](int tid) {
      MCE_NOUNROLL for (int i = tid; i < nB; i += c.nthreads()) {
        w.Bk[i] = keys[i];
        cplx a = rows[i];
        MCE_NOUNROLL for (int k = kacc + 1; k < w.ncomb; k++) {
          if (!rflags[k]) continue;
          const cplx v = rows[(long long)k * big.row_stride + i];
          if (!is_skip_addend(v)) a = cadd(a, v);
        }
        w.acc[i] = a;
      }
    });
    emit(c, w, nB, rsel);
  }
};

using KGTable2 = KGTable2T<G2_NORMAL>;

// Lists the reduction groups of [ga, gb) with more than T members and cuts them into parts of BIG_PART members.
struct KBigGroups {
  const int* grp_start; const int* roots /* device: [0] groups, [1] groups rooted at an old term */; int phase, T, Hm;
  BigGroup* groups; BigPart* parts; int* cnt /*[2]: groups, parts*/; unsigned long long* cnt64 /*[3]: rows, flags, keys*/;
  int rank = 0, world = 1;      // term-level sharding: only the groups of this rank's chunk of the phase are listed
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      int ga = phase == 0 ? 0 : roots[1], gb = phase == 0 ? roots[1] : roots[0];
      if (world > 1) { const int ch = (gb - ga + world - 1) / world, lo = ga + rank * ch, hi = lo + ch; ga = lo < gb ? lo : gb; gb = hi < gb ? hi : gb; }
      const int gi = ga + c.block() * c.nthreads() + tid;
      if (gi >= gb) return;
      const int size = grp_start[gi + 1] - grp_start[gi];
      if (size <= T) return;
      BigGroup g;
      g.gi = gi; g.ncomb = size; g.nparts = (size - 1 + BIG_PART - 1) / BIG_PART;
      const int slot = c.atomic_add(cnt, 1);
      g.part_base = c.atomic_add(cnt + 1, g.nparts);
      g.rows_off = (long long)c.atomic_add_u64(cnt64, (unsigned long long)size * Hm);
      g.flags_off = (long long)c.atomic_add_u64(cnt64 + 1, (unsigned long long)size);
      g.keys_off = (long long)c.atomic_add_u64(cnt64 + 2, (unsigned long long)Hm);
      g.hdr[0] = g.hdr[1] = g.hdr[2] = g.hdr[3] = 0;
      groups[slot] = g;
      for (int p = 0; p < g.nparts; p++) { BigPart bp; bp.slot = slot; bp.part = p; parts[g.part_base + p] = bp; }
    });
  }
};

}  // namespace mce
#endif
