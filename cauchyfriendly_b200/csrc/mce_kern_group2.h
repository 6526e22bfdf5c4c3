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

#if defined(__CUDA_ARCH__)
#define MCE_POPC(x) __popc(x)
#define MCE_FFS(x) (__ffs((int)(x)) - 1)
#define MCE_NOINLINE __noinline__
#else
#define MCE_POPC(x) __builtin_popcount(x)
#define MCE_FFS(x) (__builtin_ffs((int)(x)) - 1)
#define MCE_NOINLINE
#endif

struct Group2Sm {
  int nB, flag, cnt, owner, sigma, pad0;
  int t_m, t_phc, t_pcells, t_z, t_is_child, t_has_cmap;
  unsigned t_hflag, t_enc_lhp, t_csneg, t_mask;
  double t_c, t_d, t_psq;
  const cplx* t_pG;
  double q[MAXM];
  unsigned char cmap[MAXM];
  unsigned char ksrc[MAXM];     // coaligned child: child row feeding parent position k (255 = none)
  unsigned kflip;               // bit k: flip the sign at parent position k
};

// rank of `key` among the set bits of bm (pf = exclusive prefix popcounts per word), or -1 when absent
MCE_HD int bitmap_rank(const unsigned* bm, const unsigned short* pf, unsigned key) {
  const unsigned w = bm[key >> 5], bit = 1u << (key & 31);
  if (!(w & bit)) return -1;
  return (int)pf[key >> 5] + MCE_POPC(w & (bit - 1u));
}

struct KGTable2 {
  StepParams sp; GenView prev; GenView next; ParentWs ws; TermView tv;
  int m, g0;
  const int* order; const int* grp_start;
  int HC;                       // capacity of the per-cell arrays (max cells of any table this step)
  int NW;                       // bitmap words: 2^max_shape / 32
  unsigned char* alive_flag; int* diag;
  static MCE_HD size_t smem_bytes(int HC, int NW) {
    return ((sizeof(Group2Sm) + 15) & ~(size_t)15) + (size_t)HC * (sizeof(unsigned) + 2 * sizeof(cplx)) + (size_t)NW * 2 * (sizeof(unsigned) + sizeof(unsigned short)) + 2 * (16 + 1024) * sizeof(unsigned short) + 64;
  }

  // ---- bitmap helpers (block-wide) ----
  template <class Ctx> MCE_KERNEL_FN void bm_zero(Ctx& c, unsigned* bm, int nw) const {
    c.par([&](int tid) { for (int i = tid; i < nw; i += c.nthreads()) bm[i] = 0; });
  }
  template <class Ctx> MCE_KERNEL_FN void bm_prefix(Ctx& c, const unsigned* bm, unsigned short* pf, int nw, int* total) const {
    // nw is 2^(bits-5): 1..2048 words. Thread t sums a contiguous chunk, then a serial pass over the chunk sums.
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
      *total = acc;
    });
    c.par([&](int tid) {
      const int lo = tid * chunk, hi = lo + chunk < nw ? lo + chunk : nw;
      if (lo >= nw) return;
      const unsigned short base = pf[nw + tid];
      if (base) for (int i = lo; i < hi; i++) pf[i] = (unsigned short)(pf[i] + base);
    });
  }

  // B_mu of a parent as (source keys, count, xor mask): B^{k|k-1} ^ sign(A H) ^ in-place re-orientations
  MCE_HD void parent_B_src(int r, const unsigned** src, int* n, unsigned* mask) const {
    const int gid = prev.alive[r], phc = gen_m(prev, gid);
    *src = sp.with_tp ? ws.tpB + (long long)r * ws.tpB_stride : gen_keys(prev, gid, phc);
    *n = sp.with_tp ? ws.tpB_cells[r] : prev.cells[gid];
    *mask = ws.sgnmask[r] ^ ws.bxor[r];
  }

  // Stage term `ti`: scalars, q, coalignment maps, and the rank structure (bmP, pfP) of its parent's table.
  template <class Ctx> MCE_KERNEL_FN void stage_term(Ctx& c, Group2Sm* sm, unsigned* bmP, unsigned short* pfP, int ti) const {
    const long long gt = tv.t_begin[m] + ti;
    const SlotMeta me = tv.meta[gt];
    const int gidp = prev.alive[me.parent], phc = gen_m(prev, gidp), pc = prev.cells[gidp];
    const int nwP = phc >= 5 ? (1 << (phc - 5)) : 1;
    const unsigned* pk = gen_keys(prev, gidp, phc);
    c.par([&](int tid) {
      if (tid == 0) {
        sm->t_m = m; sm->t_phc = phc; sm->t_pcells = pc; sm->t_z = me.z; sm->t_is_child = me.flags & 1; sm->t_has_cmap = (me.flags >> 1) & 1;
        sm->t_hflag = me.hflag; sm->t_enc_lhp = me.enc_lhp; sm->t_csneg = me.csneg; sm->t_c = me.c_val; sm->t_d = me.d_val;
        sm->t_pG = gen_G(prev, gidp, phc);
        sm->t_mask = ws.sgnmask[me.parent] ^ ws.bxor[me.parent];
        const double* p = term_p(tv, m, ti);
        double s = 0; for (int i = 0; i < m; i++) s += p[i];
        sm->t_psq = s * s;
        sm->flag = 0;
        if ((me.flags & 3) == 3) {         // coaligned new child: parent position k <- child row cmap[l], flipped by cs_map[l]
          const unsigned char* cm = tv.cmap + gt * MAXM;
          unsigned flip = 0; int k = 0, l = 0;
          while (k < phc) {
            if (k == me.z) { sm->ksrc[k] = 255; k++; if (k == phc) break; }
            sm->ksrc[k] = cm[l]; if ((me.csneg >> l) & 1u) flip |= (1u << k);
            k++; l++;
          }
          sm->kflip = flip;
        }
      }
      if (tid < m) sm->q[tid] = term_q(tv, m, ti)[tid];
      for (int i = tid; i < nwP; i += c.nthreads()) bmP[i] = 0;
    });
    c.par([&](int tid) { for (int i = tid; i < pc; i += c.nthreads()) { const unsigned k = pk[i]; c.atomic_or(&bmP[k >> 5], 1u << (k & 31)); } });
    int tot;
    bm_prefix(c, bmP, pfP, nwP, &tot);
  }

  // G_p lookup through the rank structure (eval_gs.hpp:94-153 semantics: half storage, conjugate of the opposite cell, 0 when absent)
  MCE_HD cplx lookup(const Group2Sm* sm, const unsigned* bmP, const unsigned short* pfP, int enc_l) const {
    const int phc = sm->t_phc, top = 1 << (phc - 1), rev = (1 << phc) - 1;
    const bool cj = (enc_l & top) != 0;
    const int r = bitmap_rank(bmP, pfP, (unsigned)(cj ? (rev ^ enc_l) : enc_l));
    if (r < 0) return make_cplx(0, 0);
    const cplx v = sm->t_pG[r];
    return cj ? cconj(v) : v;
  }

  // G of one cell of the staged term, flattening.hpp:129-247
  MCE_HDN MCE_NOINLINE cplx eval_cell(Group2Sm* sm, const unsigned* bmP, const unsigned short* pfP, unsigned key) const {
    const int mm = sm->t_m, phc = sm->t_phc;
    double ygi = 0;
    const unsigned hf = sm->t_hflag;
    for (int k = 0; k < mm; k++) if (!((hf >> k) & 1u)) ygi += ((key >> k) & 1u) ? -sm->q[k] : sm->q[k];
    int lp, lm;
    const int phc_mask = (1 << phc) - 1;
    if (!sm->t_is_child) { lp = (int)(key & (unsigned)phc_mask); lm = lp; }
    else {
      const int z = sm->t_z;
      if (!sm->t_has_cmap) {              // insert a zero bit at position z, truncate to phc bits
        const unsigned low = key & ((1u << z) - 1u), high = (z < 31) ? ((key >> z) << (z + 1)) : 0u;
        lp = (int)((z < phc ? (low | high) : key) & (unsigned)phc_mask);
      } else {
        unsigned v = 0;
        for (int k = 0; k < phc; k++) { const unsigned s = sm->ksrc[k]; if (s != 255u) v |= ((key >> s) & 1u) << k; }
        lp = (int)((v ^ sm->kflip) & (unsigned)phc_mask);
        if (z < phc) lp &= ~(1 << z);
      }
      lm = (z < phc) ? (lp | (1 << z)) : lp;
    }
    const cplx gp = lookup(sm, bmP, pfP, lp ^ (int)sm->t_enc_lhp);
    const cplx gm = lookup(sm, bmP, pfP, lm ^ (int)sm->t_enc_lhp);
    cplx g = csub(cdiv(gp, make_cplx(ygi + sm->t_d, sm->t_c)), cdiv(gm, make_cplx(ygi - sm->t_d, sm->t_c)));
    g = cscale(g, sp.gscale);
    if (!*(volatile int*)&sm->flag)        // |G| only matters until one cell is found non-negligible (flat:242-247)
      if ((sm->t_psq * cabs_(g)) > TERM_APPROXIMATION_EPS) sm->flag = 1;
    return g;
  }

  template <class Ctx> MCE_KERNEL_FN void orient(Ctx& c, Group2Sm* sm, int ti, int tj) const {
    const double* Ai = term_A(tv, m, ti, sp.d); const double* Aj = term_A(tv, m, tj, sp.d);
    c.par([&](int tid) { if (tid == 0) sm->sigma = 0; });
    c.par([&](int tid) { if (tid < m) { unsigned b = orient_bit(Ai, Aj, tid, sp.d); if (b) c.atomic_or((unsigned*)&sm->sigma, b); } });
  }

  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int d = sp.d;
    unsigned char* base = c.smem();
    Group2Sm* sm = (Group2Sm*)base;
    cplx* acc = (cplx*)(base + ((sizeof(Group2Sm) + 15) & ~(size_t)15));
    cplx* Gm = acc + HC;
    unsigned* Bk = (unsigned*)(Gm + HC);
    unsigned* bmP = Bk + HC;             // parent-table rank structure of the staged term
    unsigned* bmA = bmP + NW;            // scratch bitmap: parent B_mu / child keys / final keys
    unsigned short* pfP = (unsigned short*)(bmA + NW);
    unsigned short* pfA = pfP + NW + 16 + c.nthreads();   // prefix arrays carry chunk totals behind them
    const int gi = g0 + c.block();
    const int start = grp_start[gi], ncomb = grp_start[gi + 1] - start;
    const int* members = order + start;
    const int gid_out = next.gid_begin[m] + gi;
    const unsigned rev_m = (1u << m) - 1u, top_m = 1u << (m - 1);
    const int nwM = m >= 5 ? (1 << (m - 5)) : 1;

    // ---- B-table of the root (K7) ----
    const int root = members[0];
    const SlotMeta meR = tv.meta[tv.t_begin[m] + root];
    int nB = 0;
    if (meR.flags & 1) {
      if (m <= d) {
        nB = 1 << (m - 1);
        c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) Bk[i] = (unsigned)i; if (tid == 0) sm->owner = -1; });
      } else {
        const unsigned* src; int ncp; unsigned mask;
        parent_B_src(meR.parent, &src, &ncp, &mask);
        const int pbc = meR.pbc, z = meR.z;
        const int nwPb = pbc >= 5 ? (1 << (pbc - 5)) : 1;
        unsigned* bmC = bmP;            // the child-key bitmap borrows bmP (not in use before the first stage_term)
        c.par([&](int tid) {
          for (int i = tid; i < nwPb; i += c.nthreads()) bmA[i] = 0;
          for (int i = tid; i < nwM; i += c.nthreads()) bmC[i] = 0;
          if (tid == 0) sm->owner = -1;
        });
        c.par([&](int tid) { for (int i = tid; i < ncp; i += c.nthreads()) { const unsigned b = src[i] ^ mask; c.atomic_or(&bmA[b >> 5], 1u << (b & 31)); } });
        const unsigned mask_z = 1u << z, hbit = 1u << (pbc - 1), rev_pbc = (pbc >= 32) ? 0xffffffffu : ((1u << pbc) - 1u), mask_low = (1u << z) - 1u;
        const bool coal = m < pbc;
        const unsigned char* cm = tv.cmap + (tv.t_begin[m] + root) * MAXM;
        c.par([&](int tid) {            // pairs (b, b ^ 2^z) both present -> two child sign vectors (ce:261-320), coalesced on the fly (ce:323-366)
          unsigned sel[MAXM]; int nsel = 0;
          if (coal) { unsigned seen = 0; for (int j = 0; j < pbc; j++) { const unsigned ci = cm[j]; if (!((seen >> ci) & 1u)) { seen |= (1u << ci); sel[nsel++] = 1u << j; } } }
          for (int i = tid; i < ncp; i += c.nthreads()) {
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
        nB = c.uniform(sm->cnt);
        c.par([&](int tid) {            // enumerate the set bits: keys in ascending order
          for (int w = tid; w < nwM; w += c.nthreads()) {
            unsigned bits = bmC[w]; int o = pfP[w];
            while (bits) { const int b = MCE_FFS(bits); bits &= bits - 1u; Bk[o++] = (unsigned)(w * 32 + b); }
          }
        });
      }
    } else {
      const unsigned* src; unsigned mask;
      parent_B_src(meR.parent, &src, &nB, &mask);
      c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) Bk[i] = src[i] ^ mask; if (tid == 0) sm->owner = meR.parent; });
    }

    // ---- root table, with re-election when the candidate is negligible (flat:399-489) ----
    int k = 0, cur = root, accepted = 0;
    for (;;) {
      stage_term(c, sm, bmP, pfP, cur);
      c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) acc[i] = eval_cell(sm, bmP, pfP, Bk[i]); });
      if (c.uniform(sm->flag)) { accepted = 1; break; }
      const int lfr = cur;
      if (++k >= ncomb) break;
      cur = members[k];
      const SlotMeta meK = tv.meta[tv.t_begin[m] + cur];
      if (meK.flags & 1) {
        orient(c, sm, lfr, cur);
        unsigned sigma = (unsigned)c.uniform(sm->sigma);
        if (sigma & top_m) sigma ^= rev_m;
        if (sigma)
          c.par([&](int tid) {
            for (int i = tid; i < nB; i += c.nthreads()) Bk[i] ^= sigma;
            if (tid == 0 && sm->owner >= 0) c.atomic_xor(ws.bxor + sm->owner, sigma);
          });
      } else {
        const unsigned* src; unsigned mask;
        parent_B_src(meK.parent, &src, &nB, &mask);
        c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) Bk[i] = src[i] ^ mask; if (tid == 0) sm->owner = meK.parent; });
      }
    }
    if (!accepted) {
      c.par([&](int tid) { if (tid == 0) { alive_flag[gid_out] = 0; next.cells[gid_out] = 0; next.g_m[gid_out] = (unsigned char)m; } });
      return;
    }
    const int rsel = cur;

    // ---- remaining members (flat:491-550) ----
    for (++k; k < ncomb; ++k) {
      const int t = members[k];
      const SlotMeta meT = tv.meta[tv.t_begin[m] + t];
      int own_cells = nB;
      if (!(meT.flags & 1)) own_cells = sp.with_tp ? ws.tpB_cells[meT.parent] : prev.cells[prev.alive[meT.parent]];
      orient(c, sm, rsel, t);
      const unsigned sigma_raw = (unsigned)c.uniform(sm->sigma);
      const bool cj = (sigma_raw & top_m) != 0;
      stage_term(c, sm, bmP, pfP, t);
      if ((meT.flags & 1) || own_cells != nB) {
        const unsigned sigma_n = cj ? (sigma_raw ^ rev_m) : sigma_raw;
        if (!(meT.flags & 1)) c.par([&](int tid) { if (tid == 0) c.atomic_add(diag, 1); });
        c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) Gm[i] = eval_cell(sm, bmP, pfP, Bk[i] ^ sigma_n); });
        if (c.uniform(sm->flag))
          c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) acc[i] = cadd(acc[i], cj ? cconj(Gm[i]) : Gm[i]); });
      } else {
        // old term with its own table: cell i of its B_mu is position i of its (sorted) previous table; add by key (flat:291-314)
        const unsigned* src; int nT; unsigned mask;
        parent_B_src(meT.parent, &src, &nT, &mask);
        if (sp.with_tp) {               // on TP steps B_mu comes from the DCE-TP table, not from the G-table keys: rank over tpB
          c.par([&](int tid) { for (int i = tid; i < nwM; i += c.nthreads()) bmA[i] = 0; });
          c.par([&](int tid) { for (int i = tid; i < nT; i += c.nthreads()) { const unsigned b = src[i]; c.atomic_or(&bmA[b >> 5], 1u << (b & 31)); } });
          int tot; bm_prefix(c, bmA, pfA, nwM, &tot);
        }
        const unsigned* bmT = sp.with_tp ? bmA : bmP; const unsigned short* pfT = sp.with_tp ? pfA : pfP;
        c.par([&](int tid) { for (int i = tid; i < nT; i += c.nthreads()) Gm[i] = eval_cell(sm, bmP, pfP, src[i] ^ mask); });
        if (c.uniform(sm->flag))
          c.par([&](int tid) {
            for (int i = tid; i < nB; i += c.nthreads()) {
              unsigned kq = Bk[i] ^ sigma_raw; bool cjj = false;
              if (kq & top_m) { cjj = true; kq ^= rev_m; }
              const int jj = bitmap_rank(bmT, pfT, kq ^ mask);      // position of the member's cell with key kq
              if (jj >= 0) acc[i] = cadd(acc[i], cjj ? cconj(Gm[jj]) : Gm[jj]);
            }
          });
      }
    }

    // ---- write the surviving term; rank of a key in the bitmap of the final keys = its sorted position (flat:251-252) ----
    c.par([&](int tid) { for (int i = tid; i < nwM; i += c.nthreads()) bmA[i] = 0; });
    c.par([&](int tid) { for (int i = tid; i < nB; i += c.nthreads()) { const unsigned b = Bk[i]; c.atomic_or(&bmA[b >> 5], 1u << (b & 31)); } });
    int tot; bm_prefix(c, bmA, pfA, nwM, &tot);
    unsigned* ko = gen_keys(next, gid_out, m); cplx* Go = gen_G(next, gid_out, m);
    const double* As = term_A(tv, m, rsel, d); const double* ps = term_p(tv, m, rsel); const double* bs = term_b(tv, m, rsel, d);
    double* Ao = gen_A(next, gid_out, m, d); double* po = gen_p(next, gid_out, m); double* bo = gen_b(next, gid_out, d);
    c.par([&](int tid) {
      for (int i = tid; i < nB; i += c.nthreads()) { const unsigned b = Bk[i]; const int pos = bitmap_rank(bmA, pfA, b); ko[pos] = b; Go[pos] = acc[i]; }
      for (int i = tid; i < m * d; i += c.nthreads()) Ao[i] = As[i];
      for (int i = tid; i < m; i += c.nthreads()) po[i] = ps[i];
      for (int i = tid; i < d; i += c.nthreads()) bo[i] = bs[i];
      if (tid == 0) {
        alive_flag[gid_out] = 1; next.cells[gid_out] = nB; next.g_m[gid_out] = (unsigned char)m;
        c.atomic_add_u64((unsigned long long*)(diag + 2), (unsigned long long)nB);
      }
    });
  }
};

}  // namespace mce
#endif
