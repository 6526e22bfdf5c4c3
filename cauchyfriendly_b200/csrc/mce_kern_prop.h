// mce_kern_prop.h -- kernels K1/K3/K4 (+K10, step_first): time propagation, measurement update,
// moment contributions, MU coalignment and the regroup-by-shape of the child terms.
// Reference loops replaced: cauchy_term.hpp:84-310 (msmt_update), 312-401 (eval_g_yei), 435-531
// (time_prop, tp_coalign), 458-472 (normalize_hps), 533-745 (mu_coalign); cauchy_estimator.hpp:307-338
// (cache_moments), 744-779 (regroup), 1179-1208 (step_first), 1312-1394 (shift / deterministic TP).
#ifndef MCE_KERN_PROP_H_
#define MCE_KERN_PROP_H_

#include "mce_exec.h"
#include "mce_types.h"

namespace mce {

// ---------------------------------------------------------------------------------------------
// Per-thread building blocks (straight restatements; dot products left to right, no FMA).
// ---------------------------------------------------------------------------------------------
MCE_HD double dot_lr(const double* x, const double* y, int n) {
  double z = 0.0;
  for (int i = 0; i < n; i++) z += x[i] * y[i];
  return z;
}

// normalize_hps, cauchy_term.hpp:458-472
MCE_HD void normalize_rows(double* A, double* p, double* q, int m, int d, bool set_q) {
  if (set_q) for (int i = 0; i < m; i++) q[i] = p[i];
  for (int i = 0; i < m; i++) {
    double norm1 = 0;
    for (int j = 0; j < d; j++) norm1 += fabs(A[i * d + j]);
    p[i] *= norm1;
    for (int j = 0; j < d; j++) A[i * d + j] /= norm1;
  }
}

// (anti)parallel test of two L1-normalised rows, cauchy_term.hpp:494-504 / 563-575
MCE_HD void coalign_gates(const double* r, const double* c, int d, bool* pos, bool* neg) {
  bool pg = true, ng = true;
  for (int l = 0; l < d; l++) {
    if (pg) pg = fabs(r[l] - c[l]) < COALIGN_EPS;
    if (ng) ng = fabs(r[l] + c[l]) < COALIGN_EPS;
    if (!(pg || ng)) break;
  }
  *pos = pg; *neg = ng;
}

// mu_coalign, cauchy_term.hpp:533-745.  Rows are compacted in place; returns the new shape.
// `gate(j, k, &pos, &neg)` answers whether the L1-normalised rows j < k are parallel / anti-parallel.
template <class Gate>
MCE_HD int mu_coalign_core(double* A, double* p, double* q, int m, int d, unsigned* hflag_io, unsigned char* cmap, unsigned* csneg_out, Gate gate) {
  unsigned F = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);   // bit set: row still unique
  unsigned hf = *hflag_io, csneg = 0;
  const bool any_h = hf != 0;
  for (int j = 0; j < m; j++) cmap[j] = 255;
  int unique_count = 0;
  for (int j = 0; j < m - 1; j++) {
    if (!((F >> j) & 1u)) continue;
    cmap[j] = (unsigned char)unique_count;
    for (int k = j + 1; k < m; k++) {
      if (!((F >> k) & 1u)) continue;
      bool pos, neg;
      gate(j, k, &pos, &neg);
      if (pos) {
        if (any_h) {
          const bool hj = (hf >> j) & 1u, hk = (hf >> k) & 1u;
          if (!hj && !hk) q[j] += q[k];
          else if (hj && !hk) { q[j] = q[k]; hf &= ~(1u << j); }
          else if (!hj && hk) hf &= ~(1u << k);
          else { hf &= ~(1u << k); q[j] += q[k]; }
        } else q[j] += q[k];
        F &= ~(1u << k); p[j] += p[k]; cmap[k] = (unsigned char)unique_count;
      }
      if (neg) {
        if (any_h) {
          const bool hj = (hf >> j) & 1u, hk = (hf >> k) & 1u;
          if (!hj && !hk) q[j] -= q[k];
          else if (hj && !hk) { q[j] = -q[k]; hf &= ~(1u << j); }
          else if (!hj && hk) hf &= ~(1u << k);
          // both H-orthogonal and anti-parallel: the reference exit(1)s here (term:674-691); the flag survives
        } else q[j] -= q[k];
        F &= ~(1u << k); p[j] += p[k]; cmap[k] = (unsigned char)unique_count; csneg |= (1u << k);
      }
    }
    unique_count += 1;
  }
  if ((F >> (m - 1)) & 1u) cmap[m - 1] = (unsigned char)unique_count;
  int new_shape = 0;
  for (int i = 0; i < m; i++) new_shape += (F >> i) & 1u;
  if (new_shape != m) {
    unique_count = 1;
    for (int j = 1; j < m; j++) {
      if ((F >> j) & 1u) {
        if (unique_count < j) {
          for (int l = 0; l < d; l++) A[unique_count * d + l] = A[j * d + l];
          p[unique_count] = p[j]; q[unique_count] = q[j];
        }
        unique_count++;
      }
    }
    if (any_h) {
      unsigned nf = 0; unique_count = 0;
      for (int j = 0; j < m; j++) if ((F >> j) & 1u) nf |= (((hf >> j) & 1u) << unique_count++);
      hf = nf;
    }
  }
  *hflag_io = hf; *csneg_out = csneg;
  return new_shape;
}

MCE_HD int mu_coalign_rows(double* A, double* p, double* q, int m, int d, unsigned* hflag_io, unsigned char* cmap, unsigned* csneg_out) {
  normalize_rows(A, p, q, m, d, true);
  return mu_coalign_core(A, p, q, m, d, hflag_io, cmap, csneg_out, [&](int j, int k, bool* pos, bool* neg) { coalign_gates(A + j * d, A + k * d, d, pos, neg); });
}

// eval_g_yei, cauchy_term.hpp:312-401, for a term that has not been L1-normalised yet.
// y gets 2d doubles: (re_j, im_j) = (-sum_l p_l s_l a_lj, b_j).
MCE_HD cplx eval_g_yei(const double* A, const double* p, const double* b, int m, int d, unsigned hflag, double c_val, double d_val,
                       const double* root_point, bool first_update, int phc, int z, unsigned enc_lhp,
                       const unsigned* pkeys, const cplx* pG, int pcells, double* y,
                       const unsigned* rbm = nullptr, const unsigned short* rpf = nullptr) {
  double tmp[MAXD];
  for (int j = 0; j < d; j++) tmp[j] = 0;
  unsigned signs = 0;
  double ygi = 0;
  for (int l = 0; l < m; l++) {
    const double s = dot_lr(A + l * d, root_point, d) > 0 ? 1.0 : -1.0;
    if (s < 0) signs |= (1u << l);
    const double sc = p[l] * s;
    for (int j = 0; j < d; j++) tmp[j] += sc * A[l * d + j];
    if (!((hflag >> l) & 1u)) ygi += p[l] * s;
  }
  cplx gp, gm;
  if (first_update) { gp = make_cplx(1, 0); gm = make_cplx(1, 0); }
  else {
    int lp, lm;
    parent_keys(signs, m, phc, z, true, nullptr, 0, &lp, &lm);   // the old term has z = m >= phc: no inserted bit
    if (rbm) {                // rank structure of the parent's table: a bit test instead of a binary search
      gp = g_lookup_rank(lp ^ (int)enc_lhp, phc, rbm, rpf, pG);
      gm = g_lookup_rank(lm ^ (int)enc_lhp, phc, rbm, rpf, pG);
    } else {
      gp = g_lookup(lp ^ (int)enc_lhp, phc, pkeys, pG, pcells);
      gm = g_lookup(lm ^ (int)enc_lhp, phc, pkeys, pG, pcells);
    }
  }
  const cplx vp = make_cplx(ygi + d_val, c_val), vm = make_cplx(ygi - d_val, c_val);
  cplx rp, rm;
  if (!cdiv2_fast(gp, vp, gm, vm, &rp, &rm)) { rp = cdiv(gp, vp); rm = cdiv(gm, vm); }   // common case: the six divisions overlap (mce_math.h)
  cplx g = csub(rp, rm);
  g = cscale(g, 1.0 / (2.0 * M_PI));
  for (int j = 0; j < d; j++) { y[2 * j] = -tmp[j]; y[2 * j + 1] = b[j]; }
  return g;
}

// ---------------------------------------------------------------------------------------------
// K1: time propagation + Gamma coalignment, one thread per parent (TP steps only).
// ---------------------------------------------------------------------------------------------
struct KTimeProp {
  StepParams sp; GenView gen; ParentWs ws;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int r = c.block() * c.nthreads() + tid;
      if (r >= gen.n_alive) return;
      const int d = sp.d, gid = gen.alive[r], m0 = gen_m(gen, gid), MS = sp.max_shape;
      const double* A0 = gen_A(gen, gid, m0, d); const double* p0 = gen_p(gen, gid, m0); const double* b0 = gen_b(gen, gid, d);
      double* A = ws.A + (long long)r * MS * d; double* p = ws.p + (long long)r * MS; double* b = ws.b + (long long)r * d;
      // time_prop, cauchy_term.hpp:435-456: A <- A Phi^T, b <- Phi b (+ B u)
      for (int i = 0; i < m0; i++)
        for (int j = 0; j < d; j++) {
          double sum = 0.0;
          for (int k = 0; k < d; k++) sum += A0[i * d + k] * sp.Phi[k + j * d];
          A[i * d + j] = sum;
        }
      for (int i = 0; i < d; i++) {
        double sum = 0.0;
        for (int j = 0; j < d; j++) sum += sp.Phi[i * d + j] * b0[j];
        b[i] = sum;
      }
      if (sp.has_bu) for (int i = 0; i < d; i++) b[i] += 1.0 * sp.bu[i];
      for (int i = 0; i < m0; i++) p[i] = p0[i];
      // tp_coalign, cauchy_term.hpp:475-531
      normalize_rows(A, p, nullptr, m0, d, false);
      int m = m0; unsigned Fg = (1u << sp.npn) - 1u;
      for (int j = 0; j < sp.npn; j++) {
        const double* gr = sp.GammaT + j * d;
        for (int k = 0; k < m0; k++) {
          if (!((Fg >> j) & 1u)) break;
          bool pos, neg;
          coalign_gates(gr, A + k * d, d, &pos, &neg);
          if (pos || neg) { Fg &= ~(1u << j); p[k] += sp.beta[j]; }
        }
      }
      for (int i = 0; i < sp.npn; i++)
        if ((Fg >> i) & 1u) {
          for (int l = 0; l < d; l++) A[m * d + l] = sp.GammaT[i * d + l];
          p[m++] = sp.beta[i];
        }
      ws.m_tp[r] = (unsigned char)m;
    });
  }
};

// ---------------------------------------------------------------------------------------------
// K3+K4, cooperative version: a CTA takes MU_PB parents at a time and keeps their mu rows and all their children in shared
// memory; the work is spread as (parent, row), (parent, child, row) and (parent, child) items, so no thread holds a
// hyperplane array of its own (a thread per (parent, child) keeps ~5 KB in local memory, which turns into HBM traffic: round 1's first kernel).
// Arithmetic and its order are those of the reference (term:84-310), item by item.
// ---------------------------------------------------------------------------------------------
constexpr int MU_PB = 4;      // one warp per parent: 32 * MU_PB threads per CTA
// (child s, element l) items of one parent walked by a lane with stride 32, without integer divisions
struct RowIt {
  int s, l, qd, rd, W;
  MCE_HD RowIt(int lane, int W_) : W(W_) { s = lane / W_; l = lane - s * W_; qd = 32 / W_; rd = 32 - qd * W_; }
  MCE_HD void next() { s += qd; l += rd; if (l >= W) { l -= W; s++; } }
};
struct KMsmtUpdate2 {
  StepParams sp; GenView gen; ParentWs ws; SlotView sl; int ms;
  struct Par { int r, gid, phc, m, pad; unsigned F_int, sgn; double zeta; const double *Ap, *pp, *bp; };
  struct Slot { int valid, t, newm; unsigned hofs, csneg, flags2; };
  static MCE_HD size_t per_parent_doubles(int MT, int d) { return (size_t)(MT + 1) * d + (MT + 1) + (size_t)(MT + 1) * ((size_t)MT * d + 2 * MT + d); }
  static MCE_HD size_t smem_bytes(int MT, int d) {
    return MU_PB * (sizeof(double) * per_parent_doubles(MT, d) + sizeof(Par) + (MT + 1) * (sizeof(Slot) + 2 * MT * sizeof(unsigned))) + 64;
  }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int MT = sl.MT[ms], spp = MT + 1, d = sp.d, NT = c.nthreads();
    const int npar = sl.par_begin[ms + 1] - sl.par_begin[ms];
    const int pb0 = c.block() * MU_PB, npb = (npar - pb0) < MU_PB ? (npar - pb0) : MU_PB;
    // shared-memory carve-up (per parent pb)
    const size_t ppd = per_parent_doubles(MT, d);
    double* dbase = (double*)c.smem();
    Par* par = (Par*)(dbase + MU_PB * ppd);
    Slot* slot = (Slot*)(par + MU_PB);
    unsigned* gates = (unsigned*)(slot + MU_PB * spp);             // [pb][s][2][MT] parallel / anti-parallel masks of row j
    auto MU = [&](int pb) { return dbase + pb * ppd; };              // [(MT+1)][d]
    auto RHO = [&](int pb) { return MU(pb) + (MT + 1) * d; };        // [MT+1]
    auto CA = [&](int pb, int s) { return RHO(pb) + (MT + 1) + (size_t)s * ((size_t)MT * d + 2 * MT + d); };   // [MT][d]
    auto CP = [&](int pb, int s) { return CA(pb, s) + MT * d; };
    auto CQ = [&](int pb, int s) { return CP(pb, s) + MT; };
    auto CB = [&](int pb, int s) { return CQ(pb, s) + MT; };
    auto GP = [&](int pb, int s) { return gates + ((size_t)(pb * spp + s) * 2) * MT; };
    // ---- P0: parent descriptors ----
    c.par([&](int tid) {
      if (tid >= npb) return;
      Par& P = par[tid];
      P.r = sl.par_begin[ms] + pb0 + tid; P.gid = gen.alive[P.r]; P.phc = gen_m(gen, P.gid); P.F_int = 0; P.sgn = 0;
      if (sp.with_tp) { P.m = ws.m_tp[P.r]; P.Ap = ws.A + (long long)P.r * sp.max_shape * d; P.pp = ws.p + (long long)P.r * sp.max_shape; P.bp = ws.b + (long long)P.r * d; }
      else { P.m = P.phc; P.Ap = gen_A(gen, P.gid, P.phc, d); P.pp = gen_p(gen, P.gid, P.phc); P.bp = gen_b(gen, P.gid, d); }
      P.zeta = sp.msmt - dot_lr(sp.H, P.bp, d);
    });
    // ---- P1: mu_l = a_l / (H a_l), rho_l = p_l |H a_l| (term:104-134), one (parent, row) per thread ----
    c.par([&](int tid) {
      const int pb = tid >> 5, lane = tid & 31;
      for (int l = lane; l <= MT && pb < npb; l += 32) {
        Par& P = par[pb];
        if (l > P.m) continue;
        double* mu_l = MU(pb) + l * d;
        if (l == P.m) { RHO(pb)[l] = sp.gamma; for (int i = 0; i < d; i++) mu_l[i] = 0; c.atomic_or(&P.F_int, 1u << l); continue; }
        double row[MAXD];
        for (int i = 0; i < d; i++) row[i] = P.Ap[l * d + i];
        const double H_mu = dot_lr(sp.H, row, d), a = fabs(H_mu);
        if (a < MU_EPS) { RHO(pb)[l] = P.pp[l]; for (int i = 0; i < d; i++) mu_l[i] = row[i]; }
        else {
          const double sc = 1.0 / H_mu;
          for (int i = 0; i < d; i++) mu_l[i] = row[i] * sc;
          RHO(pb)[l] = P.pp[l] * a; c.atomic_or(&P.F_int, 1u << l);
          if (!(H_mu > 0)) c.atomic_or(&P.sgn, 1u << l);
        }
      }
      for (int s = lane; s < spp && pb < npb; s += 32) { Slot& S = slot[pb * spp + s]; S.valid = 0; S.t = 0; S.newm = 0; S.hofs = 0; S.csneg = 0; S.flags2 = 0; }
    });
    // ---- P2: child rows (term:158-211), one (parent, child, row) per thread ----
    c.par([&](int tid) {
      const int pb = tid >> 5, lane = tid & 31;
      for (RowIt w(lane, MT); w.s < spp && pb < npb; w.next()) {
        const int s = w.s, l = w.l, ps = pb * spp + s;
        const Par& P = par[pb];
        const int m = P.m, t = (s == 0) ? m : s - 1;
        if (t > m || (s != 0 && t >= m) || !((P.F_int >> t) & 1u)) continue;        // no such child
        if (l >= m) continue;
        const int _l = l < t ? l : l + 1;
        const double* mu_l = MU(pb) + _l * d; const double* mu_t = MU(pb) + t * d;
        double* ca = CA(pb, s) + l * d;
        CP(pb, s)[l] = RHO(pb)[_l];
        if ((P.F_int >> _l) & 1u) for (int i = 0; i < d; i++) ca[i] = mu_l[i] - mu_t[i];
        else { for (int i = 0; i < d; i++) ca[i] = mu_l[i]; c.atomic_or(&slot[ps].hofs, 1u << l); }
        if (l == 0) {
          double* cb = CB(pb, s);
          for (int i = 0; i < d; i++) cb[i] = P.bp[i] + P.zeta * mu_t[i];
          slot[ps].valid = 1; slot[ps].t = t;
        }
      }
    });
    // ---- P3: moment contribution (est:307-338) + slot meta, one (parent, child) per thread ----
    c.par([&](int tid) {
      // (parent, child) items packed over the CTA's threads: a parent has at most MT + 1 children, so one warp per parent
      // would leave two thirds of its lanes idle in this phase (the heaviest one: two table lookups, two complex divisions)
      for (int it = tid; it < npb * spp; it += NT) {
        const int pb = it / spp, s = it - pb * spp;
        const int ps = pb * spp + s;
        const Par& P = par[pb]; Slot& S = slot[ps];
        const long long ls = (long long)(pb0 + pb) * spp + s, gslot = sl.slot_begin[ms] + ls;
        double* yout = sl.y + gslot * 2 * d;
        if (!S.valid) {
          SlotMeta me; me.newm = 0; me.pbc = (unsigned char)P.m; me.z = 0; me.flags = 0; me.hflag = 0; me.enc_lhp = 0; me.csneg = 0; me.parent = P.r; me.pad_ = 0; me.c_val = 0; me.d_val = 0;
          sl.g[gslot] = make_cplx(0, 0);
          for (int j = 0; j < 2 * d; j++) yout[j] = 0;
          sl.meta[gslot] = me;
          continue;
        }
        const int m = P.m, t = S.t, phc = P.phc;
        unsigned enc_lhp = P.sgn;
        if (phc < m) enc_lhp &= (1u << phc) - 1u;
        const unsigned* pkeys = gen_keys(gen, P.gid, phc); const cplx* pG = gen_G(gen, P.gid, phc);
        const long long rko = sp.max_shape <= 16 ? gen_rk_off(gen, P.gid, phc) : 0;
        sl.g[gslot] = eval_g_yei(CA(pb, s), CP(pb, s), CB(pb, s), m, d, S.hofs, P.zeta, RHO(pb)[t], sp.root_point, false, phc, t, enc_lhp, pkeys, pG,
                                 gen.cells[P.gid], yout, sp.max_shape <= 16 ? gen.rbm + rko : nullptr, sp.max_shape <= 16 ? gen.rpf + rko : nullptr);
        if (s == 0) {
          unsigned e = P.sgn;                                   // parent B ^= enc_sgn_AH, half-normalised (term:229-250)
          if (e & (1u << (m - 1))) e ^= (m >= 32 ? 0xffffffffu : ((1u << m) - 1u));
          ws.sgnmask[P.r] = e; ws.bxor[P.r] = 0;
        }
        S.flags2 = enc_lhp;
        if (sp.skip_post_mu) {
          SlotMeta me; me.newm = (unsigned char)m; me.pbc = (unsigned char)m; me.z = (unsigned char)t; me.flags = (s == 0) ? 0 : 1; me.hflag = S.hofs; me.enc_lhp = enc_lhp;
          me.csneg = 0; me.parent = P.r; me.pad_ = 0; me.c_val = P.zeta; me.d_val = RHO(pb)[t];
          sl.meta[gslot] = me;
        }
      }
    });
    if (sp.skip_post_mu) return;
    // ---- P4: L1 normalisation (normalize_hps, term:458-472), one (parent, child, row) per thread ----
    c.par([&](int tid) {
      const int pb = tid >> 5, lane = tid & 31;
      for (RowIt w(lane, MT); w.s < spp && pb < npb; w.next()) {
        const int s = w.s, l = w.l, ps = pb * spp + s;
        if (!slot[ps].valid || l >= par[pb].m) continue;
        double* ca = CA(pb, s) + l * d;
        CQ(pb, s)[l] = CP(pb, s)[l];
        double norm1 = 0;
        for (int j = 0; j < d; j++) norm1 += fabs(ca[j]);
        CP(pb, s)[l] *= norm1;
        for (int j = 0; j < d; j++) ca[j] /= norm1;
        GP(pb, s)[l] = 0; GP(pb, s)[MT + l] = 0;
      }
    });
    // ---- P5: (anti)parallel gates of every row pair of a new child (term:563-575), one (parent, child, row j) per thread ----
    c.par([&](int tid) {
      const int pb = tid >> 5, lane = tid & 31;
      for (RowIt w(lane, MT); w.s < spp && pb < npb; w.next()) {
        const int s = w.s, j = w.l, ps = pb * spp + s;
        const int m = par[pb].m;
        if (s == 0 || !slot[ps].valid || j >= m - 1) continue;
        const double* A = CA(pb, s);
        unsigned pm = 0, nm = 0;
        for (int k = j + 1; k < m; k++) {
          bool pos, neg;
          coalign_gates(A + j * d, A + k * d, d, &pos, &neg);
          if (pos) pm |= 1u << k;
          if (neg) nm |= 1u << k;
        }
        GP(pb, s)[j] = pm; GP(pb, s)[MT + j] = nm;
      }
    });
    // ---- P6: sequential merge of coaligned rows (mu_coalign, term:533-745), one (parent, child) per thread ----
    c.par([&](int tid) {
      for (int it = tid; it < npb * spp; it += NT) {      // packed like P3
        const int pb = it / spp, s = it - pb * spp;
        const int ps = pb * spp + s;
        Slot& S = slot[ps];
        if (!S.valid) continue;
        const int m = par[pb].m;
        const long long gslot = sl.slot_begin[ms] + (long long)(pb0 + pb) * spp + s;
        int newm = m; unsigned csneg = 0, hofs = S.hofs;
        if (s != 0) {
          const unsigned* gp = GP(pb, s);
          newm = mu_coalign_core(CA(pb, s), CP(pb, s), CQ(pb, s), m, d, &hofs, sl.cmap + gslot * MAXM, &csneg,
                                 [&](int j, int k, bool* pos, bool* neg) { *pos = (gp[j] >> k) & 1u; *neg = (gp[MT + j] >> k) & 1u; });
        }
        S.newm = newm; S.csneg = csneg;
        SlotMeta me; me.newm = (unsigned char)newm; me.pbc = (unsigned char)m; me.z = (unsigned char)S.t; me.flags = (unsigned char)(((s == 0) ? 0 : 1) | ((s != 0 && newm < m) ? 2 : 0));
        me.hflag = hofs; me.enc_lhp = S.flags2; me.csneg = csneg; me.parent = par[pb].r; me.pad_ = 0; me.c_val = par[pb].zeta; me.d_val = RHO(pb)[S.t];
        sl.meta[gslot] = me;
      }
    });
    // ---- P7: store the slots (coalesced over each slot's rows) ----
    c.par([&](int tid) {
      const int pb = tid >> 5, lane = tid & 31;
      for (RowIt w(lane, MT * d); w.s < spp && pb < npb; w.next()) {
        const int s = w.s, e = w.l, ps = pb * spp + s;
        const Slot& S = slot[ps];
        if (!S.valid || e >= S.newm * d) continue;
        const long long ls = (long long)(pb0 + pb) * spp + s;
        sl.A[sl.A_off[ms] + ls * (long long)MT * d + e] = CA(pb, s)[e];
      }
      for (RowIt w(lane, MT); w.s < spp && pb < npb; w.next()) {
        const int s = w.s, l = w.l, ps = pb * spp + s;
        const Slot& S = slot[ps];
        if (!S.valid) continue;
        const long long ls = (long long)(pb0 + pb) * spp + s;
        if (l < S.newm) { sl.p[sl.pq_off[ms] + ls * MT + l] = CP(pb, s)[l]; sl.q[sl.pq_off[ms] + ls * MT + l] = CQ(pb, s)[l]; }
      }
      for (RowIt w(lane, d); w.s < spp && pb < npb; w.next()) {
        const int s = w.s, i = w.l, ps = pb * spp + s;
        if (!slot[ps].valid) continue;
        sl.b[(sl.slot_begin[ms] + (long long)(pb0 + pb) * spp + s) * d + i] = CB(pb, s)[i];
      }
    });
  }
};

// ---------------------------------------------------------------------------------------------
// Moments: fz is summed in the reference's serial order (bit-identical to NUM_CPUS=1: the running
// normaliser 1/(2 pi Re fz) scales every new G and hence every term-approximation decision);
// mean / covariance sums use a fixed-shape two-level reduction (deterministic; parity bar 1e-9).
// ---------------------------------------------------------------------------------------------
// Serial-order sums of all 1 + d + d*d complex moment accumulators.  The addends are computed exactly as
// est:318-325 does -- fz += g; mean_j += g*y_j; cov_jk -= (g*y_j)*y_k -- and every real accumulator adds them for
// slot 0, 1, 2, ... in slot order, the order of the reference's cache_moments loop (parents in shape/index order,
// each followed by its children).  Unused slots hold g = y = 0 and leave the accumulators unchanged, so the sums
// are bit-identical to the NUM_CPUS = 1 reference.
// Layout: block q owns ONE complex quantity; lanes 0 and 1 of warp 0 keep its running (re, im) sums and walk the slots
// serially: that dependent DADD chain (8 cycles per slot on B200) is the critical path.  The other warps run a software
// pipeline ahead of it: cp.async stages the columns the quantity needs (g, y_j, y_k: 48-byte rows, conflict-free for
// 16-byte shared loads) of tile t+4 while the addends of tile t+1 are computed from shared memory; only warps that do not
// share warp 0's scheduler partition (warp id % 4 != 0) do fp64 work.  Large tiles amortise the per-phase barrier.
constexpr int MOM_QB = 1, MOM_TILE = 768, MOM_RAW = 4, MOM_W = 6;
struct KMomentsSerial {
  const cplx* g; const double* y; long long n; int d; double* out /*[2*(1+d+d*d)]*/;
  int dbg = 0;      // tools/ubench/mom_bench.cu only: bit 0 skips the chain, bit 1 the addend production, bit 2 the staging copies
  static MCE_HD size_t smem_bytes(int) { return sizeof(double) * (2 + 2 * MOM_TILE * 2 + MOM_RAW * MOM_TILE * MOM_W) + 64; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int nq = 1 + d + d * d, qbase = c.block(), NA = 2, W = MOM_W;
    const int q = qbase;                              // 0: fz = sum g; 1..d: sum g y_j; d+1..: -sum (g y_j) y_k
    const int j = q == 0 ? 0 : (q <= d ? q - 1 : (q - 1 - d) / d), k = q <= d ? 0 : (q - 1 - d) % d;
    const int ncol = q == 0 ? 1 : (q <= d ? 2 : 3);
    double* raw = (double*)c.smem();                  // [MOM_RAW][MOM_TILE][W]   (g.re, g.im, y_j.re, y_j.im, y_k.re, y_k.im)
    double* buf = raw + MOM_RAW * MOM_TILE * W;       // [2][MOM_TILE][NA]  addends
    double* accs = buf + 2 * MOM_TILE * NA;
    const long long ntiles = (n + MOM_TILE - 1) / MOM_TILE;
    auto tile_cnt = [&](long long t) { const long long s0 = t * MOM_TILE; return (int)((n - s0) < MOM_TILE ? (n - s0) : MOM_TILE); };
    auto stage = [&](long long t, int lane, int nlanes) {            // HBM -> raw[t % MOM_RAW], asynchronously, 16 bytes per copy
      const long long s0 = t * MOM_TILE; const int cnt = tile_cnt(t);
      double* rt = raw + (t % MOM_RAW) * MOM_TILE * W;
      const cplx* gs = g + s0; const cplx* ys = (const cplx*)(y + s0 * 2 * d);   // y rows are d (re, im) pairs
      for (int r = lane; r < cnt; r += nlanes) c.cp_async16(rt + r * W, gs + r);
      if (ncol > 1) for (int r = lane; r < cnt; r += nlanes) c.cp_async16(rt + r * W + 2, ys + (long long)r * d + j);
      if (ncol > 2) for (int r = lane; r < cnt; r += nlanes) c.cp_async16(rt + r * W + 4, ys + (long long)r * d + k);
    };
    auto produce = [&](long long t, int lane, int nlanes) {          // raw[t % MOM_RAW] -> buf[t & 1]
      const int cnt = tile_cnt(t);
      const double* rt = raw + (t % MOM_RAW) * MOM_TILE * W;
      double* bt = buf + (t & 1) * MOM_TILE * NA;
      for (int sidx = lane; sidx < cnt; sidx += nlanes) {
        cplx w = make_cplx(0, 0);
        if (q < nq) {
          const double* row = rt + sidx * W;
          const cplx gv = make_cplx(row[0], row[1]);
          if (q == 0) w = gv;
          else {
            w = cmul(gv, make_cplx(row[2], row[3]));
            if (q > d) { w = cmul(w, make_cplx(row[4], row[5])); w.re = -w.re; w.im = -w.im; }
          }
        }
        bt[sidx * NA] = w.re; bt[sidx * NA + 1] = w.im;
      }
    };
    const int nstage = c.nthreads() - 32;
    auto prod_lane = [&](int tid, int* lane, int* nlanes) {       // producer index among warps with id % 4 != 0, or -1
      const int w = tid >> 5;
      *nlanes = ((c.nthreads() >> 5) - ((c.nthreads() >> 5) + 3) / 4) * 32;
      *lane = (w & 3) ? ((w - 1 - (w >> 2)) * 32 + (tid & 31)) : -1;
    };
    // prologue: tiles 0..3 requested (one cp.async group each), tiles 0 and 1 landed, addends of tile 0 ready
    c.par([&](int tid) {
      if (tid < NA) accs[tid] = 0;
      if (tid >= 32) {
        for (int t0 = 0; t0 < MOM_RAW; t0++) { if (t0 < ntiles) stage(t0, tid - 32, nstage); c.cp_async_commit(); }
        c.cp_async_wait_pending2();
      }
    });
    c.par([&](int tid) {
      int pl, pn; prod_lane(tid, &pl, &pn);
      if (pl >= 0 && ntiles > 0) produce(0, pl, pn);
    });
    for (long long t = 0; t < ntiles; t++) {
      c.par([&](int tid) {
        if (tid < NA && !(dbg & 1)) {
          const int cnt = tile_cnt(t);
          const double* bt = buf + (t & 1) * MOM_TILE * NA + tid;
          double acc = accs[tid];
          int sidx = 0;
          if (cnt >= 8) {      // register double-buffering: the next 8 addends are loaded, THEN the current 8 are added
            double v0 = bt[0 * NA], v1 = bt[1 * NA], v2 = bt[2 * NA], v3 = bt[3 * NA], v4 = bt[4 * NA], v5 = bt[5 * NA], v6 = bt[6 * NA], v7 = bt[7 * NA];
            for (sidx = 8; sidx + 8 <= cnt; sidx += 8) {
              const double w0 = bt[(sidx + 0) * NA], w1 = bt[(sidx + 1) * NA], w2 = bt[(sidx + 2) * NA], w3 = bt[(sidx + 3) * NA];
              const double w4 = bt[(sidx + 4) * NA], w5 = bt[(sidx + 5) * NA], w6 = bt[(sidx + 6) * NA], w7 = bt[(sidx + 7) * NA];
              MCE_SCHED_FENCE();   // keeps every load a full batch (>= 64 cycles) ahead of its use: the chain never waits on shared memory
              acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
              v0 = w0; v1 = w1; v2 = w2; v3 = w3; v4 = w4; v5 = w5; v6 = w6; v7 = w7;
            }
            acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
          }
          for (; sidx < cnt; sidx++) acc += bt[sidx * NA];
          accs[tid] = acc;
        }
        // tile t+1: raw landed a phase ago -> addends; tile t+4: start its copies (raw[(t+4) % 4] == raw[t % 4] was last read by
        // produce(t) in the previous phase); then wait until at most the two youngest groups (t+3, t+4) are pending, i.e. tile
        // t+2 has landed: every copy has two full phases to arrive.
        int pl, pn; prod_lane(tid, &pl, &pn);
        if (pl >= 0 && t + 1 < ntiles && !(dbg & 2)) produce(t + 1, pl, pn);
        if (tid >= 32 && !(dbg & 4)) { if (t + MOM_RAW < ntiles) stage(t + MOM_RAW, tid - 32, nstage); c.cp_async_commit(); c.cp_async_wait_pending2(); }
      });
    }
    c.par([&](int tid) { if (tid < NA && q < nq) out[q * 2 + tid] = accs[tid]; });
  }
};

// ---------------------------------------------------------------------------------------------
// ONE serial-order sum as a parallel scan, bit-identical to  acc = 0; for (k) acc += x[k]  (used for Re fz of a partitioned estimator:
// the only moment sum that feeds back into the filter, G_SCALE_FACTOR = 1 / (2 pi Re fz), est:349).
// While a running sum s stays inside one binade, s <- fl(s + a) is integer arithmetic on its 53-bit significand S in units of ulp(s): the exact
// a / ulp = V + f rounds to V, V + 1, or -- on a tie, f = 1/2 -- to whichever makes the RESULT even (IEEE round-to-nearest-even).  So an addend
// is a map  S -> S + d[S & 1]  with two precomputed integers (d[0] == d[1] unless it is a tie), and such maps compose associatively:
//     (g o f)[p] = f[p] + g[(p + f[p]) & 1].
// A tile of addends is reduced by a block scan of maps; every prefix value is then checked to lie in [2^52 + 1, 2^53 - 2] (then every exact
// partial sum was inside the binade and the integer model held).  At the first element that fails -- the sum crosses a binade or zero, a
// non-finite or oversized addend, a zero / subnormal running sum -- the valid prefix is kept, a few elements are added by the literal loop,
// and the scan restarts behind them with the new binade.  Chains that hover around zero restart at every other element (Im fz does: that is
// why the full set of moment sums stays a dependent chain); Re fz changes binade once per ~1000 addends.
// ---------------------------------------------------------------------------------------------
struct MomMap { long long d0, d1; };
MCE_HD MomMap mom_identity() { MomMap m; m.d0 = 0; m.d1 = 0; return m; }
MCE_HD MomMap mom_compose(const MomMap& f, const MomMap& g) {      // first f, then g
  MomMap h;
  h.d0 = f.d0 + ((f.d0 & 1) ? g.d1 : g.d0);
  h.d1 = f.d1 + (((f.d1 + 1) & 1) ? g.d1 : g.d0);
  return h;
}
MCE_HD long long mom_apply(const MomMap& f, long long S) { return S + ((S & 1) ? f.d1 : f.d0); }
struct MomState { long long S; int E, sign, ok; double scale; };
MCE_HD MomState mom_state(double s) {
  union { double d; unsigned long long u; } v; v.d = s;
  MomState st; st.E = (int)((v.u >> 52) & 0x7ffull); st.sign = (int)(v.u >> 63);
  st.S = (long long)((v.u & 0xfffffffffffffull) | (1ull << 52));
  st.ok = (st.E >= 64 && st.E <= 2046) ? 1 : 0;                  // normal, finite and 1 / ulp(s) representable (zero, tiny, inf, nan take the literal loop)
  union { double d; unsigned long long u; } sc; sc.u = ((unsigned long long)(st.ok ? 2098 - st.E : 1023) << 52) | ((unsigned long long)st.sign << 63);   // +-2^(1075 - E)
  st.scale = sc.d;
  return st;
}
MCE_HD double mom_value(const MomState& st, long long S) {       // S in [2^52, 2^53)
  union { double d; unsigned long long u; } v;
  v.u = ((unsigned long long)st.sign << 63) | ((unsigned long long)st.E << 52) | ((unsigned long long)S & 0xfffffffffffffull);
  return v.d;
}
// The map of addend `a` for a running sum in the binade / sign of `st`; false when the integer model cannot hold for it.
// q = a / ulp(s) is an exact scaling by a power of two (st.scale); rint(q) is the nearest integer with ties to even, which is the increment for
// an EVEN significand (the result must be even on a tie); for an odd one the tie goes to the other neighbour.  |q| >= 2^52 (the sum would leave the
// binade), infinities and NaNs fail the single comparison.  A subnormal product only arises for |q| << 1/2 and rounds to 0 either way.
MCE_HD bool mom_classify(double a, const MomState& st, MomMap* m) {
  const double q = a * st.scale;                                  // signed so that it ADDS to the significand
  *m = mom_identity();
  if (!(fabs(q) < 4503599627370496.0)) return false;              // 2^52
  const double d0 = rint(q), r = q - d0;                          // r is exact
  const long long D0 = (long long)d0;
  m->d0 = D0;
  m->d1 = (fabs(r) == 0.5) ? D0 + (r > 0 ? 1 : -1) : D0;
  return true;
}
MCE_HD bool mom_in_binade(long long S) { return S >= (1ll << 52) + 1 && S <= (1ll << 53) - 2; }

constexpr int SS_SERIAL = 32;           // elements added by the literal loop behind a failed check ...
constexpr int SS_SERIAL_MAX = 4096;     // ... doubling up to this while the scan keeps failing early
constexpr int SS_E = 8;                 // consecutive elements per thread of one tile
constexpr int SS_NT = 1024;             // threads of the scan kernels: a tile is SS_NT * SS_E = 8192 addends

// Tiles in parallel.  The map of a whole tile only depends on the binade and sign of the running sum while the tile is added, and those can be GUESSED: the plain
// (unordered) sum of the tiles in front is within rounding noise of the true running sum.  KSumTileSums adds every tile up, KSumTileMaps composes every tile's map
// under its guess and records, for both parities of the significand at the tile's start, the total offset and the smallest / largest prefix offset.  KSumScan then
// walks the tiles in order with the EXACT running sum: a tile whose guess was right and whose prefixes all stay inside the binade (two comparisons) is applied in
// O(1); any other tile -- the sum crosses a binade inside it, the guess was off, special values -- is added the slow way.  Nothing is ever assumed: a summary is only
// used when the exact state proves the integer model held for every prefix of the tile, so the result is the serial chain's, bit for bit.
struct SumTile { long long d[2], mn[2], mx[2]; int E, sign, ok, pad; };
struct SumRange { long long mn[2], mx[2]; };
struct KSumTileSums {                   // tsum[t] = plain sum of tile t (order irrelevant: it only feeds the guess)
  const double* x; int xs /* stride in doubles: 2 = real parts of a complex array, 1 = packed reals */; long long n; double* tsum;
  static MCE_HD size_t smem_bytes(int nthreads) { return sizeof(double) * (size_t)nthreads + 64; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    double* part = (double*)c.smem();
    const long long base = (long long)c.block() * c.nthreads() * SS_E;
    c.par([&](int tid) {
      double acc = 0;
      for (int e = 0; e < SS_E; e++) { const long long k = base + (long long)tid * SS_E + e; if (k < n) acc += x[k * xs]; }
      part[tid] = acc;
    });
    for (int w = c.nthreads() >> 1; w >= 1; w >>= 1) c.par([&](int tid) { if (tid < w) part[tid] += part[tid + w]; });
    c.par([&](int tid) { if (tid == 0) tsum[c.block()] = part[0]; });
  }
};
struct KSumTileMaps {
  const double* x; int xs; long long n; const double* tsum; SumTile* tiles;
  static MCE_HD size_t smem_bytes(int nthreads) { return sizeof(MomMap) * ((size_t)nthreads + 32) + sizeof(SumRange) * ((size_t)nthreads + 32) + sizeof(double) * ((size_t)nthreads + 2) + sizeof(int) * 4 + 64; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int NT = c.nthreads(), t = c.block();
    MomMap* maps = (MomMap*)c.smem();                 // [NT] + [32] scan scratch
    SumRange* rng = (SumRange*)(maps + NT + 32);      // [NT] + [32]
    double* part = (double*)(rng + NT + 32);          // [NT] partial sums of the tiles in front, [NT] the guess
    int* ctl = (int*)(part + NT + 2);
    const long long base = (long long)t * NT * SS_E;
    c.par([&](int tid) {
      double acc = 0;
      for (int u = tid; u < t; u += NT) acc += tsum[u];
      part[tid] = acc;
      if (tid == 0) ctl[0] = 0;
    });
    for (int w = NT >> 1; w >= 1; w >>= 1) c.par([&](int tid) { if (tid < w) part[tid] += part[tid + w]; });
    const double guess = c.uniform(part[0]);
    const MomState st = mom_state(guess);
    if (!st.ok) { c.par([&](int tid) { if (tid == 0) { SumTile T; memset(&T, 0, sizeof(T)); tiles[t] = T; } }); return; }
    c.par([&](int tid) {
      MomMap m = mom_identity(); int bad = 0;
      for (int e = 0; e < SS_E; e++) {
        const long long k = base + (long long)tid * SS_E + e;
        if (k >= n) break;
        MomMap me = mom_identity();
        if (!mom_classify(x[k * xs], st, &me)) bad = 1;
        m = mom_compose(m, me);
      }
      maps[tid] = m;
      if (bad) c.atomic_or((unsigned*)&ctl[0], 1u);
    });
    c.block_scan(maps, maps + NT, [](const MomMap& f, const MomMap& g) { return mom_compose(f, g); });
    c.par([&](int tid) {                               // prefix offsets for both parities of the significand at the tile's start
      const MomMap Q = tid == 0 ? mom_identity() : maps[tid - 1];
      long long o[2] = {Q.d0, Q.d1};
      SumRange r; r.mn[0] = r.mn[1] = 0x7fffffffffffffffll; r.mx[0] = r.mx[1] = -0x7fffffffffffffffll - 1;
      for (int e = 0; e < SS_E; e++) {
        const long long k = base + (long long)tid * SS_E + e;
        if (k >= n) break;
        MomMap me = mom_identity();
        mom_classify(x[k * xs], st, &me);
        for (int pi = 0; pi < 2; pi++) {
          o[pi] += ((pi + o[pi]) & 1) ? me.d1 : me.d0;
          if (o[pi] < r.mn[pi]) r.mn[pi] = o[pi];
          if (o[pi] > r.mx[pi]) r.mx[pi] = o[pi];
        }
      }
      rng[tid] = r;
    });
    c.block_scan(rng, rng + NT, [](const SumRange& a, const SumRange& b) {
      SumRange r;
      for (int pi = 0; pi < 2; pi++) { r.mn[pi] = a.mn[pi] < b.mn[pi] ? a.mn[pi] : b.mn[pi]; r.mx[pi] = a.mx[pi] > b.mx[pi] ? a.mx[pi] : b.mx[pi]; }
      return r;
    });
    c.par([&](int tid) {
      if (tid != 0) return;
      SumTile T;
      T.d[0] = maps[NT - 1].d0; T.d[1] = maps[NT - 1].d1;
      for (int pi = 0; pi < 2; pi++) { T.mn[pi] = rng[NT - 1].mn[pi]; T.mx[pi] = rng[NT - 1].mx[pi]; }
      T.E = st.E; T.sign = st.sign; T.ok = ctl[0] ? 0 : 1; T.pad = 0;
      tiles[t] = T;
    });
  }
};
struct KSumScan {                       // ONE block; out[0] = x[0].re + x[1].re + ... (n addends) in order; out[1] = restarts, out[2] = tiles taken from their summaries (statistics)
  const double* x; int xs; long long n; double* out;
  const SumTile* tiles = nullptr;       // summaries of KSumTileMaps (same tile size), or null: every tile the slow way
  static MCE_HD size_t smem_bytes(int nthreads) {
    return sizeof(double) * (2 * (size_t)nthreads * SS_E + 2) + sizeof(MomMap) * ((size_t)nthreads + 32) + sizeof(long long) * ((size_t)nthreads + 2) + sizeof(int) * 8 + 64;
  }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int NT = c.nthreads(), TILE = NT * SS_E;
    double* vals = (double*)c.smem();                 // [2][TILE] addends of this tile and the next
    double* sv = vals + 2 * (size_t)TILE;             // [0] running sum
    MomMap* maps = (MomMap*)(sv + 2);                 // [NT] thread maps (inclusive scan in place), then [32] scan scratch
    long long* Ss = (long long*)(maps + NT + 32);     // [NT] significand in front of every thread's first element
    long long* tnext = Ss + NT;                       // [0] first tile the fast path could not take
    int* ctl = (int*)(Ss + NT + 2);                   // [0] first failing element, [1] done, [2] next start, [3] restarts, [4] tiles applied from their summaries
    const long long ntiles = (n + TILE - 1) / TILE;
    c.par([&](int tid) {
      if (tid == 0) { sv[0] = 0; ctl[3] = 0; ctl[4] = 0; ctl[5] = SS_SERIAL / 2; }
      if (!tiles) for (int e = 0; e < SS_E; e++) { const long long k = (long long)tid * SS_E + e; vals[tid * SS_E + e] = (k < n) ? x[k * xs] : 0.0; }
    });
    for (long long t = 0; t < ntiles; t++) {
      if (tiles) {
        // every consecutive tile whose summary the exact running sum validates is applied in O(1) by one thread; the first one that is not is added below
        c.par([&](int tid) {
          if (tid != 0) return;
          double s = sv[0]; long long tt = t;
          const long long LO = (1ll << 52) + 1, HI = (1ll << 53) - 2;
          while (tt < ntiles) {
            const SumTile T = tiles[tt];
            const MomState st = mom_state(s);
            if (!(T.ok && st.ok && st.E == T.E && st.sign == T.sign)) break;
            const int pi = (int)(st.S & 1);
            if (!(T.mn[pi] >= LO - st.S && T.mx[pi] <= HI - st.S)) break;      // written so that garbage offsets cannot overflow the comparison
            s = mom_value(st, st.S + T.d[pi]);
            tt++;
          }
          sv[0] = s; tnext[0] = tt; ctl[4] += (int)(tt - t);
        });
        t = c.uniform(tnext[0]);
        if (t >= ntiles) break;
        c.par([&](int tid) {
          double* buf = vals + (t & 1) * (size_t)TILE;
          for (int e = 0; e < SS_E; e++) { const long long k = t * TILE + (long long)tid * SS_E + e; buf[tid * SS_E + e] = (k < n) ? x[k * xs] : 0.0; }
        });
      }
      const int cnt = (int)((n - t * TILE) < TILE ? (n - t * TILE) : TILE);
      const double* a = vals + (t & 1) * (size_t)TILE;
      int pos = 0;
      for (;;) {
        const double s = c.uniform(sv[0]);
        const MomState st = mom_state(s);
        union { double dd; unsigned long long u; } sb; sb.dd = s;
        const bool zero = sb.u == 0ull;               // +0: addends of +-0 leave it unchanged ((+0) + (-0) = +0)
        // A: the maps of the own elements, composed; the next tile's loads are issued first and stored last, so they fly meanwhile
        c.par([&](int tid) {
          double nx[SS_E];
          const bool pre = !tiles && pos == 0 && t + 1 < ntiles;
          if (pre) for (int e = 0; e < SS_E; e++) { const long long k = (t + 1) * TILE + (long long)tid * SS_E + e; nx[e] = (k < n) ? x[k * xs] : 0.0; }
          MomMap m = mom_identity(); int bad = 0;
          for (int e = 0; e < SS_E; e++) {
            const int k = tid * SS_E + e;
            if (k < pos || k >= cnt) continue;
            MomMap me = mom_identity();
            if (zero) bad |= !(a[k] == 0.0);
            else if (!st.ok || !mom_classify(a[k], st, &me)) bad = 1;
            m = mom_compose(m, me);
          }
          maps[tid] = m;
          if (bad) Ss[tid] = -1; else Ss[tid] = 0;
          if (tid == 0) ctl[0] = 0x7fffffff;
          if (pre) { double* nb = vals + ((t + 1) & 1) * (size_t)TILE; for (int e = 0; e < SS_E; e++) nb[tid * SS_E + e] = nx[e]; }
        });
        c.block_scan(maps, maps + NT, [](const MomMap& f, const MomMap& g) { return mom_compose(f, g); });
        // B: every prefix value must stay inside the binade; the first element that fails is the restart point
        c.par([&](int tid) {
          const bool bad = Ss[tid] < 0;
          long long S = tid == 0 ? st.S : mom_apply(maps[tid - 1], st.S);
          Ss[tid] = S;
          for (int e = 0; e < SS_E; e++) {
            const int k = tid * SS_E + e;
            if (k < pos || k >= cnt) continue;
            bool fail;
            if (zero) fail = !(a[k] == 0.0);
            else {
              MomMap me = mom_identity();
              fail = !st.ok || !mom_classify(a[k], st, &me);
              S = mom_apply(me, S);
              fail = fail || !mom_in_binade(S);
            }
            if (fail) { c.atomic_min(&ctl[0], k); break; }
          }
          (void)bad;
        });
        // C: the new running sum, or a few literal additions behind the failing element
        c.par([&](int tid) {
          if (tid != 0) return;
          const int v = ctl[0];
          if (v == 0x7fffffff) { if (!zero && cnt > pos) sv[0] = mom_value(st, mom_apply(maps[NT - 1], st.S)); ctl[1] = 1; return; }
          double acc = s;
          if (!zero && v > pos) {                         // the valid prefix: walk the owning thread's elements up to v - 1
            const int tv = v / SS_E;
            long long S = Ss[tv];
            for (int k = tv * SS_E; k < v; k++) { if (k < pos) continue; MomMap me = mom_identity(); mom_classify(a[k], st, &me); S = mom_apply(me, S); }
            acc = mom_value(st, S);
          }
          // Crossings come in clusters (the first thousands of addends of a real chain change binade every few elements): when the scan got nowhere, the literal
          // run behind the failure doubles (5 ns per addend beats a 10 us scan pass that dies after a handful); a scan that carried a long stretch resets it.
          int L = ctl[5];
          L = (v - pos < 256) ? (L * 2 < SS_SERIAL_MAX ? L * 2 : SS_SERIAL_MAX) : SS_SERIAL;
          ctl[5] = L;
          const int e = v + L < cnt ? v + L : cnt;
          for (int i = v; i < e; i++) acc += a[i];
          sv[0] = acc; ctl[2] = e; ctl[1] = (e >= cnt) ? 1 : 0; ctl[3] += 1;
        });
        if (c.uniform(ctl[1])) break;
        pos = c.uniform(ctl[2]);
      }
    }
    c.par([&](int tid) { if (tid == 0) { out[0] = sv[0]; out[1] = (double)ctl[3]; out[2] = (double)ctl[4]; } });
  }
};

constexpr int MOM_CHUNK = 4096;   // slots per block of the partial-moment reduction
struct KMomentsPartial {          // partial[block][2*(fz + d + d*d)] = sum over the block's slots of ([g,] g*y_j, -g*y_j*y_k)
  const cplx* g; const double* y; long long n; int d; double* partial;
  int fz = 0;                     // 1: the normalisation sum (g itself) is quantity 0, so that KMomentsFinal writes the whole moment vector
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int nq = fz + d + d * d;
    double* sm = (double*)c.smem();                 // [nthreads][2] scratch per quantity
    const long long lo = (long long)c.block() * MOM_CHUNK;
    const long long hi = lo + MOM_CHUNK < n ? lo + MOM_CHUNK : n;
    for (int qn = 0; qn < nq; qn++) {
      c.par([&](int tid) {
        double ar = 0, ai = 0;
        for (long long i = lo + tid; i < hi; i += c.nthreads()) {
          const cplx gv = g[i];
          const double* yy = y + i * 2 * d;
          cplx v;
          const int qm = qn - fz;
          if (qm < 0) v = gv;
          else if (qm < d) v = cmul(gv, make_cplx(yy[2 * qm], yy[2 * qm + 1]));
          else {
            const int j = (qm - d) / d, k = (qm - d) % d;
            v = cmul(cmul(gv, make_cplx(yy[2 * j], yy[2 * j + 1])), make_cplx(yy[2 * k], yy[2 * k + 1]));
            v.re = -v.re; v.im = -v.im;
          }
          ar += v.re; ai += v.im;
        }
        sm[2 * tid] = ar; sm[2 * tid + 1] = ai;
      });
      for (int stride = c.nthreads() / 2; stride > 0; stride >>= 1)
        c.par([&](int tid) { if (tid < stride) { sm[2 * tid] += sm[2 * (tid + stride)]; sm[2 * tid + 1] += sm[2 * (tid + stride) + 1]; } });
      c.par([&](int tid) { if (tid == 0) { partial[((long long)c.block() * nq + qn) * 2] = sm[0]; partial[((long long)c.block() * nq + qn) * 2 + 1] = sm[1]; } });
    }
  }
};
struct KMomentsFinal {            // out[2*nq] = ordered sum of the block partials
  const double* partial; int nblocks_in; int nq; double* out;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int q2 = c.block() * c.nthreads() + tid;
      if (q2 >= 2 * nq) return;
      double acc = 0;
      for (int b = 0; b < nblocks_in; b++) acc += partial[(long long)b * 2 * nq + q2];
      out[q2] = acc;
    });
  }
};

// ---------------------------------------------------------------------------------------------
// Regroup by new shape (cauchy_estimator.hpp:744-779): canonical rank of every slot inside its new
// shape = [old terms in (old shape, parent) order] ++ [children in (old shape, parent, t) order].
// Pass 1 counts per chunk, pass 2 scans the chunk counts per (kind, shape) bin, pass 3 scatters.
// ---------------------------------------------------------------------------------------------
constexpr int RANK_CHUNK = 256;
MCE_HD int slot_region(const SlotView& sl, long long slot) {
  int m = 0;
  for (int k = 1; k < NSHAPE; k++) if (sl.slot_begin[k] <= slot && slot < sl.slot_begin[k + 1]) m = k;
  return m;
}
struct KRankCount {     // counts[(kind*NSHAPE + shape) * nchunks + chunk]
  SlotView sl; int nchunks; int* counts;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    int* hist = (int*)c.smem();      // [2*NSHAPE]
    c.par([&](int tid) { for (int i = tid; i < 2 * NSHAPE; i += c.nthreads()) hist[i] = 0; });
    c.par([&](int tid) {
      const long long slot = (long long)c.block() * RANK_CHUNK + tid;
      if (tid >= RANK_CHUNK || slot >= sl.n_slots) return;
      const SlotMeta& me = sl.meta[slot];
      if (me.newm) c.atomic_add(&hist[(me.flags & 1) * NSHAPE + me.newm], 1);
    });
    c.par([&](int tid) { for (int i = tid; i < 2 * NSHAPE; i += c.nthreads()) counts[(long long)i * nchunks + c.block()] = hist[i]; });
  }
};
struct KRankScan {      // exclusive scan over chunks for every bin; one block per bin: segment sums, scan of the sums, segment scans
  int nchunks; int* counts; int* totals;
  static MCE_HD size_t smem_bytes(int nthreads) { return sizeof(int) * (nthreads + 1); }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    int* part = (int*)c.smem();
    int* cc = counts + (long long)c.block() * nchunks;
    const int NT = c.nthreads(), seg = (nchunks + NT - 1) / NT;
    c.par([&](int tid) {
      const int lo = tid * seg, hi = lo + seg < nchunks ? lo + seg : nchunks;
      int acc = 0;
      for (int i = lo; i < hi; i++) acc += cc[i];
      part[tid] = acc;
    });
    c.par([&](int tid) {
      if (tid != 0) return;
      int acc = 0;
      for (int t = 0; t < NT; t++) { const int v = part[t]; part[t] = acc; acc += v; }
      totals[c.block()] = acc;
    });
    c.par([&](int tid) {
      const int lo = tid * seg, hi = lo + seg < nchunks ? lo + seg : nchunks;
      int acc = part[tid];
      for (int i = lo; i < hi; i++) { const int v = cc[i]; cc[i] = acc; acc += v; }
    });
  }
};
// One block per chunk of RANK_CHUNK slots.  Phase 1: the bin (kind, new shape) of every slot.  Phase 2: the canonical rank of
// every slot = chunk base of its bin + number of earlier slots of the chunk in the same bin; the small per-term fields
// (meta, b, coalignment map) move here.  Phase 3: the hyperplane payload moves as one flattened (slot, element) loop, so
// every thread has many independent loads in flight (the copy is pure HBM traffic).
struct RegroupDesc { int rank, m; long long a_src, a_dst, pq_src, pq_dst; };
struct KRegroup {
  StepParams sp; SlotView sl; TermView tv; int nchunks; const int* counts; long long* slot_of_term;
  static MCE_HD size_t smem_bytes() { return (sizeof(RegroupDesc) + sizeof(int)) * RANK_CHUNK + 16; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    RegroupDesc* desc = (RegroupDesc*)c.smem();            // [RANK_CHUNK]
    int* bin = (int*)(desc + RANK_CHUNK);                  // [RANK_CHUNK] kind * NSHAPE + new shape, -1 = unused slot
    const int d = sp.d;
    const long long slot0 = (long long)c.block() * RANK_CHUNK;
    c.par([&](int tid) {
      for (int k = tid; k < RANK_CHUNK; k += c.nthreads()) {
        const long long slot = slot0 + k;
        int b = -1;
        if (slot < sl.n_slots) { const SlotMeta& me = sl.meta[slot]; if (me.newm) b = (me.flags & 1) * NSHAPE + me.newm; }
        bin[k] = b;
      }
    });
    c.par([&](int tid) {
      for (int k = tid; k < RANK_CHUNK; k += c.nthreads()) {
        const int b = bin[k];
        RegroupDesc e; e.rank = -1; e.m = 0; e.a_src = e.a_dst = e.pq_src = e.pq_dst = 0;
        if (b >= 0) {
          int before = 0;
          for (int j = 0; j < k; j++) before += (bin[j] == b);
          const long long slot = slot0 + k;
          const SlotMeta me = sl.meta[slot];
          const int m = me.newm, kind = me.flags & 1, ms = slot_region(sl, slot), MT = sl.MT[ms];
          const int rank = (kind ? tv.n_old[m] : 0) + counts[(long long)b * nchunks + c.block()] + before;
          const long long ls = slot - sl.slot_begin[ms];
          e.rank = rank; e.m = m;
          e.a_src = sl.A_off[ms] + ls * (long long)MT * d; e.a_dst = tv.A_base[m] + (long long)rank * m * d;
          e.pq_src = sl.pq_off[ms] + ls * MT; e.pq_dst = tv.pq_base[m] + (long long)rank * m;
          const long long gt = tv.t_begin[m] + rank;
          double* bo = tv.b + gt * d; const double* bi = sl.b + slot * d;
          for (int i = 0; i < d; i++) bo[i] = bi[i];
          const unsigned char* ci = sl.cmap + slot * MAXM; unsigned char* co = tv.cmap + gt * MAXM;
          if (me.flags & 2) for (int i = 0; i < me.pbc; i++) co[i] = ci[i];         // only coaligned children own a map, of pbc entries (initcheck-clean)
          tv.meta[gt] = me; slot_of_term[gt] = slot;
        }
        desc[k] = e;
      }
    });
    const int EA = sp.max_shape * d, EP = sp.max_shape;
    c.par([&](int tid) {
      for (int e = tid; e < RANK_CHUNK * EA; e += c.nthreads()) {
        const int k = e / EA, i = e - k * EA;
        const RegroupDesc& ds = desc[k];
        if (ds.rank >= 0 && i < ds.m * d) tv.A[ds.a_dst + i] = sl.A[ds.a_src + i];
      }
      for (int e = tid; e < RANK_CHUNK * EP; e += c.nthreads()) {
        const int k = e / EP, i = e - k * EP;
        const RegroupDesc& ds = desc[k];
        if (ds.rank >= 0 && i < ds.m) { tv.p[ds.pq_dst + i] = sl.p[ds.pq_src + i]; tv.q[ds.pq_dst + i] = sl.q[ds.pq_src + i]; }
      }
    });
  }
};

// ---------------------------------------------------------------------------------------------
// K11: moment contributions of the surviving terms from their NEW tables (eval_g_yei_after_ftr, cauchy_term.hpp:403-433),
// used by compute_moments(false) when print_basic_info is set (cauchy_estimator.hpp:1166-1171, quirk A.9 iii).
// ---------------------------------------------------------------------------------------------
struct KPostFtrMoments {
  StepParams sp; GenView gen; cplx* g; double* y;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int r = c.block() * c.nthreads() + tid;
      if (r >= gen.n_alive) return;
      const int d = sp.d, gid = gen.alive[r], m = gen_m(gen, gid);
      const double* A = gen_A(gen, gid, m, d); const double* p = gen_p(gen, gid, m); const double* b = gen_b(gen, gid, d);
      double tmp[MAXD];
      for (int j = 0; j < d; j++) tmp[j] = 0;
      int enc_sv = 0;
      for (int l = 0; l < m; l++) {
        const double s = dot_lr(A + l * d, sp.root_point, d) > 0 ? 1.0 : -1.0;
        const double sc = p[l] * s;
        for (int j = 0; j < d; j++) tmp[j] += sc * A[l * d + j];
        if (s < 0) enc_sv |= 1 << l;
      }
      g[r] = g_lookup(enc_sv, m, gen_keys(gen, gid, m), gen_G(gen, gid, m), gen.cells[gid]);
      double* yo = y + (long long)r * 2 * d;
      for (int j = 0; j < d; j++) { yo[2 * j] = -tmp[j]; yo[2 * j + 1] = b[j]; }
    });
  }
};

// ---------------------------------------------------------------------------------------------
// K10: b <- b - delta over all parents (finalize_extended_moments est:1365-1383, shift_cf_by_bias est:1312).
// ---------------------------------------------------------------------------------------------
struct KShiftB {
  GenView gen; int d; double delta[MAXD]; double sign;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int r = c.block() * c.nthreads() + tid;
      if (r >= gen.n_alive) return;
      double* b = gen_b(gen, gen.alive[r], d);
      if (sign < 0) for (int j = 0; j < d; j++) b[j] -= delta[j];
      else for (int j = 0; j < d; j++) b[j] += delta[j];
    });
  }
};
// deterministic_time_prop, est:1331-1355: A <- A T^T, b <- T b (+ B u) without adding process noise.
struct KDetTimeProp {
  GenView gen; int d; double T[MAXD * MAXD]; double bu[MAXD]; int has_bu;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int r = c.block() * c.nthreads() + tid;
      if (r >= gen.n_alive) return;
      const int gid = gen.alive[r], m = gen_m(gen, gid);
      double* A = gen_A(gen, gid, m, d); double* b = gen_b(gen, gid, d);
      double work[MAXD];
      for (int i = 0; i < m; i++) {
        for (int k = 0; k < d; k++) work[k] = A[i * d + k];
        for (int j = 0; j < d; j++) { double sum = 0.0; for (int k = 0; k < d; k++) sum += work[k] * T[k + j * d]; A[i * d + j] = sum; }
      }
      for (int k = 0; k < d; k++) work[k] = b[k];
      for (int i = 0; i < d; i++) { double sum = 0.0; for (int j = 0; j < d; j++) sum += T[i * d + j] * work[j]; b[i] = sum; }
      if (has_bu) for (int i = 0; i < d; i++) b[i] += 1.0 * bu[i];
    });
  }
};

// ---------------------------------------------------------------------------------------------
// First step (cauchy_estimator.hpp:1179-1208): d+1 terms from (A0, p0, b0), closed-form tables
// (make_gtable_first, flattening.hpp:14-67).  One block; thread s handles slot s (0 = old term).
// out_mom: [2*(1+d+d*d)] raw sums (fz, mean, cov) in term order.
// ---------------------------------------------------------------------------------------------
struct KFirstStep {
  StepParams sp; const double *A0, *p0, *b0; GenView out; double* out_mom; int* out_count;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    const int d = sp.d, nq = 1 + d + d * d;
    double* sm_g = (double*)c.smem();                  // [(d+1)][2 + 2d] g and y per slot
    int* sm_valid = (int*)(sm_g + (d + 1) * (2 + 2 * d)); // [d+1] term index or -1
    double* sm_scale = (double*)(sm_valid + ((d + 2) & ~1));
    c.par([&](int tid) {
      if (tid > d) return;
      const int s = tid, m = d, t = (s == 0) ? m : s - 1;
      double mu[(MAXD + 1) * MAXD], rho[MAXD + 1]; unsigned F_int = 0;
      for (int l = 0; l < m; l++) {
        double* mu_l = mu + l * d;
        for (int i = 0; i < d; i++) mu_l[i] = A0[l * d + i];
        const double H_mu = dot_lr(sp.H, mu_l, d), a = fabs(H_mu);
        if (a < MU_EPS) rho[l] = p0[l];
        else { const double sc = 1.0 / H_mu; for (int i = 0; i < d; i++) mu_l[i] *= sc; rho[l] = p0[l] * a; F_int |= (1u << l); }
      }
      rho[m] = sp.gamma; for (int i = 0; i < d; i++) mu[m * d + i] = 0; F_int |= (1u << m);
      sm_valid[s] = ((F_int >> t) & 1u) ? 1 : -1;
      double* gy = sm_g + s * (2 + 2 * d);
      for (int i = 0; i < 2 + 2 * d; i++) gy[i] = 0;
      if (!((F_int >> t) & 1u)) return;
      const double zeta = sp.msmt - dot_lr(sp.H, b0, d);
      double cA[MAXD * MAXD], cp[MAXD], cb[MAXD];
      const double* mu_t = mu + t * d;
      for (int i = 0; i < d; i++) cb[i] = b0[i] + zeta * mu_t[i];
      unsigned hofs = 0; int l = 0;
      for (int _l = 0; _l < m + 1; _l++) {
        if (_l == t) continue;
        cp[l] = rho[_l];
        if ((F_int >> _l) & 1u) for (int i = 0; i < d; i++) cA[l * d + i] = mu[_l * d + i] - mu_t[i];
        else { for (int i = 0; i < d; i++) cA[l * d + i] = mu[_l * d + i]; hofs |= (1u << l); }
        l++;
      }
      cplx g = eval_g_yei(cA, cp, cb, m, d, hofs, zeta, rho[t], sp.root_point, true, 0, 0, 0, nullptr, nullptr, 0, gy + 2);
      gy[0] = g.re; gy[1] = g.im;
      // term index: old term first, then the integrable children in t order (est:1182)
      int idx = 0;
      if (s > 0) { idx = 1; for (int tt = 0; tt < t; tt++) idx += (F_int >> tt) & 1u; }
      sm_valid[s] = idx;
      double* Ao = gen_A(out, idx, d, d); double* po = gen_p(out, idx, d); double* bo = gen_b(out, idx, d);
      for (int i = 0; i < m * d; i++) Ao[i] = cA[i];
      for (int i = 0; i < m; i++) po[i] = cp[i];
      for (int i = 0; i < d; i++) bo[i] = cb[i];
      // c, d and Horthog are needed by the table pass; stash them behind y in shared memory is not enough (y is d complex),
      // so recompute-free: store in the G slot 0 of the table (overwritten below after being read back).
      cplx* Gt = gen_G(out, idx, d);
      Gt[0] = make_cplx(zeta, rho[t]);
      unsigned* Kt = gen_keys(out, idx, d);
      Kt[0] = hofs;
    });
    c.par([&](int tid) {          // compute_moments(true), est:524-579: serial sums in term order
      if (tid != 0) return;
      cplx acc[1 + MAXD + MAXD * MAXD];
      for (int i = 0; i < nq; i++) acc[i] = make_cplx(0, 0);
      int nt = 0;
      for (int idx = 0; idx <= d; idx++)
        for (int s = 0; s <= d; s++) {
          if (sm_valid[s] != idx) continue;
          nt++;
          const double* gy = sm_g + s * (2 + 2 * d);
          const cplx g = make_cplx(gy[0], gy[1]);
          acc[0] = cadd(acc[0], g);
          for (int j = 0; j < d; j++) {
            const cplx yj = make_cplx(gy[2 + 2 * j], gy[3 + 2 * j]);
            acc[1 + j] = cadd(acc[1 + j], cmul(g, yj));
            for (int k = 0; k < d; k++) acc[1 + d + j * d + k] = csub(acc[1 + d + j * d + k], cmul(cmul(g, yj), make_cplx(gy[2 + 2 * k], gy[3 + 2 * k])));
          }
        }
      for (int i = 0; i < nq; i++) { out_mom[2 * i] = acc[i].re; out_mom[2 * i + 1] = acc[i].im; }
      *out_count = nt;
      sm_scale[0] = (1.0 / (2.0 * M_PI)) / acc[0].re;       // G_SCALE_FACTOR, est:567
    });
    c.par([&](int tid) {          // make_gtable_first, flattening.hpp:14-67 (one thread per term)
      if (tid > d || sm_valid[tid] < 0) return;
      const int idx = sm_valid[tid], m = d, cells = 1 << (d - 1);
      cplx* Gt = gen_G(out, idx, d); unsigned* Kt = gen_keys(out, idx, d);
      const double c_val = Gt[0].re, d_val = Gt[0].im; const unsigned hofs = Kt[0];
      const double* p = gen_p(out, idx, d);
      for (int j = 0; j < cells; j++) {
        double ygi = 0;
        for (int k = 0; k < m; k++) if (!((hofs >> k) & 1u)) ygi += p[k] * ((((j >> k) & 1) == 0) ? 1.0 : -1.0);
        cplx v = csub(cdiv(make_cplx(1, 0), make_cplx(ygi + d_val, c_val)), cdiv(make_cplx(1, 0), make_cplx(ygi - d_val, c_val)));
        Gt[j] = cscale(v, sm_scale[0]); Kt[j] = (unsigned)j;
      }
      out.cells[idx] = cells; out.g_m[idx] = (unsigned char)d; out.alive[idx] = idx;
    });
  }
};

}  // namespace mce
#endif
