// mce_engine.h -- host-side step driver of the B200 MCE path: owns the device-resident term store and
// sequences the kernels of mce_kern_*.h for CauchyEstimator::step() (cauchy_estimator.hpp:1211-1245).
// The driver is written against a Backend policy (allocation, copies, kernel launch, sort/scan primitives);
// libmce_b200.so instantiates it with the CUDA backend only (backend_cuda.cuh).
#ifndef MCE_ENGINE_H_
#define MCE_ENGINE_H_

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "mce_kern_ftr.h"
#include "mce_kern_group.h"
#include "mce_kern_group2.h"
#include "mce_kern_prop.h"
#include "mce_kern_cpdf.h"
#include "mce_kern_part.h"

namespace mce {

// reference error bits, cauchy_constants.hpp:104-116
enum { ERROR_COVARIANCE_UNSTABLE_ANY_STEP = 0, ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_FINAL_MSMT = 1,
       ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT = 2, ERROR_COVARIANCE_AT_CURRENT_STEP_DNE = 3,
       ERROR_MEAN_UNSTABLE_ANY_STEP = 4, ERROR_MEAN_UNSTABLE_CURRENT_STEP_FINAL_MSMT = 5,
       ERROR_MEAN_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT = 6, ERROR_MEAN_AT_CURRENT_STEP_DNE = 7,
       ERROR_FZ_UNSTABLE = 8, ERROR_FZ_NEGATIVE = 9 };

struct StepStats {
  double ms_total = 0, ms_tp = 0, ms_mu = 0, ms_moments = 0, ms_regroup = 0, ms_ftr = 0, ms_gtable = 0, ms_compact = 0;
  long long parents = 0, slots = 0, terms_after_muc = 0, groups = 0, survivors = 0;
  long long bytes_gtable = 0, bytes_step = 0, launches = 0;
  int ftr_rounds_max = 0, diag_alias = 0, diag_hash = 0;
  double ev_step_ms = 0, ev_gtable_ms = 0, ev_mu_ms = 0, ev_moments_ms = 0, ev_ftr_ms = 0;   // CUDA-event durations on the stream
  long long cells_parents = 0, cells_survivors = 0, gtable_launches = 0, gtable_lean_launches = 0;
  int big_groups = 0;
};

template <class BE>
struct DevBuf {           // grow-only device buffer; contents are not preserved across growth
  BE* be = nullptr; void* p = nullptr; size_t cap = 0;
  void* ensure(size_t bytes) {
    if (bytes > cap) {
      if (p) be->free(p);
      size_t want = bytes + bytes / 4 + 256;
      p = be->alloc(want); cap = want;
    }
    return p;
  }
  template <class T> T* as() const { return (T*)p; }
  void release() { if (p) be->free(p); p = nullptr; cap = 0; }
};

// Host-side covariance checks (cauchy_util.hpp:1882-2004); eigenvalues of the lower triangle by cyclic Jacobi
// (the reference's NR tred2/tqli also reads the lower triangle, eig_solve.hpp:263-404).
inline void sym_eigvals(const double* A, double* ev, int n) {
  std::vector<double> a(n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) a[i * n + j] = a[j * n + i] = A[i * n + j];
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
      if (fabs(a[p * n + q]) < 1e-300) continue;
      double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * a[p * n + q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < n; k++) { double akp = a[k * n + p], akq = a[k * n + q]; a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = a[p * n + k], aqk = a[q * n + k]; a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk; }
    }
  }
  for (int i = 0; i < n; i++) ev[i] = a[i * n + i];
}
inline int covariance_checker(const cplx* cov, int d) {
  int flags = 0; std::vector<double> cr(d * d), ci(d * d), eig(d);
  for (int i = 0; i < d * d; i++) { cr[i] = cov[i].re; ci[i] = cov[i].im; }
  sym_eigvals(cr.data(), eig.data(), d);
  for (int i = 0; i < d; i++) if (eig[i] < -1e-5) flags |= 1 << 0;
  for (int i = 0; i < d - 1; i++) {
    double sig_ii = sqrt(cr[i * d + i]);
    for (int j = i + 1; j < d; j++) { double corr = cr[i * d + j] / (sig_ii * sqrt(cr[j * d + j])); if (fabs(corr) > 1) flags |= 1 << 1; }
  }
  for (int i = 0; i < d; i++) for (int j = i; j < d; j++) { double ratio = fabs(ci[i * d + j]) / (fabs(cr[i * d + j]) + 1e-15); if (ratio > 10) flags |= 1 << 2; }
  for (int i = 0; i < d * d; i++) if (fabs(ci[i]) > 2000) flags |= 1 << 3;
  return flags;
}

template <class BE>
class Engine {
 public:
  BE be;
  int d, cmcc, pncc, p, steps, num_estimation_steps, max_shape, shape_range;
  int master_step = 0, Nt = 1, Nt_muc = 1, numeric_moment_errors = 0, skip_post_mu = 0, print_basic_info = 0;
  int tr_order[12];
  std::vector<double> A0, p0, b0, root_point, b_pert;
  std::vector<double> A1, p1, b1;   // the first term as the next first step will read it (the reference keeps it in childterms_workspace: term:772-785)
  double G_SCALE_FACTOR = 0;
  cplx fz, last_fz, fz_mu; std::vector<cplx> mean, var, last_mean, last_var, mean_mu, var_mu;   // *_mu: as of the end of the measurement update
  std::vector<int> terms_per_shape, muc_per_shape;
  StepStats stats;
  std::string error;
  bool finished = false;     // the window's last step has run: only reset() is valid now
  bool fast_moments = false; // every moment sum but Re fz by a two-level tree instead of the reference's serial order; Re fz by the exact scan (bit-identical)

  // ---- device state ----
  struct GenStore {
    GenView v; DevBuf<BE> g_m, cells, alive, A, p, b, keys, G, rbm, rpf, gmax;
    DevBuf<BE> gpos;                    // partitioned estimator: global alive rank of every local survivor
    std::vector<int> alive_per_shape;   // survivors per shape
    long long sum_cells = 0;            // total table cells of the survivors
  } gen[2], imp;                        // imp: the parents this rank's terms descend from, fetched from their home ranks (partitioned estimator)
  int cur = 0;
  DevBuf<BE> wsA, wsp, wsb, wsm, wsSgn, wsXor, wsTpB, wsTpBc;
  DevBuf<BE> cpVal0, cpCache, cpXs, cpOut, cpYs, cpRecs, cpV, cpBad;   // point-wise marginal cpdf (mce_kern_cpdf.h)
  double cpdf_ms = 0;                          // device time of the last marginal_1d_points call (CUDA events)
  DevBuf<BE> slA, slp, slq, slb, slmeta, slcmap, slg, sly;
  DevBuf<BE> tvA, tvp, tvq, tvb, tvmeta, tvcmap, slotOfTerm;
  DevBuf<BE> sumTiles;
  DevBuf<BE> rankCounts, rankTotals, momPartial, momOut, scratchI0, scratchI1, scratchI2, scratchI3, scratchK0, scratchK1;
  DevBuf<BE> ftrF, ftrWide, grpOrder, grpStart, aliveFlag, diagBuf, unkBuf, initBuf;
  DevBuf<BE> bigGroups, bigParts, bigCnt, bigRows, bigFlags, bigKeys;
  // ---- partitioned estimator (mce_kern_part.h): every term lives on one rank ----
  DevBuf<BE> ptKeys, ptAllKeys, ptAllSorted, ptIdx0, ptIdx1, ptSplit, ptDest, ptCnt, ptRunOff, ptSend, ptRecv, ptGk, ptGkS, ptOrd, ptGidx, ptNold;
  DevBuf<BE> ptIKey, ptIKeyS, ptIIdx, ptIIdxS, ptFlag, ptPos, ptIList, ptRList, ptSeg, ptNImp, ptReq, ptPRecS, ptPRecR, ptIgpos, ptBxG, ptSKey, ptSAll, ptHost, ptGg;
  DevBuf<BE> iwsSgn, iwsXor, iwsTpB, iwsTpBc, txA, txp, txq, txb, txmeta, txcmap;
  std::vector<int> g_alive_per_shape;           // survivors per shape over all ranks
  long long fast_moments_slots = 300000;        // mce_options.fast_moments_min_slots: below, the 114 dependent chains take < 1.5 ms and hide behind the term reduction: fast_moments keeps them (bit-exact)
  long long early_scale_slots = 400000;         // steps with at least this many slots take G_SCALE_FACTOR from the exact scan of Re fz (mce_options.early_scale_min_slots)
  int moments_mode = 0;                         // 0: every rank adds ALL slots in the reference's order (bit-exact); 1: per-rank serial sums added in rank order;
                                                // 2: like 1, but Re fz -- the one sum that feeds back into the filter -- is the exact scan over ALL slots (KSumScan)
  int scan_restarts = 0;
  struct PartStats { long long bytes_terms = 0, bytes_parents = 0, bytes_moments = 0, bytes_keys = 0; int imports = 0, owned = 0; double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0}; } pstats;   // ms: CUDA-event stage times of the exchange
  bool phase_timing = false;                    // mce_options.phase_timing
  bool lean_groups = false;                     // mce_options.lean_group_kernel
  int big_T = BIG_T;                            // groups with more members are split (mce_options.group_split_threshold)
  long long big_scratch_cap = 6LL << 30;       // bytes of addend rows above which group splitting is skipped for a step
  // debug capture
  bool capture = false;
  struct CapShape { int m = 0, n = 0; std::vector<double> A, p, q, b, cd; std::vector<int> meta, F; std::vector<unsigned char> cmap; std::vector<signed char> csmap; };
  std::vector<CapShape> cap;

  Engine(int d_, int cmcc_, int pncc_, int p_, int steps_, const double* A0_, const double* p0_, const double* b0_,
         const double* root_point_, const double* b_pert_, const int* tr_order_, int print_info)
      : d(d_), cmcc(cmcc_), pncc(pncc_), p(p_), steps(steps_) {
    num_estimation_steps = p * steps;
    max_shape = d > 1 ? (steps - 1) * pncc + d : d + pncc;     // est:97
    shape_range = max_shape + 1;
    print_basic_info = print_info;
    for (int i = 0; i < 12; i++) tr_order[i] = tr_order_ ? tr_order_[i] : i;
    A0.assign(A0_, A0_ + d * d); p0.assign(p0_, p0_ + d); b0.assign(b0_, b0_ + d);
    A1 = A0; p1 = p0; b1 = b0;
    root_point.assign(root_point_, root_point_ + d);
    b_pert.assign(MAXM, 0.0);
    for (int i = 0; i < max_shape && i < MAXM; i++) b_pert[i] = b_pert_[i];
    mean.assign(d, make_cplx(0, 0)); var.assign(d * d, make_cplx(0, 0)); last_mean = mean; last_var = var; mean_mu = mean; var_mu = var;
    fz = last_fz = make_cplx(0, 0);
    terms_per_shape.assign(shape_range, 0); muc_per_shape.assign(shape_range, 0);
    terms_per_shape[d] = 1;
    DevBuf<BE>* all[] = {&gen[0].g_m, &gen[0].cells, &gen[0].alive, &gen[0].A, &gen[0].p, &gen[0].b, &gen[0].keys, &gen[0].G, &gen[0].rbm, &gen[0].rpf, &gen[1].rbm, &gen[1].rpf, &gen[0].gmax, &gen[1].gmax,
                         &gen[1].g_m, &gen[1].cells, &gen[1].alive, &gen[1].A, &gen[1].p, &gen[1].b, &gen[1].keys, &gen[1].G,
                         &wsA, &wsp, &wsb, &wsm, &wsSgn, &wsXor, &wsTpB, &wsTpBc, &slA, &slp, &slq, &slb, &slmeta, &slcmap, &slg, &sly,
                         &tvA, &tvp, &tvq, &tvb, &tvmeta, &tvcmap, &slotOfTerm, &rankCounts, &rankTotals, &momPartial, &momOut, &sumTiles,
                         &gen[0].gpos, &gen[1].gpos, &imp.g_m, &imp.cells, &imp.alive, &imp.A, &imp.p, &imp.b, &imp.keys, &imp.G, &imp.rbm, &imp.rpf, &imp.gmax, &imp.gpos,
                         &ptKeys, &ptAllKeys, &ptAllSorted, &ptIdx0, &ptIdx1, &ptSplit, &ptDest, &ptCnt, &ptRunOff, &ptSend, &ptRecv, &ptGk, &ptGkS, &ptOrd, &ptGidx, &ptNold,
                         &ptIKey, &ptIKeyS, &ptIIdx, &ptIIdxS, &ptFlag, &ptPos, &ptIList, &ptRList, &ptSeg, &ptNImp, &ptReq, &ptPRecS, &ptPRecR, &ptIgpos, &ptBxG, &ptSKey, &ptSAll, &ptHost, &ptGg,
                         &iwsSgn, &iwsXor, &iwsTpB, &iwsTpBc, &txA, &txp, &txq, &txb, &txmeta, &txcmap,
                         &scratchI0, &scratchI1, &scratchI2, &scratchI3, &scratchK0, &scratchK1, &ftrF, &ftrWide, &grpOrder, &grpStart,
                         &aliveFlag, &diagBuf, &unkBuf, &initBuf, &bigGroups, &bigParts, &bigCnt, &bigRows, &bigFlags, &bigKeys, &cpVal0, &cpCache, &cpXs, &cpOut, &cpYs, &cpRecs, &cpV, &cpBad};
    for (auto* b : all) { b->be = &be; all_bufs.push_back(b); }
    gen[0].alive_per_shape.assign(NSHAPE, 0); gen[1].alive_per_shape.assign(NSHAPE, 0); imp.alive_per_shape.assign(NSHAPE, 0);
    g_alive_per_shape.assign(NSHAPE, 0);
  }
  ~Engine() {}
  std::vector<DevBuf<BE>*> all_bufs;
  void release_all() { for (auto* b : all_bufs) b->release(); }   // device memory of this estimator (mce_destroy)

  static bool supported(int d, int max_shape, int pncc, std::string* why) {
    if (d < 2 || d > MAXD) { *why = "state dimension must be in [2, 8]"; return false; }
    if (max_shape > MAXM - 1) { *why = "max hyperplane count exceeds 31 (reference cap, est:231-235)"; return false; }
    if (pncc > MAXPN) { *why = "pncc > 4 not supported"; return false; }
    // the group kernel keeps one table in shared memory: keys + two complex values per cell, or -- lean variant, used where that does not fit -- keys + one
    const int Hcap = cell_count_central_half(max_shape, d);
    const size_t need = max_shape <= 16 ? KGTable2T<G2_NORMAL, true>::smem_bytes(Hcap, (1 << max_shape) / 32 > 1 ? (1 << max_shape) / 32 : 1)
                                        : KGTable::smem_bytes(next_pow2(Hcap < 4 ? 4 : Hcap));
    if (need > 227 * 1024) {
      *why = "a window this deep gives tables of up to " + std::to_string(Hcap) + " cells; the group kernel holds a table in shared memory (227 KB, about 11000 cells for up to 16 hyperplanes, 5500 beyond)";
      return false;
    }
    // every other kernel whose shared memory grows with the deepest shape of the window, at that shape
    const int cgen = cell_count_general(max_shape, d);
    const size_t need_mu = KMsmtUpdate2::smem_bytes(max_shape, d), need_ftr = KFtrRoundTiled::smem_bytes(max_shape, d);
    const size_t need_tp = max_shape <= 16 ? KTpDce2T<2>::smem_bytes((1 << max_shape) / 32 > 1 ? (1 << max_shape) / 32 : 1, 128, d)
                                           : KTpDce::smem_bytes(next_pow2(DCE_STORAGE_MULT * cgen), next_pow2(cgen + 1), 128);
    const size_t worst = need_mu > need_ftr ? (need_mu > need_tp ? need_mu : need_tp) : (need_ftr > need_tp ? need_ftr : need_tp);
    if (worst > 227 * 1024) {
      *why = "a window this deep (" + std::to_string(max_shape) + " hyperplanes in " + std::to_string(d) + " dimensions) needs " + std::to_string(worst / 1024) +
             " KB of shared memory in the " + (worst == need_mu ? "measurement-update" : worst == need_ftr ? "term-reduction" : "time-propagation cell-enumeration") + " kernel; the limit is 227 KB";
      return false;
    }
    return true;
  }

  // ------------------------------------------------------------------------------------------
  void fill_gen_layout(GenStore& g, const std::vector<int>& groups_per_shape) {
    GenView& v = g.v;
    long long gA = 0, gp = 0, gt = 0, gr = 0; int gid = 0;
    for (int m = 0; m < NSHAPE; m++) {
      v.gid_begin[m] = gid; v.A_base[m] = gA; v.p_base[m] = gp; v.tab_base[m] = gt; v.rk_base[m] = gr;
      const int n = m < (int)groups_per_shape.size() ? groups_per_shape[m] : 0;
      const int stride = m >= 1 ? cell_count_central_half(m, d) : 0;
      v.tab_stride[m] = stride;
      gid += n; gA += (long long)n * m * d; gp += (long long)n * m; gt += (long long)n * stride;
      if (max_shape <= 16 && m <= 16) gr += (long long)n * rank_words(m);
    }
    v.gid_begin[NSHAPE] = gid; v.n_groups = gid;
    v.g_m = (unsigned char*)g.g_m.ensure((size_t)gid + 16);
    v.cells = (int*)g.cells.ensure(sizeof(int) * ((size_t)gid + 4));
    v.alive = (int*)g.alive.ensure(sizeof(int) * ((size_t)gid + 4));
    v.A = (double*)g.A.ensure(sizeof(double) * (size_t)(gA + 8));
    v.p = (double*)g.p.ensure(sizeof(double) * (size_t)(gp + 8));
    v.b = (double*)g.b.ensure(sizeof(double) * ((size_t)gid * d + 8));
    v.keys = (unsigned*)g.keys.ensure(sizeof(unsigned) * (size_t)(gt + 8));
    v.G = (cplx*)g.G.ensure(sizeof(cplx) * (size_t)(gt + 8));
    v.rbm = (unsigned*)g.rbm.ensure(sizeof(unsigned) * (size_t)(gr + 8));
    v.rpf = (unsigned short*)g.rpf.ensure(sizeof(unsigned short) * (size_t)(gr + 8));
    v.gmax = (double*)g.gmax.ensure(sizeof(double) * ((size_t)gid + 8));
  }

  StepParams make_params(double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H, double gamma,
                         const double* B, const double* u, bool with_tp) {
    StepParams sp; memset(&sp, 0, sizeof(sp));
    sp.d = d; sp.with_tp = with_tp; sp.skip_post_mu = skip_post_mu; sp.max_shape = max_shape;
    for (int i = 0; i < 12; i++) sp.tr_order[i] = tr_order[i];
    for (int i = 0; i < d; i++) { sp.H[i] = H[i]; sp.root_point[i] = root_point[i]; }
    for (int i = 0; i < MAXM; i++) sp.b_pert[i] = b_pert[i];
    sp.msmt = msmt; sp.gamma = gamma; sp.gscale = G_SCALE_FACTOR;
    if (with_tp) {
      for (int i = 0; i < d * d; i++) sp.Phi[i] = Phi[i];
      // precoalign_Gamma_beta, cauchy_util.hpp:128-210 (host: pncc x d values)
      double tG[MAXPN * MAXD], tb[MAXPN]; int cm = pncc;
      for (int i = 0; i < cm; i++) { for (int j = 0; j < d; j++) tG[i * d + j] = Gamma[j * cm + i]; tb[i] = beta[i]; }
      for (int i = 0; i < cm; i++) {
        double nf = 0; for (int j = 0; j < d; j++) nf += fabs(tG[i * d + j]);
        for (int j = 0; j < d; j++) tG[i * d + j] /= nf;
        tb[i] *= nf;
      }
      bool F[MAXPN]; for (int i = 0; i < cm; i++) F[i] = true;
      for (int i = 0; i < cm - 1; i++) if (F[i])
        for (int j = i + 1; j < cm; j++) if (F[j]) {
          bool pos, neg; coalign_gates(tG + i * d, tG + j * d, d, &pos, &neg);
          if (pos || neg) { F[j] = false; tb[i] += tb[j]; }
        }
      int n = 0;
      for (int i = 0; i < cm; i++) if (F[i]) { for (int j = 0; j < d; j++) sp.GammaT[n * d + j] = tG[i * d + j]; sp.beta[n] = tb[i]; n++; }
      sp.npn = n;
      if (cmcc > 0 && B && u) {
        sp.has_bu = 1;
        for (int i = 0; i < d; i++) { double sum = 0.0; for (int j = 0; j < cmcc; j++) sum += B[i * cmcc + j] * u[j]; sp.bu[i] = sum; }
      }
    }
    return sp;
  }

  // finalize_cached_moments (est:340-357) on the raw sums; `check` adds moments_numerical_check (est:359-461)
  void finalize_moments(const double* raw /*2*(1+d+d*d)*/, bool check) {
    fz = make_cplx(raw[0], raw[1]);
    for (int i = 0; i < d; i++) mean[i] = make_cplx(raw[2 + 2 * i], raw[3 + 2 * i]);
    for (int i = 0; i < d * d; i++) var[i] = make_cplx(raw[2 + 2 * d + 2 * i], raw[3 + 2 * d + 2 * i]);
    G_SCALE_FACTOR = (1.0 / (2.0 * M_PI)) / fz.re;
    const cplx Ifz = make_cplx(0, fz.re);
    for (int i = 0; i < d; i++) mean[i] = cdiv(mean[i], Ifz);
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) var[i * d + j] = csub(cdiv(var[i * d + j], fz), cmul(mean[i], mean[j]));
    fz_mu = fz; mean_mu = mean; var_mu = var;
    if (!check) return;
    const bool first_msmt = (master_step % p) == 0, not_last = (master_step % p) != (p - 1), last = (master_step % p) == (p - 1);
    if (first_msmt) {
      numeric_moment_errors |= (1 << ERROR_MEAN_AT_CURRENT_STEP_DNE) | (1 << ERROR_COVARIANCE_AT_CURRENT_STEP_DNE);
      numeric_moment_errors &= ~((1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_FINAL_MSMT) | (1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT) |
                                 (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT) | (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_FINAL_MSMT));
    }
    int nw = 0;
    if (fz.re <= 0) nw |= (1 << ERROR_FZ_NEGATIVE);
    if (fabs(fz.im / (1e-15 + fz.re)) > 1e-3) nw |= (1 << ERROR_FZ_UNSTABLE);
    bool mean_okay = true;
    for (int i = 0; i < d; i++) {
      double mr = fabs(mean[i].re), mi = fabs(mean[i].im), ratio = mi / (1e-15 + mr);
      if ((ratio > 1e-1) || (mi > 0.001)) {
        nw |= (1 << ERROR_MEAN_UNSTABLE_ANY_STEP);
        if (not_last) nw |= (1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT);
        if (last) nw |= (1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_FINAL_MSMT);
        mean_okay = false;
      }
    }
    bool cov_okay = true;
    if (covariance_checker(var.data(), d)) {
      nw |= (1 << ERROR_COVARIANCE_UNSTABLE_ANY_STEP);
      if (not_last) nw |= (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT);
      if (last) nw |= (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_FINAL_MSMT);
      cov_okay = false;
    }
    numeric_moment_errors |= nw;
    if (mean_okay) numeric_moment_errors &= ~(1 << ERROR_MEAN_AT_CURRENT_STEP_DNE);
    if (cov_okay) numeric_moment_errors &= ~(1 << ERROR_COVARIANCE_AT_CURRENT_STEP_DNE);
    if (mean_okay) last_mean = mean; else mean = last_mean;
    if (cov_okay) last_var = var; else var = last_var;
    if (!mean_okay && !cov_okay) fz = last_fz; else last_fz = fz;
    fz_mu = fz; mean_mu = mean; var_mu = var;
  }

  // ------------------------------------------------------------------------------------------
  int step(double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H, double gamma,
           const double* B, const double* u) {
    if (numeric_moment_errors & (1 << ERROR_FZ_NEGATIVE)) return numeric_moment_errors;     // est:1214-1219
    if (master_step < 0 || master_step > num_estimation_steps) { error = "master_step was set outside [0, num_estimation_steps]"; return -4; }
    if (master_step == num_estimation_steps || finished) { error = "master_step == num_estimation_steps: reset() the estimator first (est:1220-1225)"; return -4; }
    skip_post_mu = (master_step == num_estimation_steps - 1);                               // SKIP_LAST_STEP, est:1229
    stats = StepStats();
    double t0 = be.tic();
    be.ev_record(0);
    int rc = (master_step == 0) ? step_first(msmt, H, gamma) : step_general(msmt, Phi, Gamma, beta, H, gamma, B, u);
    if (rc < 0) return rc;
    be.ev_record(1);
    stats.ev_step_ms = be.ev_elapsed(0, 1);
    stats.ms_total = be.toc(t0);
    stats.launches = be.launch_count; be.launch_count = 0;
    master_step++;
    return numeric_moment_errors;
  }

  int step_first(double msmt, const double* H, double gamma) {
    StepParams sp = make_params(msmt, nullptr, nullptr, nullptr, H, gamma, nullptr, nullptr, false);
    const int W = be.shard.world, R = be.shard.rank;        // partitioned estimator: the first term lives on rank 0
    GenStore& ng = gen[cur];
    std::vector<int> groups(NSHAPE, 0); if (R == 0) groups[d] = d + 1;
    fill_gen_layout(ng, groups);
    const int nq = 1 + d + d * d;
    std::vector<double> raw(2 * nq); int nt = 0;
    if (R == 0) {
      double* init = (double*)initBuf.ensure(sizeof(double) * (d * d + 2 * d));
      be.h2d(init, A1.data(), sizeof(double) * d * d);
      be.h2d(init + d * d, p1.data(), sizeof(double) * d);
      be.h2d(init + d * d + d, b1.data(), sizeof(double) * d);
      double* mom = (double*)momOut.ensure(sizeof(double) * 4 * nq + 16);
      int* cnt = (int*)unkBuf.ensure(64);
      KFirstStep k{sp, init, init + d * d, init + d * d + d, ng.v, mom, cnt};
      size_t smem = sizeof(double) * ((d + 1) * (2 + 2 * d) + 4) + sizeof(int) * (d + 4);
      be.launch(k, 1, 32, smem);
      be.d2h(raw.data(), mom, sizeof(double) * 2 * nq); be.d2h(&nt, cnt, sizeof(int));
    }
    if (W > 1) {
      std::vector<long long> mine(2 * nq + 1, 0);
      memcpy(mine.data(), raw.data(), sizeof(double) * 2 * nq); mine[2 * nq] = nt;
      const std::vector<long long> all = allgather_ll(mine);
      memcpy(raw.data(), all.data(), sizeof(double) * 2 * nq); nt = (int)all[2 * nq];
    }
    finalize_moments(raw.data(), false);        // compute_moments(true): no numerical check on the first step (quirk A.9 iv)
    const int nt_loc = R == 0 ? nt : 0;
    ng.v.n_alive = nt_loc;
    std::fill(ng.alive_per_shape.begin(), ng.alive_per_shape.end(), 0); ng.alive_per_shape[d] = nt_loc;
    std::fill(g_alive_per_shape.begin(), g_alive_per_shape.end(), 0); g_alive_per_shape[d] = nt;
    if (W > 1) { int* gp = (int*)ng.gpos.ensure(sizeof(int) * (size_t)(nt_loc + 4)); be.launch(KIota{nt_loc, gp}, 1, 64, 0); }
    Nt = nt; Nt_muc = nt; ng.sum_cells = (long long)nt_loc << (d - 1);
    std::fill(terms_per_shape.begin(), terms_per_shape.end(), 0); terms_per_shape[d] = nt; muc_per_shape = terms_per_shape;
    if (print_basic_info && W == 1) post_ftr_moments(sp); else fz = make_cplx(1, 0);       // est:1198-1204
    last_mean = mean; last_var = var; last_fz = fz;
    stats.parents = 1; stats.slots = d + 1; stats.terms_after_muc = nt; stats.groups = nt; stats.survivors = nt;
    return 0;
  }

  // ---- partitioned estimator: host-side collectives ----
  // all-gather of n 64-bit values per rank through a device buffer (the transports move device memory)
  std::vector<long long> allgather_ll(const std::vector<long long>& mine) {
    const int W = be.shard.world, R = be.shard.rank; const size_t n = mine.size();
    long long* buf = (long long*)ptHost.ensure(sizeof(long long) * n * W);
    be.h2d(buf + (size_t)R * n, mine.data(), sizeof(long long) * n);
    be.xchg_begin(); be.xchg_allgather(buf, sizeof(long long) * n); be.xchg_end();
    std::vector<long long> all(n * W);
    be.d2h(all.data(), buf, sizeof(long long) * n * W);
    return all;
  }
  // the same block of `bytes` from every rank lands at recv + roff[h] (all-gather with unequal contributions)
  void allgatherv(const void* send, long long bytes, void* recv, const std::vector<long long>& roff, const std::vector<long long>& rcnt) {
    const int W = be.shard.world;
    std::vector<long long> soff(W, 0), scnt(W, bytes);
    be.xchg_alltoallv(send, soff.data(), scnt.data(), recv, roff.data(), rcnt.data());
  }

  // Hybrid moments: only Re g of every slot, packed, at its canonical position of the global slot list (8 bytes per slot to combine instead of 16 + 16 d).
  long long part_gather_re(const SlotView& sl, const GenStore& pg, const double** re_out) {
    KSlotScatter k; memset(&k, 0, sizeof(k));
    long long ng_slots = 0; int base = 0;
    for (int m = 0; m < NSHAPE; m++) { k.gslot_begin[m] = ng_slots; k.gshape_base[m] = base; ng_slots += (long long)g_alive_per_shape[m] * (sl.MT[m] + 1); base += g_alive_per_shape[m]; }
    double* buf = (double*)ptGg.ensure(sizeof(double) * ((size_t)ng_slots + 8));
    be.memset(buf, 0, sizeof(double) * (size_t)ng_slots);
    k.sl = sl; k.d = d; k.gpos = pg.gpos.template as<int>(); k.re_out = buf;
    if (sl.n_slots > 0) be.launch(k, (int)((sl.n_slots + 127) / 128), 128, 0);
    be.xchg_begin(); be.xchg_allreduce_u32(buf, (size_t)ng_slots * 2); be.xchg_end();
    pstats.bytes_moments = (long long)(sizeof(double) * (size_t)ng_slots);
    *re_out = buf;
    return ng_slots;
  }

  // One serial-order sum (Re of n complex values), bit-identical to the dependent chain: tile summaries in parallel, then the exact walk (csrc/mce_kern_prop.h).
  void launch_sum_scan(bool side, const double* g, int stride, long long n, double* out) {
    const long long TILE = (long long)SS_NT * SS_E, ntiles = (n + TILE - 1) / TILE;
    KSumScan k{g, stride, n, out};
    static const bool no_tiles = getenv("MCE_SCAN_NO_TILES") != nullptr;      // measurement switch (tools/scan_real.py)
    if (ntiles >= 4 && !no_tiles) {
      const size_t off = (sizeof(double) * (size_t)ntiles + 63) & ~(size_t)63;
      unsigned char* buf = (unsigned char*)sumTiles.ensure(off + sizeof(SumTile) * (size_t)ntiles + 64);
      double* tsum = (double*)buf; SumTile* tiles = (SumTile*)(buf + off);
      KSumTileSums k1{g, stride, n, tsum}; KSumTileMaps k2{g, stride, n, tsum, tiles};
      if (side) { be.launch_side(k1, (int)ntiles, SS_NT, KSumTileSums::smem_bytes(SS_NT)); be.launch_side(k2, (int)ntiles, SS_NT, KSumTileMaps::smem_bytes(SS_NT)); }
      else { be.launch(k1, (int)ntiles, SS_NT, KSumTileSums::smem_bytes(SS_NT)); be.launch(k2, (int)ntiles, SS_NT, KSumTileMaps::smem_bytes(SS_NT)); }
      k.tiles = tiles;
    }
    if (side) be.launch_side(k, 1, SS_NT, KSumScan::smem_bytes(SS_NT)); else be.launch(k, 1, SS_NT, KSumScan::smem_bytes(SS_NT));
  }

  // Ordered moments (moments_mode 0): every slot's (g, y) is placed at its canonical position of the global slot list and the
  // list is combined over the ranks (each word has one non-zero contributor, so the integer sum is the value); every rank then
  // adds ALL slots in the reference's order.  Returns the global slot count; *g_out / *y_out point at the combined list.
  long long part_gather_slots(const SlotView& sl, const GenStore& pg, cplx** g_out, double** y_out, bool with_y = true) {
    KSlotScatter k; memset(&k, 0, sizeof(k));
    long long ng_slots = 0; int base = 0;
    for (int m = 0; m < NSHAPE; m++) { k.gslot_begin[m] = ng_slots; k.gshape_base[m] = base; ng_slots += (long long)g_alive_per_shape[m] * (sl.MT[m] + 1); base += g_alive_per_shape[m]; }
    const size_t words = (size_t)ng_slots * (2 + (with_y ? 2 * d : 0));
    double* buf = (double*)ptGg.ensure(sizeof(double) * (words + 8));
    be.memset(buf, 0, sizeof(double) * words);
    k.sl = sl; k.d = d; k.gpos = pg.gpos.template as<int>(); k.g_out = (cplx*)buf; k.y_out = with_y ? buf + 2 * ng_slots : nullptr;
    if (sl.n_slots > 0) be.launch(k, (int)((sl.n_slots + 127) / 128), 128, 0);
    be.xchg_begin(); be.xchg_allreduce_u32(buf, words * 2); be.xchg_end();
    pstats.bytes_moments = (long long)(sizeof(double) * words);
    *g_out = (cplx*)buf; if (y_out) *y_out = buf + 2 * ng_slots;
    return ng_slots;
  }

  // Routes the post-coalignment terms to the owners of their reduction keys and fetches the parents the owned terms descend
  // from (mce_kern_part.h).  In: the local TermView `tl` (+ slot_of_term) and every rank's (old, child) counts per shape.
  // Out: *tvx = the owned terms in canonical (gidx) order, `imp` + *iws = the imported parents the G-table kernel reads.
  void part_exchange(const StepParams& sp, const SlotView& sl, const TermView& tl, const long long* sot, const std::vector<long long>& all_tot,
                     bool with_tp, GenStore& pg, const ParentWs& ws, TermView* tvx, ParentWs* iws) {
    const int W = be.shard.world, R = be.shard.rank;
    std::vector<std::vector<long long>> nloc(W, std::vector<long long>(NSHAPE, 0));
    std::vector<long long> gN(NSHAPE, 0), gtb(NSHAPE + 1, 0);
    for (int h = 0; h < W; h++) for (int m = 0; m < NSHAPE; m++) { nloc[h][m] = all_tot[(size_t)h * 2 * NSHAPE + m] + all_tot[(size_t)h * 2 * NSHAPE + NSHAPE + m]; gN[m] += nloc[h][m]; }
    for (int m = 0; m < NSHAPE; m++) gtb[m + 1] = gtb[m] + gN[m];
    const long long nl = tl.t_begin[NSHAPE], gtot = gtb[NSHAPE];
    pstats = PartStats();
    be.ev_record(12);
    // 1. reduction keys of the local terms, all ranks' keys per shape, splitters
    unsigned long long* kloc = (unsigned long long*)ptKeys.ensure(sizeof(unsigned long long) * (size_t)(nl + 4));
    unsigned long long* kall = (unsigned long long*)ptAllKeys.ensure(sizeof(unsigned long long) * (size_t)(gtot + 4));
    unsigned long long* ksrt = (unsigned long long*)ptAllSorted.ensure(sizeof(unsigned long long) * (size_t)(gtot + 4));
    int* dmy0 = (int*)ptIdx0.ensure(sizeof(int) * (size_t)(gtot + 4)); int* dmy1 = (int*)ptIdx1.ensure(sizeof(int) * (size_t)(gtot + 4));
    unsigned long long* split_d = (unsigned long long*)ptSplit.ensure(sizeof(unsigned long long) * NSHAPE * PART_MAXW);
    for (int m = 1; m < NSHAPE; m++) if (tl.n[m] > 0) be.launch(KPartKeys{tl, m, d, tr_order[0], kloc + tl.t_begin[m]}, (tl.n[m] + 127) / 128, 128, 0);
    be.xchg_begin();
    for (int m = 1; m < NSHAPE; m++) if (gN[m] > 0) {
      std::vector<long long> roff(W), rcnt(W); long long o = gtb[m];
      for (int h = 0; h < W; h++) { roff[h] = o * 8; rcnt[h] = nloc[h][m] * 8; o += nloc[h][m]; }
      allgatherv(kloc + tl.t_begin[m], 8LL * tl.n[m], kall, roff, rcnt);
    }
    be.xchg_end();
    pstats.bytes_keys = 8 * gtot;
    be.memset(dmy0, 0, sizeof(int) * (size_t)gtot);
    for (int m = 1; m < NSHAPE; m++) if (gN[m] > 0) {
      be.sort_pairs(kall + gtb[m], ksrt + gtb[m], dmy0 + gtb[m], dmy1 + gtb[m], (int)gN[m]);
      be.launch(KPartSplit{ksrt + gtb[m], (int)gN[m], W, split_d + m * PART_MAXW}, W - 1, 128, sizeof(int) * 130);
    }
    be.ev_record(13);
    // 2. destination of every local term, send counts
    unsigned char* dest = (unsigned char*)ptDest.ensure((size_t)nl + 16);
    int* cnt_d = (int*)ptCnt.ensure(sizeof(int) * 2 * NSHAPE * PART_MAXW); int* cur_d = cnt_d + NSHAPE * PART_MAXW;
    be.memset(cnt_d, 0, sizeof(int) * 2 * NSHAPE * PART_MAXW);
    for (int m = 1; m < NSHAPE; m++) if (tl.n[m] > 0)
      be.launch(KPartDest{kloc + tl.t_begin[m], tl.n[m], W, split_d + m * PART_MAXW, dest + tl.t_begin[m], cnt_d + m * PART_MAXW}, (tl.n[m] + 127) / 128, 128, sizeof(int) * PART_MAXW);
    std::vector<int> hc(NSHAPE * PART_MAXW, 0);
    be.d2h(hc.data(), cnt_d, sizeof(int) * NSHAPE * PART_MAXW);
    std::vector<long long> mine((size_t)NSHAPE * W, 0);
    for (int m = 0; m < NSHAPE; m++) for (int h = 0; h < W; h++) mine[(size_t)m * W + h] = hc[m * PART_MAXW + h];
    const std::vector<long long> cm = allgather_ll(mine);                 // cm[(src * NSHAPE + m) * W + dst]
    auto C = [&](int src, int m, int dst) { return cm[((size_t)src * NSHAPE + m) * W + dst]; };
    be.ev_record(14);
    // 3. pack, exchange
    std::vector<long long> sbase(NSHAPE + 1, 0), rbase(NSHAPE + 1, 0), nrecv(NSHAPE, 0), run_off((size_t)NSHAPE * PART_MAXW, 0);
    for (int m = 0; m < NSHAPE; m++) {
      long long o = 0;
      for (int h = 0; h < W; h++) { run_off[(size_t)m * PART_MAXW + h] = o; o += C(R, m, h); nrecv[m] += C(h, m, R); }
      const long long RW = m >= 1 ? part_rec_words(m, d) : 0;
      sbase[m + 1] = sbase[m] + (long long)tl.n[m] * RW; rbase[m + 1] = rbase[m] + nrecv[m] * RW;
    }
    unsigned long long* sendb = (unsigned long long*)ptSend.ensure(sizeof(unsigned long long) * (size_t)(sbase[NSHAPE] + 4));
    unsigned long long* recvb = (unsigned long long*)ptRecv.ensure(sizeof(unsigned long long) * (size_t)(rbase[NSHAPE] + 4));
    long long* run_d = (long long*)ptRunOff.ensure(sizeof(long long) * NSHAPE * PART_MAXW);
    be.h2d(run_d, run_off.data(), sizeof(long long) * NSHAPE * PART_MAXW);
    for (int m = 1; m < NSHAPE; m++) if (tl.n[m] > 0) {
      KPartPack k{sp, pg.v, sl, tl, m, R, pg.gpos.template as<int>(), sot, dest + tl.t_begin[m], run_d + m * PART_MAXW, cur_d + m * PART_MAXW, sendb + sbase[m]};
      be.launch(k, (tl.n[m] + PART_TB - 1) / PART_TB, 128, KPartPack::smem_bytes());
    }
    be.xchg_begin();
    for (int m = 1; m < NSHAPE; m++) if (gN[m] > 0) {
      const long long RB = 8LL * part_rec_words(m, d);
      std::vector<long long> soff(W), scnt(W), roff(W), rcnt(W); long long o = 0;
      for (int h = 0; h < W; h++) { soff[h] = run_off[(size_t)m * PART_MAXW + h] * RB; scnt[h] = C(R, m, h) * RB; roff[h] = o * RB; rcnt[h] = C(h, m, R) * RB; o += C(h, m, R); }
      be.xchg_alltoallv(sendb + sbase[m], soff.data(), scnt.data(), recvb + rbase[m], roff.data(), rcnt.data());
      pstats.bytes_terms += (nrecv[m] - C(R, m, R)) * RB;
    }
    be.xchg_end();
    be.ev_record(15);
    // 4. owned terms in canonical order
    TermView tv; memset(&tv, 0, sizeof(tv));
    long long nt = 0, tA = 0, tpq = 0;
    for (int m = 0; m < NSHAPE; m++) {
      tv.n[m] = (int)nrecv[m]; tv.t_begin[m] = nt; tv.A_base[m] = tA; tv.pq_base[m] = tpq;
      nt += nrecv[m]; tA += nrecv[m] * m * d; tpq += nrecv[m] * m;
    }
    tv.t_begin[NSHAPE] = nt;
    tv.A = (double*)txA.ensure(sizeof(double) * (size_t)(tA + 8)); tv.p = (double*)txp.ensure(sizeof(double) * (size_t)(tpq + 8));
    tv.q = (double*)txq.ensure(sizeof(double) * (size_t)(tpq + 8)); tv.b = (double*)txb.ensure(sizeof(double) * (size_t)(nt + 1) * d);
    tv.meta = (SlotMeta*)txmeta.ensure(sizeof(SlotMeta) * (size_t)(nt + 1)); tv.cmap = (unsigned char*)txcmap.ensure((size_t)(nt + 1) * MAXM);
    unsigned long long* gk = (unsigned long long*)ptGk.ensure(sizeof(unsigned long long) * (size_t)(nt + 4));
    unsigned long long* gks = (unsigned long long*)ptGkS.ensure(sizeof(unsigned long long) * (size_t)(nt + 4));
    int* oi = (int*)ptIdx0.ensure(sizeof(int) * (size_t)((nt > gtot ? nt : gtot) + 4)); int* ord = (int*)ptOrd.ensure(sizeof(int) * (size_t)(nt + 4));
    unsigned long long* gidx = (unsigned long long*)ptGidx.ensure(sizeof(unsigned long long) * (size_t)(nt + 4));
    int* nold_d = (int*)ptNold.ensure(sizeof(int) * NSHAPE);
    be.memset(nold_d, 0, sizeof(int) * NSHAPE);
    for (int m = 1; m < NSHAPE; m++) if (tv.n[m] > 0) {
      const int n = tv.n[m], RW = part_rec_words(m, d); const long long tb = tv.t_begin[m];
      be.launch(KPartRecKeys{recvb + rbase[m], n, RW, gk + tb, oi + tb}, (n + 127) / 128, 128, 0);
      be.sort_pairs(gk + tb, gks + tb, oi + tb, ord + tb, n);
      be.launch(KPartCountOld{gks + tb, n, nold_d + m}, 1, 32, 0);
      be.launch(KPartUnpack{d, m, tv, recvb + rbase[m], ord + tb, gidx}, (n + PART_TB - 1) / PART_TB, 128, 0);
    }
    be.ev_record(16);
    // 5. import list: distinct (parent shape, home rank, alive rank) of the owned terms; every term learns its import index
    unsigned long long* ik = (unsigned long long*)ptIKey.ensure(sizeof(unsigned long long) * (size_t)(nt + 4));
    unsigned long long* iks = (unsigned long long*)ptIKeyS.ensure(sizeof(unsigned long long) * (size_t)(nt + 4));
    int* ii = (int*)ptIIdx.ensure(sizeof(int) * (size_t)(nt + 4)); int* iis = (int*)ptIIdxS.ensure(sizeof(int) * (size_t)(nt + 4));
    int* flag = (int*)ptFlag.ensure(sizeof(int) * (size_t)(nt + 4)); int* pos = (int*)ptPos.ensure(sizeof(int) * (size_t)(nt + 4));
    unsigned long long* ilist = (unsigned long long*)ptIList.ensure(sizeof(unsigned long long) * (size_t)(nt + 4));
    int* rlist = (int*)ptRList.ensure(sizeof(int) * (size_t)(nt + 4));
    int* seg_d = (int*)ptSeg.ensure(sizeof(int) * (NSHAPE * PART_MAXW + 4)); int* nimp_d = (int*)ptNImp.ensure(sizeof(int) * 4);
    be.memset(nimp_d, 0, sizeof(int) * 4);
    if (nt > 0) {
      const int nb = (int)((nt + 127) / 128);
      be.launch(KImportKeys{tv.meta, nt, ik, ii}, nb, 128, 0);
      be.sort_pairs(ik, iks, ii, iis, (int)nt);
      be.launch(KUniqFlags{iks, nt, flag}, nb, 128, 0);
      be.exclusive_scan(flag, pos, (int)nt);
      be.launch(KImportAssign{iks, iis, flag, pos, nt, tv.meta, ilist, rlist, nimp_d}, nb, 128, 0);
    }
    be.launch(KImportSegs{ilist, nimp_d, W, seg_d}, (NSHAPE * W + 1 + 127) / 128, 128, 0);
    std::vector<int> seg(NSHAPE * W + 1, 0), nold(NSHAPE, 0);
    be.d2h(seg.data(), seg_d, sizeof(int) * (NSHAPE * W + 1));
    be.d2h(nold.data(), nold_d, sizeof(int) * NSHAPE);
    for (int m = 0; m < NSHAPE; m++) tv.n_old[m] = nold[m];
    const int n_import = seg[NSHAPE * W];
    be.ev_record(17);
    // 6. requests to the home ranks, parent records back
    for (int m = 0; m < NSHAPE; m++) for (int h = 0; h < W; h++) mine[(size_t)m * W + h] = seg[m * W + h + 1] - seg[m * W + h];
    const std::vector<long long> rq = allgather_ll(mine);                 // rq[(requester * NSHAPE + m) * W + home]
    auto Q = [&](int req, int m, int home) { return rq[((size_t)req * NSHAPE + m) * W + home]; };
    std::vector<long long> nreq(NSHAPE, 0), rqb(NSHAPE + 1, 0), gimp(NSHAPE, 0);
    for (int m = 0; m < NSHAPE; m++) { for (int q = 0; q < W; q++) { nreq[m] += Q(q, m, R); for (int h = 0; h < W; h++) gimp[m] += Q(q, m, h); } rqb[m + 1] = rqb[m] + nreq[m]; }
    int* reqb = (int*)ptReq.ensure(sizeof(int) * (size_t)(rqb[NSHAPE] + 4));
    be.xchg_begin();
    for (int m = 1; m < NSHAPE; m++) if (gimp[m] > 0) {
      std::vector<long long> soff(W), scnt(W), roff(W), rcnt(W); long long o = 0;
      for (int h = 0; h < W; h++) { soff[h] = 4LL * seg[m * W + h]; scnt[h] = 4 * Q(R, m, h); roff[h] = 4 * o; rcnt[h] = 4 * Q(h, m, R); o += Q(h, m, R); }
      be.xchg_alltoallv(rlist, soff.data(), scnt.data(), reqb + rqb[m], roff.data(), rcnt.data());
    }
    be.xchg_end();
    const int Hcap = cell_count_central_half(max_shape, d);
    std::vector<ParentRecLayout> lay(NSHAPE);
    std::vector<long long> psb(NSHAPE + 1, 0), prb(NSHAPE + 1, 0), nimp(NSHAPE, 0);
    for (int m = 0; m < NSHAPE; m++) {
      if (m >= 1) { int mt = m + sp.npn; if (mt > max_shape) mt = max_shape; lay[m] = parent_rec_layout(m, d, with_tp ? cell_count_central_half(mt, d) : 0); } else memset(&lay[m], 0, sizeof(lay[m]));
      nimp[m] = seg[(m + 1) * W > NSHAPE * W ? NSHAPE * W : (m + 1) * W] - seg[m * W];
      psb[m + 1] = psb[m] + nreq[m] * lay[m].bytes; prb[m + 1] = prb[m] + nimp[m] * lay[m].bytes;
    }
    unsigned char* prs = (unsigned char*)ptPRecS.ensure((size_t)psb[NSHAPE] + 64);
    unsigned char* prr = (unsigned char*)ptPRecR.ensure((size_t)prb[NSHAPE] + 64);
    be.ev_record(18);
    for (int m = 1; m < NSHAPE; m++) if (nreq[m] > 0)
      be.launch(KImportPack{pg.v, ws, with_tp ? 1 : 0, m, d, lay[m], reqb + rqb[m], pg.gpos.template as<int>(), prs + psb[m]}, (int)nreq[m], 64, 0);
    be.xchg_begin();
    for (int m = 1; m < NSHAPE; m++) if (gimp[m] > 0) {
      const long long B = lay[m].bytes;
      std::vector<long long> soff(W), scnt(W), roff(W), rcnt(W); long long o = 0;
      for (int h = 0; h < W; h++) { soff[h] = o * B; scnt[h] = Q(h, m, R) * B; o += Q(h, m, R); roff[h] = (long long)(seg[m * W + h] - seg[m * W]) * B; rcnt[h] = Q(R, m, h) * B; }
      be.xchg_alltoallv(prs + psb[m], soff.data(), scnt.data(), prr + prb[m], roff.data(), rcnt.data());
      pstats.bytes_parents += (nimp[m] - Q(R, m, R)) * B;
    }
    be.xchg_end();
    be.ev_record(19);
    // 7. the import store: a generation store of its own, addressed like the local one
    std::vector<int> per(NSHAPE, 0);
    for (int m = 0; m < NSHAPE; m++) per[m] = (int)nimp[m];
    fill_gen_layout(imp, per);
    imp.v.alive = (int*)imp.alive.ensure(sizeof(int) * ((size_t)n_import + 4));
    imp.v.n_alive = n_import; imp.alive_per_shape = per;
    memset(iws, 0, sizeof(*iws));
    iws->sgnmask = (unsigned*)iwsSgn.ensure(sizeof(unsigned) * ((size_t)n_import + 4));
    iws->bxor = (unsigned*)iwsXor.ensure(sizeof(unsigned) * ((size_t)n_import + 4));
    iws->tpB_stride = Hcap;
    if (with_tp) {
      iws->tpB = (unsigned*)iwsTpB.ensure(sizeof(unsigned) * (size_t)n_import * Hcap + 16);
      iws->tpB_cells = (int*)iwsTpBc.ensure(sizeof(int) * ((size_t)n_import + 4));
    }
    int* igpos = (int*)ptIgpos.ensure(sizeof(int) * ((size_t)n_import + 4));
    for (int m = 1; m < NSHAPE; m++) if (nimp[m] > 0)
      be.launch(KImportUnpack{imp.v, *iws, with_tp ? 1 : 0, m, d, lay[m], prr + prb[m], seg[m * W], igpos}, (int)nimp[m], 64, 0);
    be.ev_record(20);
    pstats.imports = n_import; pstats.owned = (int)nt;
    for (int i = 0; i < 8; i++) pstats.ms[i] = be.ev_elapsed(12 + i, 13 + i);     // keys+splitters, destinations, term exchange, unpack, import list, requests, parent pack+exchange, import store
    *tvx = tv;
  }

  // in-place re-orientation masks of the imported parents, combined over the ranks between the two G-table phases
  void part_bxor_sync(ParentWs& iws) {
    int gn = 0; for (int m = 0; m < NSHAPE; m++) gn += g_alive_per_shape[m];
    const int n = imp.v.n_alive;
    unsigned* glob = (unsigned*)ptBxG.ensure(sizeof(unsigned) * ((size_t)gn + 4));
    const int* igpos = ptIgpos.template as<int>();
    be.memset(glob, 0, sizeof(unsigned) * (size_t)gn);
    if (n > 0) be.launch(KBxorScatter{n, iws.bxor, igpos, glob}, (n + 127) / 128, 128, 0);
    be.xchg_begin(); be.xchg_allreduce_u32(glob, (size_t)gn); be.xchg_end();
    if (n > 0) be.launch(KBxorGather{n, iws.bxor, igpos, glob}, (n + 127) / 128, 128, 0);
  }

  // global alive ranks of the new generation's local survivors + the survivor counts over all ranks
  void part_assign_gpos(GenStore& ng, const TermView& tv, const int* order_all, const int* gstart_all, const std::vector<int>& gstart_off) {
    const int W = be.shard.world, n_surv = ng.v.n_alive;
    unsigned long long* skey = (unsigned long long*)ptSKey.ensure(sizeof(unsigned long long) * ((size_t)n_surv + 4));
    if (n_surv > 0) {
      KSurvKeys k; memset(&k, 0, sizeof(k));
      k.next = ng.v; k.tv = tv; k.n_surv = n_surv; k.order_all = order_all; k.gstart_all = gstart_all; k.gidx = ptGidx.template as<unsigned long long>(); k.skey = skey;
      for (int m = 0; m < NSHAPE; m++) k.gstart_off[m] = gstart_off[m];
      be.launch(k, (n_surv + 127) / 128, 128, 0);
    }
    std::vector<long long> mine(NSHAPE, 0);
    for (int m = 0; m < NSHAPE; m++) mine[m] = ng.alive_per_shape[m];
    const std::vector<long long> cs = allgather_ll(mine);                 // cs[h * NSHAPE + m]
    KSurvRank k; memset(&k, 0, sizeof(k));
    std::vector<long long> roff(W), rcnt(W); long long tot = 0;
    for (int h = 0; h < W; h++) {
      long long nh = 0;
      for (int m = 0; m < NSHAPE; m++) { k.L.off[h][m] = tot + nh; k.L.cnt[h][m] = (int)cs[(size_t)h * NSHAPE + m]; nh += cs[(size_t)h * NSHAPE + m]; }
      roff[h] = 8 * tot; rcnt[h] = 8 * nh; tot += nh;
    }
    int base = 0;
    for (int m = 0; m < NSHAPE; m++) { int g = 0; for (int h = 0; h < W; h++) g += (int)cs[(size_t)h * NSHAPE + m]; k.L.shape_base[m] = base; g_alive_per_shape[m] = g; base += g; }
    unsigned long long* all = (unsigned long long*)ptSAll.ensure(sizeof(unsigned long long) * ((size_t)tot + 4));
    be.xchg_begin(); allgatherv(skey, 8LL * n_surv, all, roff, rcnt); be.xchg_end();
    int* gp = (int*)ng.gpos.ensure(sizeof(int) * ((size_t)n_surv + 4));
    if (n_surv > 0) {
      k.next = ng.v; k.n_surv = n_surv; k.W = W; k.skey = skey; k.all = all; k.gpos = gp;
      be.launch(k, (n_surv + 127) / 128, 128, 0);
    }
  }
  // host copy of the global alive ranks of the local survivors (tests, exporters: canonical order = ascending rank)
  int export_gpos(int* out, int cap) {
    const GenStore& g = gen[cur];
    const int n = g.v.n_alive;
    if (be.shard.world <= 1) { for (int i = 0; i < n && i < cap; i++) out[i] = i; return n; }
    if (n > 0 && out) be.d2h(out, g.gpos.p, sizeof(int) * (size_t)(n < cap ? n : cap));
    return n;
  }

  int step_general(double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H, double gamma,
                   const double* B, const double* u) {
    const bool with_tp = (master_step % p) == 0;
    StepParams sp = make_params(msmt, Phi, Gamma, beta, H, gamma, B, u, with_tp);
    GenStore& pg = gen[cur]; GenStore& ng = gen[1 - cur];
    const int n_alive = pg.v.n_alive;
    const int W = be.shard.world, R = be.shard.rank; (void)R;
    const bool part = W > 1;                 // partitioned estimator: this rank holds n_alive of the parents (possibly none)
    // A caller that rewinds master_step (cauchy_windows.hpp:538, 659 write the field) can drive a window deeper than declared:
    // the per-parent buffers hold max_shape rows, so a time propagation that would append beyond that is refused, not clamped.
    if (with_tp)
      for (int m = 1; m < NSHAPE; m++)
        if (pg.alive_per_shape[m] > 0 && m + sp.npn > max_shape) { error = "time propagation would give a term more than max_shape = " + std::to_string(max_shape) + " hyperplanes (window stepped past its declared depth)"; return -5; }
    const int nq = 1 + d + d * d;
    stats.parents = n_alive;
    if (part) { long long gn = 0; for (int m = 0; m < NSHAPE; m++) gn += g_alive_per_shape[m]; stats.parents = gn; }
    int* diag = (int*)diagBuf.ensure(sizeof(int) * (16 + NSHAPE + 8)); be.memset(diag, 0, sizeof(int) * (16 + NSHAPE + 8));   // [16] diagnostics, then the survivor bounds per shape
    if (max_shape <= 16 && n_alive > 0) {        // rank structures of the parents' tables (lookups without binary search)
      int mmax = 1; for (int m = 1; m < NSHAPE; m++) if (pg.alive_per_shape[m] > 0) mmax = m;
      be.launch(KBuildRank{pg.v}, n_alive, 64, KBuildRank::smem_bytes(rank_words(mmax), 64));
    }
    // per-phase host timers synchronise the stream; they are off unless mce_options.phase_timing is set
    auto tic = [&]() { return phase_timing ? be.tic() : 0.0; };
    auto toc = [&](double t) { return phase_timing ? be.toc(t) : 0.0; };
    double tph = tic();

    // ---- K1/K2: time propagation, Gamma coalignment, B^{k|k-1} ----
    ParentWs ws; memset(&ws, 0, sizeof(ws));
    ws.sgnmask = (unsigned*)wsSgn.ensure(sizeof(unsigned) * (n_alive + 4));
    ws.bxor = (unsigned*)wsXor.ensure(sizeof(unsigned) * (n_alive + 4));
    const int Hcap = cell_count_central_half(max_shape, d);
    if (with_tp) {
      ws.A = (double*)wsA.ensure(sizeof(double) * (size_t)n_alive * max_shape * d);
      ws.p = (double*)wsp.ensure(sizeof(double) * (size_t)n_alive * max_shape);
      ws.b = (double*)wsb.ensure(sizeof(double) * (size_t)n_alive * d);
      ws.m_tp = (unsigned char*)wsm.ensure((size_t)n_alive + 16);
      be.launch(KTimeProp{sp, pg.v, ws}, (n_alive + 127) / 128, 128, 0);
      if (!skip_post_mu) {
        const int tp0 = 0, tp1 = n_alive;
        ws.tpB_stride = Hcap;
        ws.tpB = (unsigned*)wsTpB.ensure(sizeof(unsigned) * (size_t)n_alive * Hcap + 16);
        ws.tpB_cells = (int*)wsTpBc.ensure(sizeof(int) * ((size_t)n_alive + 4));
        int max_mtp = 0;
        for (int m = 1; m < NSHAPE; m++) if (pg.alive_per_shape[m] > 0) max_mtp = m + sp.npn;
        if (max_mtp > max_shape) max_mtp = max_shape;
        const int cgen = cell_count_general(max_mtp, d);
        const int acc_cap = next_pow2(cgen + 1), vis_cap = next_pow2(DCE_STORAGE_MULT * cgen);
        const int nth = 128;
        if (max_shape <= 16) {
          const int NWt = (1 << max_shape) / 32 > 1 ? (1 << max_shape) / 32 : 1;
          switch (d) {        // the vertex solver is specialised per state dimension (its d x d systems live in registers)
            case 2: be.launch(KTpDce2T<2>{sp, pg.v, ws, NWt, diag, tp0}, tp1 - tp0, nth, KTpDce2T<2>::smem_bytes(NWt, nth, d)); break;
            case 3: be.launch(KTpDce2T<3>{sp, pg.v, ws, NWt, diag, tp0}, tp1 - tp0, nth, KTpDce2T<3>::smem_bytes(NWt, nth, d)); break;
            case 4: be.launch(KTpDce2T<4>{sp, pg.v, ws, NWt, diag, tp0}, tp1 - tp0, nth, KTpDce2T<4>::smem_bytes(NWt, nth, d)); break;
            case 5: be.launch(KTpDce2T<5>{sp, pg.v, ws, NWt, diag, tp0}, tp1 - tp0, nth, KTpDce2T<5>::smem_bytes(NWt, nth, d)); break;
            case 6: be.launch(KTpDce2T<6>{sp, pg.v, ws, NWt, diag, tp0}, tp1 - tp0, nth, KTpDce2T<6>::smem_bytes(NWt, nth, d)); break;
            case 7: be.launch(KTpDce2T<7>{sp, pg.v, ws, NWt, diag, tp0}, tp1 - tp0, nth, KTpDce2T<7>::smem_bytes(NWt, nth, d)); break;
            default: be.launch(KTpDce2T<8>{sp, pg.v, ws, NWt, diag, tp0}, tp1 - tp0, nth, KTpDce2T<8>::smem_bytes(NWt, nth, d)); break;
          }
        } else {
          be.launch(KTpDce{sp, pg.v, ws, vis_cap, acc_cap, diag, tp0}, tp1 - tp0, nth, KTpDce::smem_bytes(vis_cap, acc_cap, nth));
        }
      }
    }
    stats.ms_tp = toc(tph); tph = tic();

    // ---- K3/K4: measurement update, moment contributions, MU coalignment ----
    SlotView sl; memset(&sl, 0, sizeof(sl));
    long long nslots = 0, offA = 0, offpq = 0; int rank = 0;
    for (int m = 0; m < NSHAPE; m++) {
      sl.par_begin[m] = rank; sl.slot_begin[m] = nslots; sl.A_off[m] = offA; sl.pq_off[m] = offpq;
      const int n = pg.alive_per_shape[m];
      int MT = m + (with_tp ? sp.npn : 0);
      if (MT > max_shape) MT = max_shape;
      sl.MT[m] = MT;
      rank += n; nslots += (long long)n * (MT + 1); offA += (long long)n * (MT + 1) * MT * d; offpq += (long long)n * (MT + 1) * MT;
    }
    sl.par_begin[NSHAPE] = rank; sl.slot_begin[NSHAPE] = nslots; sl.n_slots = nslots;
    stats.slots = nslots; last_nslots = nslots;
    sl.meta = (SlotMeta*)slmeta.ensure(sizeof(SlotMeta) * (size_t)(nslots + 1));
    sl.g = (cplx*)slg.ensure(sizeof(cplx) * (size_t)(nslots + 8));
    sl.y = (double*)sly.ensure(sizeof(double) * (size_t)(nslots + 1) * 2 * d);
    sl.cmap = (unsigned char*)slcmap.ensure(skip_post_mu ? 64 : (size_t)(nslots + 1) * MAXM);
    if (!skip_post_mu) {
      sl.A = (double*)slA.ensure(sizeof(double) * (size_t)(offA + 8));
      sl.p = (double*)slp.ensure(sizeof(double) * (size_t)(offpq + 8));
      sl.q = (double*)slq.ensure(sizeof(double) * (size_t)(offpq + 8));
      sl.b = (double*)slb.ensure(sizeof(double) * (size_t)(nslots + 1) * d);
    }
    for (int m = 1; m < NSHAPE; m++) {
      const long long n = (long long)pg.alive_per_shape[m] * (sl.MT[m] + 1);
      if (n > 0) {
        const int npar = pg.alive_per_shape[m];
        be.launch(KMsmtUpdate2{sp, pg.v, ws, sl, m}, (npar + MU_PB - 1) / MU_PB, 128, KMsmtUpdate2::smem_bytes(sl.MT[m], d));
      }
    }
    stats.ms_mu = toc(tph); tph = tic();

    // ---- moments (K3 tail): serial-order fz, two-level mean/covariance sums ----
    // The sums run on the side stream: nothing before the G-table build needs them, so they overlap regroup + FTR.
    double* mom = (double*)momOut.ensure(sizeof(double) * 4 * nq + 16);
    // The moment kernel isolates its accumulator warp on one scheduler partition (warp id % 4); that mapping only holds when
    // its CTAs are placed on idle SMs, so the main stream is drained first (measured: 8.4 ms instead of 11.6 ms at 1.1 M slots).
    // partitioned estimator, ordered moments: the sums run over ALL ranks' slots in the reference's order (bit-exact)
    const cplx* mom_g = sl.g; const double* mom_y = sl.y; long long mom_n = nslots;
    const double* scan_g = nullptr; long long scan_n = 0; int scan_stride = 2;
    if (part) {
      if (moments_mode == 0) { cplx* gg; double* gy; mom_n = part_gather_slots(sl, pg, &gg, &gy); mom_g = gg; mom_y = gy; stats.slots = mom_n; }
      if (moments_mode == 2) { scan_n = part_gather_re(sl, pg, &scan_g); scan_stride = 1; stats.slots = scan_n; }
    }
    // Large steps: the G-table build only needs G_SCALE_FACTOR = 1 / (2 pi Re fz), and Re fz alone is available early -- KSumScan adds that one chain
    // as an exact parallel scan, 3x faster than the dependent chain (bit-identical; checked against the chain's own Re fz below).  The G-table kernels
    // then start as soon as the term reduction is done, and the 2 (1 + d + d^2) serial chains finish beside them instead of in front of them.
    const bool early_scale = !part && !skip_post_mu && !fast_moments && nslots >= early_scale_slots;
    be.ev_record(6);
    be.sync();
    be.side_begin();
    be.ev_record_side(4);
    if (early_scale) { launch_sum_scan(true, (const double*)sl.g, 2, nslots, mom + 2 * nq); be.ev_record_side(9); }
    if ((fast_moments && !part && nslots >= fast_moments_slots) || (part && moments_mode == 2)) {
      // mce_options.fast_moments on one GPU / hybrid on a partitioned estimator: no dependent chain.  Every sum is a fixed-shape two-level reduction (deterministic
      // for a given partition; a rank's sums are added to the other ranks' in rank order anyway) EXCEPT the one whose bits matter downstream: Re fz comes from the
      // exact scan -- over all ranks' slots when partitioned -- so G_SCALE_FACTOR, and with it every count, key and G value, stays bit-identical.
      if (!part) { scan_g = (const double*)sl.g; scan_stride = 2; scan_n = nslots; }
      const int nblk = (int)((nslots + MOM_CHUNK - 1) / MOM_CHUNK);
      double* partial = (double*)momPartial.ensure(sizeof(double) * (size_t)(nblk > 0 ? nblk : 1) * 2 * nq + 64);
      if (nblk > 0) be.launch_side(KMomentsPartial{sl.g, sl.y, nslots, d, partial, 1}, nblk, 128, sizeof(double) * 2 * 128);
      be.launch_side(KMomentsFinal{partial, nblk, nq, mom}, (2 * nq + 127) / 128, 128, 0);
    } else {
      be.launch_side(KMomentsSerial{mom_g, mom_y, mom_n, d, mom}, (nq + MOM_QB - 1) / MOM_QB, 512, KMomentsSerial::smem_bytes(d));
    }
    if (scan_g) {       // one CTA; on the window's last step nothing else waits on the main stream, so it runs there, beside the rank's own chains
      launch_sum_scan(!skip_post_mu, scan_g, scan_stride, scan_n, mom + 2 * nq);
    }
    be.ev_record_side(5);
    double early_refz = 0, chain_refz = 0;
    auto early_gscale = [&]() {        // Re fz from the scan -> the scale of the new G values
      double sc2[2];
      be.ev_wait(9);
      be.d2h(sc2, mom + 2 * nq, sizeof(sc2));
      early_refz = sc2[0]; scan_restarts = (int)sc2[1];
      sp.gscale = (1.0 / (2.0 * M_PI)) / early_refz;       // finalize_moments' expression
    };
    auto finish_moments = [&]() {
      std::vector<double> raw(2 * nq);
      be.side_join();
      stats.ev_moments_ms = be.ev_elapsed(4, 5); stats.ev_mu_ms = be.ev_elapsed(0, 6);
      be.d2h(raw.data(), mom, sizeof(double) * 2 * nq);
      if (!part && fast_moments && nslots >= fast_moments_slots) { double sc2[2]; be.d2h(sc2, mom + 2 * nq, sizeof(sc2)); raw[0] = sc2[0]; scan_restarts = (int)sc2[1]; }
      if (part && moments_mode != 0) {        // per-rank sums, added in rank order: the same result on every rank for a given world size
        std::vector<long long> mine(2 * nq);
        memcpy(mine.data(), raw.data(), sizeof(double) * 2 * nq);
        const std::vector<long long> all = allgather_ll(mine);
        for (int i = 0; i < 2 * nq; i++) {
          double acc = 0;
          for (int h = 0; h < W; h++) { double v; memcpy(&v, &all[(size_t)h * 2 * nq + i], sizeof(double)); acc += v; }
          raw[i] = acc;
        }
        if (moments_mode == 2) {                // Re fz as the reference's chain over ALL slots gives it, bit for bit
          double sc2[2]; be.d2h(sc2, mom + 2 * nq, sizeof(sc2));
          raw[0] = sc2[0]; scan_restarts = (int)sc2[1];
        }
      }
      chain_refz = raw[0];
      finalize_moments(raw.data(), true);
      sp.gscale = G_SCALE_FACTOR;
    };

    // ---- canonical ranks of the new terms inside their new shapes ----
    const int nchunks = (int)((nslots + RANK_CHUNK - 1) / RANK_CHUNK);
    int* counts = (int*)rankCounts.ensure(sizeof(int) * (size_t)nchunks * 2 * NSHAPE + 64);
    int* totals = (int*)rankTotals.ensure(sizeof(int) * 2 * NSHAPE + 64);
    be.launch(KRankCount{sl, nchunks, counts}, nchunks, RANK_CHUNK, sizeof(int) * 2 * NSHAPE);
    be.launch(KRankScan{nchunks, counts, totals}, 2 * NSHAPE, 64, KRankScan::smem_bytes(64));
    std::vector<int> tot(2 * NSHAPE);
    be.d2h(tot.data(), totals, sizeof(int) * 2 * NSHAPE);
    stats.ms_moments = toc(tph); tph = tic();
    std::vector<long long> all_tot;          // partitioned estimator: every rank's (old, child) counts per new shape
    if (part) { std::vector<long long> mine(tot.begin(), tot.end()); all_tot = allgather_ll(mine); }

    TermView tv; memset(&tv, 0, sizeof(tv));
    long long nterms = 0, tA = 0, tpq = 0;
    std::fill(muc_per_shape.begin(), muc_per_shape.end(), 0);
    for (int m = 0; m < NSHAPE; m++) {
      tv.n_old[m] = tot[m]; tv.n[m] = tot[m] + tot[NSHAPE + m];
      tv.t_begin[m] = nterms; tv.A_base[m] = tA; tv.pq_base[m] = tpq;
      nterms += tv.n[m]; tA += (long long)tv.n[m] * m * d; tpq += (long long)tv.n[m] * m;
      if (m < shape_range) muc_per_shape[m] = tv.n[m];
    }
    tv.t_begin[NSHAPE] = nterms;
    long long g_nterms = nterms;
    if (part) {
      g_nterms = 0;
      for (int m = 0; m < NSHAPE; m++) {
        long long g = 0; for (int h = 0; h < W; h++) g += all_tot[(size_t)h * 2 * NSHAPE + m] + all_tot[(size_t)h * 2 * NSHAPE + NSHAPE + m];
        if (m < shape_range) muc_per_shape[m] = (int)g;
        g_nterms += g;
      }
    }
    Nt_muc = (int)g_nterms; stats.terms_after_muc = g_nterms;
    if (skip_post_mu) {          // est:732-733, 806-828: only the counts change on the window's last step
      finish_moments();
      terms_per_shape = muc_per_shape; Nt = (int)g_nterms; finished = true;
      return 0;
    }

    tv.A = (double*)tvA.ensure(sizeof(double) * (size_t)(tA + 8));
    tv.p = (double*)tvp.ensure(sizeof(double) * (size_t)(tpq + 8));
    tv.q = (double*)tvq.ensure(sizeof(double) * (size_t)(tpq + 8));
    tv.b = (double*)tvb.ensure(sizeof(double) * (size_t)(nterms + 1) * d);
    tv.meta = (SlotMeta*)tvmeta.ensure(sizeof(SlotMeta) * (size_t)(nterms + 1));
    tv.cmap = (unsigned char*)tvcmap.ensure((size_t)(nterms + 1) * MAXM);
    long long* sot = (long long*)slotOfTerm.ensure(sizeof(long long) * (size_t)(nterms + 1));
    be.launch(KRegroup{sp, sl, tv, nchunks, counts, sot}, nchunks, RANK_CHUNK, KRegroup::smem_bytes());
    ParentWs gws = ws;                       // the parents the G-table kernel reads: the local ones, or the imported ones
    if (part) {                              // the terms move to the owners of their reduction keys; from here on `tv` holds the owned terms
      TermView tvx;
      part_exchange(sp, sl, tv, sot, all_tot, with_tp, pg, ws, &tvx, &gws);
      tv = tvx; nterms = tv.t_begin[NSHAPE];
    }
    const GenView& pv = part ? imp.v : pg.v;
    stats.ms_regroup = toc(tph); tph = tic();

    be.ev_record(7);
    // ---- K5/K6: fast term reduction per new shape, reduction groups ----
    int* F_all = (int*)ftrF.ensure(sizeof(int) * (size_t)(nterms + 4));
    unsigned char* wide_all = (unsigned char*)ftrWide.ensure((size_t)nterms + 16);
    int* order_all = (int*)grpOrder.ensure(sizeof(int) * (size_t)(nterms + 4));
    int* gstart_all = (int*)grpStart.ensure(sizeof(int) * (size_t)(nterms + NSHAPE + 4));
    // The shapes advance in lock-step (every stage is launched for all shapes before its counters are read back), so a step
    // costs 2 + max_rounds host round trips instead of (2 + rounds) per shape; scratch is laid out per shape by term offset.
    unsigned long long* k0_all = (unsigned long long*)scratchK0.ensure(sizeof(unsigned long long) * (size_t)(nterms + 4));
    unsigned long long* k1_all = (unsigned long long*)scratchK1.ensure(sizeof(unsigned long long) * (size_t)(nterms + 4));
    int* i0_all = (int*)scratchI0.ensure(sizeof(int) * (size_t)(nterms + 4));
    int* i1_all = (int*)scratchI1.ensure(sizeof(int) * (size_t)(nterms + 4));
    int* i2_all = (int*)scratchI2.ensure(sizeof(int) * (size_t)(nterms + 4));
    int* i3_all = (int*)scratchI3.ensure(sizeof(int) * (size_t)(nterms + 4));
    const size_t unk_bytes = (sizeof(unsigned long long) + sizeof(int)) * NSHAPE;
    unsigned long long* dens_d = (unsigned long long*)unkBuf.ensure(unk_bytes + 64);    // [NSHAPE] window populations
    int* nu_d = (int*)(dens_d + NSHAPE);                                                // [NSHAPE] undecided terms of the round
    std::vector<int> n_groups(NSHAPE, 0), n_phase1(NSHAPE, 0), gstart_off(NSHAPE, 0);
    if (capture) cap.clear();
    // split-group bookkeeping (KBigGroups): per (phase, shape) a region of group/part descriptors and a counter block
    std::vector<long long> big_slot_base(2 * NSHAPE, 0), big_part_base(2 * NSHAPE, 0);
    long long big_slots = 0, big_parts = 0;
    for (int ph = 0; ph < 2; ph++)
      for (int m = 1; m < NSHAPE; m++) {
        big_slot_base[ph * NSHAPE + m] = big_slots; big_part_base[ph * NSHAPE + m] = big_parts;
        big_slots += tv.n[m] / big_T + 1; big_parts += tv.n[m] / BIG_PART + tv.n[m] / big_T + 2;
      }
    BigGroup* bgroups = (BigGroup*)bigGroups.ensure(sizeof(BigGroup) * (size_t)big_slots);
    BigPart* bparts = (BigPart*)bigParts.ensure(sizeof(BigPart) * (size_t)big_parts);
    const size_t bigcnt_bytes = (sizeof(unsigned long long) * 3 + sizeof(int) * 2) * 2 * NSHAPE + sizeof(int) * 2 * NSHAPE;
    unsigned long long* bcnt64 = (unsigned long long*)bigCnt.ensure(bigcnt_bytes);      // [2*NSHAPE][3], then int [2*NSHAPE][2], then cr
    int* bcnt = (int*)(bcnt64 + 3 * 2 * NSHAPE);
    int* cr_d = bcnt + 2 * 2 * NSHAPE;                                                  // [NSHAPE][2] roots, old-term roots (KCountRoots)
    be.memset(bcnt64, 0, bigcnt_bytes);
    std::vector<int> shapes, reg_shapes;          // shapes present this step; those too large for the single-CTA kernel
    KFtrSmall ks; memset(&ks, 0, sizeof(ks));
    int n_small = 0;
    for (int m = 1; m < NSHAPE; m++) if (tv.n[m] > 0) {
      shapes.push_back(m);
      if (tv.n[m] <= FTR_SMALL_N) ks.shapes[n_small++] = m; else reg_shapes.push_back(m);
    }
    be.memset(dens_d, 0, unk_bytes);
    if (n_small > 0) {       // small shapes: the whole reduction in one launch, no host round trip
      ks.tv = tv; ks.sp = sp; ks.k1_all = k1_all; ks.i1_all = i1_all; ks.wide_all = wide_all; ks.F_all = F_all; ks.order_all = order_all;
      ks.gstart_all = gstart_all; ks.cr = cr_d;
      be.launch(ks, n_small, 512, KFtrSmall::smem_bytes());
    }
    for (int m : reg_shapes) {
      const int n = tv.n[m], nb = (n + 127) / 128; const long long tb = tv.t_begin[m];
      be.launch(KFtrKeys{tv, m, d, tr_order[0], k0_all + tb, i0_all + tb, F_all + tb}, nb, 128, 0);
      be.sort_pairs(k0_all + tb, k1_all + tb, i0_all + tb, i1_all + tb, n);      // k1 = sorted keys, i1 = term index at each sorted position
      be.launch(KFtrWide{tv, m, d, tr_order[0], k1_all + tb, wide_all + tb, dens_d + m}, nb, 128, 0);
    }
    std::vector<unsigned long long> dens(NSHAPE, 0);
    if (!reg_shapes.empty()) be.d2h(dens.data(), dens_d, sizeof(unsigned long long) * NSHAPE);
    std::vector<int> active = reg_shapes, nu(NSHAPE, 0);
    int rounds = 0;
    while (!active.empty()) {
      // rounds that find nothing to do are cheap, host round trips are not: the first batch runs several rounds per readback
      const int batch = rounds == 0 ? (nterms < 65536 ? 3 : 2) : 1;
      for (int b = 0; b < batch; b++) {
        be.memset(nu_d, 0, sizeof(int) * NSHAPE);
        for (int m : active) {
          const int n = tv.n[m], nb = (n + 127) / 128; const long long tb = tv.t_begin[m];
          // mean epsilon-window population on the sorted axis decides the round kernel: dense clusters (many candidates per
          // term) amortise the shared-memory staging of the tiled kernel, sparse data is faster with direct scans
          const bool tiled = (double)dens[m] * 16.0 / (double)n > 24.0;
          if (tiled) be.launch(KFtrRoundTiled{tv, sp, m, k1_all + tb, i1_all + tb, wide_all + tb, F_all + tb, nu_d + m}, (n + FTR_TB - 1) / FTR_TB, FTR_TB, KFtrRoundTiled::smem_bytes(m, d));
          else be.launch(KFtrRound{tv, sp, m, k1_all + tb, i1_all + tb, wide_all + tb, F_all + tb, nu_d + m}, nb, 128, 0);
        }
        rounds++;
      }
      be.d2h(nu.data(), nu_d, sizeof(int) * NSHAPE);
      std::vector<int> still;
      for (int m : active) if (nu[m] != 0) still.push_back(m);
      active.swap(still);
      if (rounds > (int)nterms + 4) { be.side_join(); error = "FTR resolution did not converge"; return -6; }
    }
    stats.ftr_rounds_max = rounds;
    if (capture) for (int m : shapes) capture_shape(tv, m, F_all + tv.t_begin[m], gws, with_tp, part ? imp : pg);
    for (int m : shapes) {
      const int n = tv.n[m], nb = (n + 127) / 128; const long long tb = tv.t_begin[m];
      int* gstart = gstart_all + tb + m;                       // at most n + 1 entries per shape
      gstart_off[m] = (int)(tb + m);
      if (n > FTR_SMALL_N) {
        be.launch(KRootKeys{n, F_all + tb, k0_all + tb, i0_all + tb}, nb, 128, 0);
        be.sort_pairs(k0_all + tb, k1_all + tb, i0_all + tb, order_all + tb, n);
        be.launch(KGroupHeads{n, F_all + tb, order_all + tb, i2_all + tb}, nb, 128, 0);
        be.exclusive_scan(i2_all + tb, i3_all + tb, n);
        be.launch(KCountRoots{n, tv.n_old[m], F_all + tb, cr_d + 2 * m}, nb, 128, 2 * sizeof(int));
        be.launch(KGroupFill{n, i2_all + tb, i3_all + tb, gstart, cr_d + 2 * m}, nb, 128, 0);
      }
      if (max_shape <= 16 && n > big_T) {
        const int Hm = cell_count_central_half(m, d);
        for (int ph = 0; ph < 2; ph++) {
          const int ix = ph * NSHAPE + m;
          be.launch(KBigGroups{gstart, cr_d + 2 * m, ph, big_T, Hm, bgroups + big_slot_base[ix], bparts + big_part_base[ix], bcnt + 2 * ix, bcnt64 + 3 * ix, 0, 1}, nb, 128, 0);
        }
      }
    }
    be.ev_record(8);
    stats.ms_ftr = toc(tph); tph = tic();

    // ---- K7/K8: child B-tables and G-tables, one CTA per reduction group ----
    if (early_scale) early_gscale(); else finish_moments();            // G_SCALE_FACTOR = 1 / (2 pi Re fz) scales every new G (flat:227)
    stats.ms_moments += toc(tph); tph = tic();
    // one D2H brings the root counts of every shape (KCountRoots) and the split-group counters (KBigGroups)
    std::vector<unsigned long long> h64(3 * 2 * NSHAPE, 0); std::vector<int> h32(2 * 2 * NSHAPE, 0), cr(2 * NSHAPE, 0);
    {
      std::vector<unsigned char> hb(bigcnt_bytes);
      be.d2h(hb.data(), bcnt64, bigcnt_bytes);
      memcpy(h64.data(), hb.data(), sizeof(unsigned long long) * h64.size());
      memcpy(h32.data(), hb.data() + sizeof(unsigned long long) * h64.size(), sizeof(int) * h32.size());
      memcpy(cr.data(), hb.data() + sizeof(unsigned long long) * h64.size() + sizeof(int) * h32.size(), sizeof(int) * cr.size());
    }
    for (int m : shapes) {
      n_groups[m] = cr[2 * m];
      n_phase1[m] = cr[2 * m + 1];      // groups rooted at an old term come first (roots ascend, old terms precede children)
    }
    fill_gen_layout(ng, n_groups);
    unsigned char* aflag = (unsigned char*)aliveFlag.ensure((size_t)ng.v.n_groups + 16);
    const int HC2 = next_pow2(Hcap < 4 ? 4 : Hcap);
    const size_t gsm = KGTable::smem_bytes(HC2);
    long long total_groups = 0;
    // split groups: sizes of the addend scratch, scratch bases per (phase, shape)
    std::vector<long long> rows_base(2 * NSHAPE, 0), flags_base(2 * NSHAPE, 0), keys_base(2 * NSHAPE, 0);
    bool split = false;
    cplx* brows = nullptr; int* bflags = nullptr; unsigned* bkeys = nullptr;
    if (max_shape <= 16) {
      long long tr = 0, tf = 0, tk = 0, nbig = 0;
      for (int ix = 0; ix < 2 * NSHAPE; ix++) {
        rows_base[ix] = tr; flags_base[ix] = tf; keys_base[ix] = tk;
        tr += (long long)h64[3 * ix]; tf += (long long)h64[3 * ix + 1]; tk += (long long)h64[3 * ix + 2]; nbig += h32[2 * ix];
      }
      stats.big_groups = (int)nbig;
      if (nbig > 0 && tr * (long long)sizeof(cplx) <= big_scratch_cap) {
        split = true;
        brows = (cplx*)bigRows.ensure(sizeof(cplx) * (size_t)tr);
        bflags = (int*)bigFlags.ensure(sizeof(int) * (size_t)tf);
        bkeys = (unsigned*)bigKeys.ensure(sizeof(unsigned) * (size_t)tk);
        be.memset(bflags, 0, sizeof(int) * (size_t)tf);
      }
    }
    be.ev_record(2);
    for (int phase = 0; phase < 2; phase++) {
      for (int m = 1; m < NSHAPE; m++) {
        if (n_groups[m] == 0) continue;
        const int g0 = phase == 0 ? 0 : n_phase1[m], g1 = phase == 0 ? n_phase1[m] : n_groups[m];
        if (g1 <= g0) continue;
        total_groups += g1 - g0;
        const int lo = g0, hi = g1, gshift = 0;
        const int Hm = cell_count_central_half(m, d);
        int nth = Hm <= 32 ? 32 : (Hm <= 64 ? 64 : 128);
        { static const char* ev = getenv("MCE_G2_NT"); if (ev && max_shape <= 16) { const int v = atoi(ev); if (v == 32 || v == 64 || v == 128) nth = v < nth ? v : nth; } }
        if (max_shape <= 16) {
          const int NW = (1 << max_shape) / 32 > 1 ? (1 << max_shape) / 32 : 1;
          const int ix = phase * NSHAPE + m;
          BigArgs ba; memset(&ba, 0, sizeof(ba));
          const int nbg = split ? h32[2 * ix] : 0, nbp = split ? h32[2 * ix + 1] : 0;
          if (nbg > 0) {
            ba.groups = bgroups + big_slot_base[ix]; ba.parts = bparts + big_part_base[ix];
            ba.rows = brows + rows_base[ix]; ba.flags = bflags + flags_base[ix]; ba.keys = bkeys + keys_base[ix]; ba.row_stride = Hm;
          }
          // the standard variant wherever its two value tables fit; the lean one (half the bytes per cell, ~10 % slower) beyond
          const bool lean = lean_groups || KGTable2::smem_bytes(Hm, NW) > (size_t)(227 * 1024);
          const int* ord = order_all + tv.t_begin[m]; const int* gst = gstart_all + gstart_off[m];
          auto launch_all = [&](auto lean_tag) {
            constexpr bool L = decltype(lean_tag)::value;
            const size_t smb = KGTable2T<G2_NORMAL, L>::smem_bytes(Hm, NW);      // per-cell arrays sized by this shape's largest table: what is left of the SM's 256 KB is L1 for the parent tables
            be.launch(KGTable2T<G2_NORMAL, L>{sp, pv, ng.v, gws, tv, m, lo, ord, gst, Hm, NW, aflag, diag, split ? big_T : 0x7fffffff, ba, gshift}, hi - lo, nth, smb);
            if (nbg > 0) {       // root election, then the members in parts, then the ordered sums of the stored addends
              be.launch(KGTable2T<G2_BIG_ROOT, L>{sp, pv, ng.v, gws, tv, m, lo, ord, gst, Hm, NW, aflag, diag, big_T, ba, gshift}, nbg, nth, smb);
              be.launch(KGTable2T<G2_BIG_PARTS, L>{sp, pv, ng.v, gws, tv, m, lo, ord, gst, Hm, NW, aflag, diag, big_T, ba, gshift}, nbp, nth, smb);
              be.launch(KGTable2T<G2_BIG_FINAL, L>{sp, pv, ng.v, gws, tv, m, lo, ord, gst, Hm, NW, aflag, diag, big_T, ba, gshift}, nbg, nth, smb);
            }
          };
          if (lean) { launch_all(std::true_type{}); stats.gtable_lean_launches++; } else launch_all(std::false_type{});
        } else {
          KGTable k{sp, pv, ng.v, gws, tv, m, g0, order_all + tv.t_begin[m], gstart_all + gstart_off[m], HC2, aflag, diag};
          be.launch(k, g1 - g0, nth, gsm);
        }
        stats.gtable_launches++;
      }
      if (part && phase == 0) part_bxor_sync(gws);      // phase 1 reads the re-orientation masks phase 0 wrote, on whichever rank
    }
    be.ev_record(3);
    if (early_scale) {           // the full set of sums, finished beside the G-table kernels; its Re fz must be the scan's, bit for bit
      finish_moments();
      if (memcmp(&early_refz, &chain_refz, sizeof(double)) != 0) { error = "the exact scan of Re fz disagrees with the serial chain"; return -7; }
    }
    stats.ev_gtable_ms = be.ev_elapsed(2, 3); stats.ev_ftr_ms = be.ev_elapsed(7, 8);
    stats.groups = total_groups;
    stats.ms_gtable = toc(tph); tph = tic();

    // ---- K9: survivor list of the new generation, parent/child generation swap (util:895) ----
    const int ngr = ng.v.n_groups;
    int n_surv = 0;
    std::fill(ng.alive_per_shape.begin(), ng.alive_per_shape.end(), 0);
    if (ngr > 0) {
      int* fi = (int*)scratchI0.ensure(sizeof(int) * (size_t)(ngr + 4));
      int* fr = (int*)scratchI1.ensure(sizeof(int) * (size_t)(ngr + 4));
      be.launch(KFlagsToInt{ngr, aflag, fi}, (ngr + 127) / 128, 128, 0);
      be.exclusive_scan(fi, fr, ngr);
      be.launch(KAliveCompact{ngr, aflag, fr, ng.v.alive}, (ngr + 127) / 128, 128, 0);
      be.launch(KShapeBounds{ng.v, ngr, fi, fr, diag + 16}, 1, 64, 0);
    }
    std::vector<int> hd(16 + NSHAPE + 1, 0);
    be.d2h(hd.data(), diag, sizeof(int) * hd.size());          // diagnostics + survivor bounds in one read
    if (ngr > 0) {
      const int* bounds = hd.data() + 16;
      for (int m = 1; m < NSHAPE; m++) ng.alive_per_shape[m] = bounds[m + 1] - bounds[m];
      n_surv = bounds[NSHAPE];
    }
    ng.v.n_alive = n_surv;
    if (part) part_assign_gpos(ng, tv, order_all, gstart_all, gstart_off);
    stats.diag_alias = hd[0]; stats.diag_hash = hd[1];
    stats.cells_parents = pg.sum_cells; ng.sum_cells = (long long)(unsigned)hd[2] | ((long long)hd[3] << 32); stats.cells_survivors = ng.sum_cells;
    stats.survivors = n_surv;
    std::fill(terms_per_shape.begin(), terms_per_shape.end(), 0);
    for (int m = 1; m < shape_range; m++) terms_per_shape[m] = part ? g_alive_per_shape[m] : ng.alive_per_shape[m];
    Nt = n_surv;
    if (part) { Nt = 0; for (int m = 0; m < NSHAPE; m++) Nt += g_alive_per_shape[m]; stats.survivors = Nt; }
    cur = 1 - cur;
    if (print_basic_info && !part) post_ftr_moments(sp); else fz = make_cplx(1, 0);           // est:1166-1176 (quirk A.9 iii)
    stats.ms_compact = toc(tph);
    // algorithmic bytes with the ACTUAL table sizes: every compulsory input read once, every output
    // written once, the post-MUC term payload written + read once (FTR is a global barrier). A table cell is
    // 4 B key + 16 B complex value in this layout (the reference's padded KeyCValue + B entry is 28 B).
    {
      long long term_in = 0, payload = 0, term_out = 0;
      for (int m = 1; m < NSHAPE; m++) {
        term_in += (long long)pg.alive_per_shape[m] * 8LL * (m * d + m + d);
        payload += (long long)tv.n[m] * (8LL * (m * d + 2 * m + d) + 2 * m);
        term_out += (long long)ng.alive_per_shape[m] * 8LL * (m * d + m + d);
      }
      const long long tab_in = 20LL * stats.cells_parents, tab_out = 20LL * stats.cells_survivors;
      stats.bytes_step = term_in + tab_in + 2 * payload + term_out + tab_out;
      stats.bytes_gtable = tab_in + payload + term_out + tab_out;     // what the group kernel itself must move
    }
    return 0;
  }

  struct KRootFlags {       // flags[j] = (F[j] == j) for j < n
    int n; const int* F; int* flags;
    template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
      c.par([&](int tid) { const int j = c.block() * c.nthreads() + tid; if (j < n) flags[j] = (F[j] == j) ? 1 : 0; });
    }
  };

  void capture_shape(const TermView& tv, int m, const int* F, const ParentWs& ws, bool with_tp, GenStore& pg) {
    CapShape cs; cs.m = m; cs.n = tv.n[m];
    const int n = cs.n;
    cs.A.resize((size_t)n * m * d); cs.p.resize((size_t)n * m); cs.q.resize((size_t)n * m); cs.b.resize((size_t)n * d); cs.cd.resize((size_t)n * 2);
    cs.meta.resize((size_t)n * 8); cs.F.resize(n); cs.cmap.resize((size_t)n * MAXM); cs.csmap.resize((size_t)n * MAXM);
    std::vector<SlotMeta> me(n);
    be.d2h(cs.A.data(), term_A(tv, m, 0, d), sizeof(double) * cs.A.size());
    be.d2h(cs.p.data(), term_p(tv, m, 0), sizeof(double) * cs.p.size());
    be.d2h(cs.q.data(), term_q(tv, m, 0), sizeof(double) * cs.q.size());
    be.d2h(cs.b.data(), term_b(tv, m, 0, d), sizeof(double) * cs.b.size());
    be.d2h(me.data(), tv.meta + tv.t_begin[m], sizeof(SlotMeta) * n);
    be.d2h(cs.cmap.data(), tv.cmap + tv.t_begin[m] * MAXM, (size_t)n * MAXM);
    be.d2h(cs.F.data(), F, sizeof(int) * n);
    std::vector<int> alive(pg.v.n_alive), cells(pg.v.n_groups); std::vector<unsigned char> gm(pg.v.n_groups);
    be.d2h(alive.data(), pg.v.alive, sizeof(int) * alive.size());
    be.d2h(cells.data(), pg.v.cells, sizeof(int) * cells.size());
    be.d2h(gm.data(), pg.v.g_m, gm.size());
    std::vector<int> tpc(pg.v.n_alive, 0);
    if (with_tp) be.d2h(tpc.data(), ws.tpB_cells, sizeof(int) * tpc.size());
    for (int i = 0; i < n; i++) {
      const SlotMeta& s = me[i];
      cs.cd[2 * i] = s.c_val; cs.cd[2 * i + 1] = s.d_val;
      int* o = &cs.meta[(size_t)i * 8];
      o[0] = gm[alive[s.parent]]; o[1] = s.pbc; o[2] = s.z; o[3] = (int)s.enc_lhp; o[4] = (int)s.hflag; o[5] = s.flags & 1; o[6] = with_tp ? tpc[s.parent] : cells[alive[s.parent]]; o[7] = (s.flags >> 1) & 1;
      for (int l = 0; l < MAXM; l++) {
        const bool has = ((s.flags >> 1) & 1) && l < s.pbc;
        cs.csmap[(size_t)i * MAXM + l] = has ? (((s.csneg >> l) & 1u) ? -1 : 1) : 0;
        if (!has) cs.cmap[(size_t)i * MAXM + l] = 255;
      }
    }
    cap.push_back(std::move(cs));
  }

  // compute_moments(false), est:524-602: moments of the surviving terms from their new tables; no numeric check.
  void post_ftr_moments(const StepParams& sp) {
    GenStore& g = gen[cur];
    const int n = g.v.n_alive, nq = 1 + d + d * d;
    if (n == 0) return;
    cplx* gg = (cplx*)slg.ensure(sizeof(cplx) * (size_t)(n + 8));
    double* yy = (double*)sly.ensure(sizeof(double) * (size_t)(n + 1) * 2 * d);
    double* mom = (double*)momOut.ensure(sizeof(double) * 4 * nq + 16);
    be.launch(KPostFtrMoments{sp, g.v, gg, yy}, (n + 127) / 128, 128, 0);
    be.launch(KMomentsSerial{gg, yy, (long long)n, d, mom}, (nq + MOM_QB - 1) / MOM_QB, 512, KMomentsSerial::smem_bytes(d));
    std::vector<double> raw(2 * nq);
    be.d2h(raw.data(), mom, sizeof(double) * 2 * nq);
    const cplx keep = fz_mu; const std::vector<cplx> km = mean_mu, kv = var_mu;
    finalize_moments(raw.data(), false);
    fz_mu = keep; mean_mu = km; var_mu = kv;
  }

  // ------------------------------------------------------------------------------------------
  int shift_b(const double* delta, double sign) {
    if (skip_post_mu) return 0;                 // est:1314, 1365
    if (master_step == 0) {                     // before the first step the CF is the initial term (terms_dp[d][0], est:1316-1326)
      for (int j = 0; j < d; j++) { if (sign < 0) b1[j] -= delta[j]; else b1[j] += delta[j]; }
      return 0;
    }
    GenStore& g = gen[cur];
    if (g.v.n_alive == 0) return 0;
    KShiftB k; k.gen = g.v; k.d = d; k.sign = sign;
    for (int i = 0; i < d; i++) k.delta[i] = delta[i];
    be.launch(k, (g.v.n_alive + 127) / 128, 128, 0);
    return 0;
  }
  int det_time_prop(const double* T, const double* B, const double* u) {
    GenStore& g = gen[cur];
    KDetTimeProp k; memset(&k, 0, sizeof(k)); k.gen = g.v; k.d = d;
    for (int i = 0; i < d * d; i++) k.T[i] = T[i];
    if (B && u && cmcc > 0) { k.has_bu = 1; for (int i = 0; i < d; i++) { double s = 0.0; for (int j = 0; j < cmcc; j++) s += B[i * cmcc + j] * u[j]; k.bu[i] = s; } }
    if (master_step == 0) {                     // the initial term is propagated like any other (est:1346-1354 walks terms_dp[d][0]); d x d host work, same operation order as KDetTimeProp
      double work[MAXD];
      for (int i = 0; i < d; i++) {
        for (int kk = 0; kk < d; kk++) work[kk] = A1[i * d + kk];
        for (int j = 0; j < d; j++) { double sum = 0.0; for (int kk = 0; kk < d; kk++) sum += work[kk] * T[kk + j * d]; A1[i * d + j] = sum; }
      }
      for (int kk = 0; kk < d; kk++) work[kk] = b1[kk];
      for (int i = 0; i < d; i++) { double sum = 0.0; for (int j = 0; j < d; j++) sum += T[i * d + j] * work[j]; b1[i] = sum; }
      if (k.has_bu) for (int i = 0; i < d; i++) b1[i] += 1.0 * k.bu[i];
      return 0;
    }
    if (g.v.n_alive == 0) return 0;
    be.launch(k, (g.v.n_alive + 127) / 128, 128, 0);
    return 0;
  }
  void reset() {                                 // est:1247-1300
    master_step = 0; Nt = 1; numeric_moment_errors = 0; finished = false; skip_post_mu = 0;
    std::fill(terms_per_shape.begin(), terms_per_shape.end(), 0); terms_per_shape[d] = 1;
    gen[0].v.n_alive = 0; gen[1].v.n_alive = 0;
    std::fill(g_alive_per_shape.begin(), g_alive_per_shape.end(), 0);
    A1 = A0; p1 = p0; b1 = b0;                   // setup_first_term(A0_init, p0_init, b0_init), est:1280
    cur = 0;          // same buffer parity on every pass of a window: the grow-only buffers settle after the first pass
  }


  // Point-wise 1-D marginal cpdf at the points xs[0..n) (cpdf_ndim.hpp:1233-1354 as driven by the grid dispatcher,
  // cpdf_ndim.hpp:2074-2139: the first point is evaluated uncached, the others from the per-term cache).  Returns n, 0 when
  // the estimator holds no tables (the window's last step, SKIP_LAST_STEP: cpdf_ndim.hpp:2079-2083), < 0 on misuse.
  int marginal_1d_points(int marg_idx, const double* bar_nu, int n, const double* xs, double* ys) {
    if (master_step < 1) { error = "marginal cpdf: the estimator has not been stepped (cpdf_ndim.hpp:1239)"; return -2; }
    if (marg_idx < 0 || marg_idx >= d || n < 1) { error = "marginal cpdf: bad state index or point count"; return -2; }
    if (master_step == num_estimation_steps || skip_post_mu) return 0;
    GenStore& g = gen[cur];
    const int nt = g.v.n_alive;
    if (nt <= 0) return 0;
    double* val0 = (double*)cpVal0.ensure(sizeof(double) * (size_t)nt);
    Cpdf1dTerm* cache = (Cpdf1dTerm*)cpCache.ensure(sizeof(Cpdf1dTerm) * (size_t)nt);
    double* dxs = (double*)cpXs.ensure(sizeof(double) * (size_t)n);
    double* dout = (double*)cpOut.ensure(sizeof(double) * (size_t)n);
    be.h2d(dxs, xs, sizeof(double) * (size_t)n);
    be.ev_record(10);
    KCpdf1dTerms kt; memset(&kt, 0, sizeof(kt));
    kt.gen = g.v; kt.d = d; kt.marg_idx = marg_idx; kt.x0 = xs[0]; kt.val0 = val0; kt.cache = cache;
    for (int i = 0; i < d; i++) kt.bar_nu[i] = bar_nu[i];
    be.launch(kt, (nt + 127) / 128, 128, 0);
    const int NTH = 32;      // one warp per CTA: a few thousand grid points still reach every SM
    KCpdf1dGrid kg{nt, n, dxs, val0, cache, dout};
    be.launch(kg, (n + NTH - 1) / NTH, NTH, KCpdf1dGrid::smem_bytes());
    be.ev_record(11);
    be.d2h(ys, dout, sizeof(double) * (size_t)n);
    cpdf_ms = be.ev_elapsed(10, 11);
    const double norm_factor = fz.re, RECIPRICAL_TWO_PI = 1.0 / (2.0 * M_PI);   // cauchy_constants.hpp:25
    ys[0] = ys[0] * RECIPRICAL_TWO_PI / norm_factor;                           // cpdf_ndim.hpp:1351-1352
    for (int k = 1; k < n; k++) ys[k] = ys[k] / norm_factor;                   // cpdf_ndim.hpp:1349-1350
    return n;
  }



  // Point-wise 2-D marginal cpdf of states (idx1 < idx2) at the points (xs[k], ys[k]) (cpdf_ndim.hpp:1356-1455 as driven by
  // CauchyCPDFGridDispatcher2D, :1850-1919).  Returns n, 0 when no tables exist, < 0 on misuse, -5 on the reference's
  // "possible singularity" exit (cpdf_ndim.hpp:1624-1628).
  int marginal_2d_points(int idx1, int idx2, const double* bar_nu, int n, const double* xs, const double* ys, double* zs) {
    if (master_step < 1) { error = "marginal cpdf: the estimator has not been stepped (cpdf_ndim.hpp:1359)"; return -2; }
    if (idx1 < 0 || idx1 >= idx2 || idx2 >= d || n < 1) { error = "marginal cpdf: state indices must satisfy 0 <= idx1 < idx2 < d (cpdf_ndim.hpp:1852-1855)"; return -2; }
    if (master_step == num_estimation_steps || skip_post_mu) return 0;
    GenStore& g = gen[cur];
    const int nt = g.v.n_alive;
    if (nt <= 0) return 0;
    const int S = max_shape, R = cpdf2_rec_doubles(S);
    double* recs = (double*)cpRecs.ensure(sizeof(double) * (size_t)R * nt);
    double* dxs = (double*)cpXs.ensure(sizeof(double) * (size_t)n);
    double* dys = (double*)cpYs.ensure(sizeof(double) * (size_t)n);
    double* dout = (double*)cpOut.ensure(sizeof(double) * (size_t)n);
    int* bad = (int*)cpBad.ensure(sizeof(int) * 4);
    // terms per chunk: the value matrix V[chunk][n] stays below ~256 MB and a chunk is a whole number of term tiles
    long long chunk = (256ll << 20) / (8ll * n);
    chunk = chunk < CPDF2_TERMS ? CPDF2_TERMS : (chunk / CPDF2_TERMS) * CPDF2_TERMS;
    if (chunk > nt) chunk = ((nt + CPDF2_TERMS - 1) / CPDF2_TERMS) * CPDF2_TERMS;
    double* V = (double*)cpV.ensure(sizeof(double) * (size_t)chunk * n);
    be.h2d(dxs, xs, sizeof(double) * (size_t)n);
    be.h2d(dys, ys, sizeof(double) * (size_t)n);
    be.memset(bad, 0, sizeof(int) * 4);
    be.ev_record(10);
    KCpdf2dTerms kt; memset(&kt, 0, sizeof(kt));
    kt.gen = g.v; kt.d = d; kt.idx1 = idx1; kt.idx2 = idx2; kt.S = S; kt.recs = recs;
    for (int i = 0; i < d; i++) kt.bar_nu[i] = bar_nu[i];
    be.launch(kt, (nt + 63) / 64, 64, 0);
    const int n_ptiles = (n + CPDF2_PTS - 1) / CPDF2_PTS;
    for (long long t0 = 0; t0 < nt; t0 += chunk) {
      const int cnt = (int)(nt - t0 < chunk ? nt - t0 : chunk), n_ttiles = (cnt + CPDF2_TERMS - 1) / CPDF2_TERMS;
      KCpdf2dValues kv{S, (int)t0, cnt, n, n_ptiles, dxs, dys, recs, V, bad};
      be.launch(kv, n_ptiles * n_ttiles, CPDF2_PTS, KCpdf2dValues::smem_bytes(S));
      KCpdf2dSum ks{cnt, n, t0 == 0 ? 1 : 0, V, dout};
      be.launch(ks, (n + 31) / 32, 32, 0);
    }
    be.ev_record(11);
    int hbad = 0;
    be.d2h(zs, dout, sizeof(double) * (size_t)n);
    be.d2h(&hbad, bad, sizeof(int));
    cpdf_ms = be.ev_elapsed(10, 11);
    if (hbad) { error = "marginal cpdf: possible singularity, gamma1 and gamma2 both vanish at a grid point (cpdf_ndim.hpp:1624-1628)"; return -5; }
    const double norm_factor = fz.re, RECIPRICAL_TWO_PI = 1.0 / (2.0 * M_PI);
    for (int k = 0; k < n; k++) zs[k] = 2 * zs[k] * RECIPRICAL_TWO_PI * RECIPRICAL_TWO_PI / norm_factor;   // cpdf_ndim.hpp:1446
    return n;
  }

  // Test hook: the serial-order moment sums (KMomentsSerial) of caller-supplied per-slot values g[n] (complex) and y[n][dd] (complex);
  // out[2 * (1 + dd + dd * dd)] receives the sums.
  int debug_moment_sums(long long n, int dd, const double* g, const double* y, double* out) {
    const int nq = 1 + dd + dd * dd;
    cplx* gg = (cplx*)slg.ensure(sizeof(cplx) * (size_t)(n + 8));
    double* yy = (double*)sly.ensure(sizeof(double) * ((size_t)(n + 1) * 2 * (dd > 0 ? dd : 1)));
    double* mom = (double*)momOut.ensure(sizeof(double) * 4 * nq + 16);
    if (n > 0) be.h2d(gg, g, sizeof(cplx) * (size_t)n);
    if (n > 0 && dd > 0) be.h2d(yy, y, sizeof(double) * (size_t)n * 2 * dd);
    be.launch(KMomentsSerial{gg, yy, n, dd, mom}, nq, 512, KMomentsSerial::smem_bytes(dd));
    be.d2h(out, mom, sizeof(double) * 2 * nq);
    return 0;
  }

  // Test hook: the per-slot moment inputs of the LAST step (g[n] complex, y[n][d] complex: what the moment kernels add up), valid until the next step.
  long long last_nslots = 0;
  long long debug_export_slots(long long cap, double* g, double* y) {
    const long long n = last_nslots < cap ? last_nslots : cap;
    if (n > 0 && g) be.d2h(g, slg.p, sizeof(cplx) * (size_t)n);
    if (n > 0 && y) be.d2h(y, sly.p, sizeof(double) * (size_t)n * 2 * d);
    return last_nslots;
  }

  // Test hook: KSumScan over the real parts of n complex values; out[0] = the serial-order sum, out[1] = restarts of the scan.
  int debug_sum_scan(long long n, const double* g, double* out) {
    cplx* gg = (cplx*)slg.ensure(sizeof(cplx) * (size_t)(n + 8));
    double* mom = (double*)momOut.ensure(sizeof(double) * 4 * (1 + d + d * d) + 16);
    if (n > 0) be.h2d(gg, g, sizeof(cplx) * (size_t)n);
    be.ev_record(10);
    launch_sum_scan(false, (const double*)gg, 2, n, mom);
    be.ev_record(11);
    be.d2h(out, mom, sizeof(double) * 3);
    cpdf_ms = be.ev_elapsed(10, 11);          // device time of the kernel, read back through mce_cpdf_last_ms
    return 0;
  }

  // device self-test of div_nobranch (mce_math.h): out[0] = flagged-ok pairs that differ from a / b, out[1] = ok pairs
  int div_selftest(long long n, unsigned long long seed, unsigned long long* out) {
    unsigned long long* cnt = (unsigned long long*)cpOut.ensure(sizeof(unsigned long long) * 2);
    be.memset(cnt, 0, sizeof(unsigned long long) * 2);
    be.launch(KDivSelfTest{seed, n, cnt}, 148 * 8, 128, 0);
    be.d2h(out, cnt, sizeof(unsigned long long) * 2);
    return 0;
  }

  // ---- term-list export: gather kernels + one bulk copy per array ----
  struct KExportCells {     // cells[i] of the i-th surviving term of a shape
    GenView gen; int rank0, n; int* cells;
    template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
      c.par([&](int tid) { const int i = c.block() * c.nthreads() + tid; if (i < n) cells[i] = gen.cells[gen.alive[rank0 + i]]; });
    }
  };
  struct KExportGather {    // one CTA per term: hyperplanes, weights, offset and the table (keys + values) into contiguous staging arrays
    GenView gen; int rank0, m, d; const int* off; double *A, *p, *b; unsigned* keys; cplx* G;
    template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
      const int i = c.block(), gid = gen.alive[rank0 + i], nc = gen.cells[gid]; const long long o = off[i];
      const double* As = gen_A(gen, gid, m, d); const double* ps = gen_p(gen, gid, m); const double* bs = gen_b(gen, gid, d);
      const unsigned* ks = gen_keys(gen, gid, m); const cplx* Gs = gen_G(gen, gid, m);
      c.par([&](int tid) {
        for (int k = tid; k < m * d; k += c.nthreads()) A[(long long)i * m * d + k] = As[k];
        for (int k = tid; k < m; k += c.nthreads()) p[(long long)i * m + k] = ps[k];
        for (int k = tid; k < d; k += c.nthreads()) b[(long long)i * d + k] = bs[k];
        for (int k = tid; k < nc; k += c.nthreads()) { keys[o + k] = ks[k]; G[o + k] = Gs[k]; }
      });
    }
  };
  // Host copy of the parents of shape m (canonical order): two kernels and six device-to-host copies per shape.
  int export_shape(int m, int* n_terms, long long* n_cells_total, double* A, double* pp, double* b, int* cells, uint32_t* keys, double* G) {
    GenStore& g = gen[cur];
    if (m < 1 || m >= NSHAPE || master_step == 0) { *n_terms = 0; *n_cells_total = 0; return 0; }
    const int n = g.alive_per_shape[m];
    *n_terms = n;
    if (n == 0) { *n_cells_total = 0; return 0; }
    int rank0 = 0; for (int k = 1; k < m; k++) rank0 += g.alive_per_shape[k];
    int* dcells = (int*)scratchI0.ensure(sizeof(int) * (size_t)(n + 4));
    int* doff = (int*)scratchI1.ensure(sizeof(int) * (size_t)(n + 4));
    be.launch(KExportCells{g.v, rank0, n, dcells}, (n + 127) / 128, 128, 0);
    std::vector<int> hc(n);
    be.d2h(hc.data(), dcells, sizeof(int) * n);
    long long tot = 0;
    for (int i = 0; i < n; i++) tot += hc[i];
    *n_cells_total = tot;
    if (!A) return 0;
    if (tot > 0x7fffffffLL) { error = "export_shape: more than 2^31 table cells in one shape"; return -5; }
    be.exclusive_scan(dcells, doff, n);
    const size_t nA = (size_t)n * m * d, np = (size_t)n * m, nb = (size_t)n * d;
    cplx* sG = (cplx*)cpV.ensure(sizeof(double) * (nA + np + nb + 2 * (size_t)tot + 8) + sizeof(unsigned) * ((size_t)tot + 8));   // 16-byte values first
    double* sA = (double*)(sG + tot); double* sp_ = sA + nA; double* sb = sp_ + np; unsigned* sk = (unsigned*)(sb + nb);
    be.launch(KExportGather{g.v, rank0, m, d, doff, sA, sp_, sb, sk, sG}, n, 128, 0);
    be.d2h(A, sA, sizeof(double) * nA); be.d2h(pp, sp_, sizeof(double) * np); be.d2h(b, sb, sizeof(double) * nb);
    if (tot > 0) { be.d2h(keys, sk, sizeof(unsigned) * (size_t)tot); be.d2h(G, sG, sizeof(cplx) * (size_t)tot); }
    memcpy(cells, hc.data(), sizeof(int) * n);
    return 0;
  }
};

}  // namespace mce
#endif
