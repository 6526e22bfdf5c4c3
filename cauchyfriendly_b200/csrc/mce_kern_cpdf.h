// mce_kern_cpdf.h -- point-wise 1-D marginal conditional pdf on a grid, evaluated on the device from the resident
// term list.  Replaces, for the grid use case, PointWiseNDimCauchyCPDF::
// evaluate_1D_marginal_cpdf (cpdf_ndim.hpp:1233-1354) driven by CauchyCPDFGridDispatcher1D::evaluate_point_grid
// (cpdf_ndim.hpp:2074-2139):
//   * KCpdf1dTerms, thread / term: the uncached evaluation of the FIRST grid point (two G-table lookups, two complex
//     divisions in libgcc's algorithm) and the per-term cache {a, b, w^2, s} of cpdf_ndim.hpp:216-222, 1329-1339;
//   * KCpdf1dGrid, thread / grid point: every point walks ALL terms in canonical order and adds
//     (a x + b) / (w^2 + (x - s)^2) to its own running sum -- the reference's summation order, so the values are
//     bit-identical; the cached terms are staged through shared memory in tiles and broadcast to the threads.
// Both kernels read the generation store in place (no term export); nothing here touches the host.
#ifndef MCE_KERN_CPDF_H_
#define MCE_KERN_CPDF_H_

#include "mce_kern_prop.h"

namespace mce {

struct Cpdf1dTerm { double a, b, w2, s; };     // Cached1DCPDFTerm with w squared once (w*w is what every point computes)

// binary search of a key in a sorted table (gtable.hpp:283-301), half storage + conjugate (eval_gs.hpp:94-153)
MCE_HD cplx cpdf_lookup(const unsigned* keys, const cplx* G, int cells, int enc, int top, int rev) {
  const bool cj = (enc & top) != 0;
  const unsigned target = (unsigned)(cj ? (rev ^ enc) : enc);
  int lo = 0, hi = cells - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) / 2;
    const unsigned mk = keys[mid];
    if (mk == target) { const cplx v = G[mid]; return cj ? cconj(v) : v; }
    if (mk > target) hi = mid - 1; else lo = mid + 1;
  }
  return make_cplx(0, 0);
}

struct KCpdf1dTerms {
  GenView gen; int d, marg_idx; double x0; double bar_nu[MAXD];
  double* val0;          // [n_alive] creal(g_val) of the first grid point, term order
  Cpdf1dTerm* cache;     // [n_alive]
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int r = c.block() * c.nthreads() + tid;
      if (r >= gen.n_alive) return;
      const int gid = gen.alive[r], m = gen_m(gen, gid), cells = gen.cells[gid];
      const double* A = gen_A(gen, gid, m, d); const double* p = gen_p(gen, gid, m); const double* b = gen_b(gen, gid, d);
      const double b_c = b[marg_idx] - x0;
      double p_cc = 0; int lhs = 0, rhs = 0;
      for (int j = 0; j < m; j++) {                                       // cpdf_ndim.hpp:1300-1322
        const double A_cj = A[j * d + marg_idx], f = fabs(A_cj);
        p_cc += p[j] * f;
        bool left;
        if (f > 1e-15) left = A_cj > 0;
        else left = dot_lr(A + j * d, bar_nu, d) > 0;
        if (left) lhs |= 1 << j; else rhs |= 1 << j;
      }
      const unsigned* keys = gen_keys(gen, gid, m); const cplx* G = gen_G(gen, gid, m);
      const int top = 1 << (m - 1), rev = (1 << m) - 1;
      const cplx gl = cpdf_lookup(keys, G, cells, lhs, top, rev), gr = cpdf_lookup(keys, G, cells, rhs, top, rev);
      const cplx gv = csub(cdiv(gl, make_cplx(p_cc, b_c)), cdiv(gr, make_cplx(-p_cc, b_c)));   // cpdf_ndim.hpp:1326
      val0[r] = gv.re;
      Cpdf1dTerm t;
      t.s = b[marg_idx];
      t.a = gr.im / M_PI;
      t.b = (gr.re * p_cc - gr.im * t.s) / M_PI;
      t.w2 = p_cc * p_cc;
      cache[r] = t;
    });
  }
};

constexpr int CPDF_TILE = 256;     // cached terms staged per barrier
constexpr int CPDF_UNROLL = 8;     // independent quotients in flight per thread

// one cached term at x (cpdf_ndim.hpp:1272-1279)
MCE_HD double cpdf1d_term(const Cpdf1dTerm& t, double x1) {
  double x1ms = x1 - t.s;
  x1ms *= x1ms;
  return (t.a * x1 + t.b) / (t.w2 + x1ms);
}

struct KCpdf1dGrid {
  static constexpr int kMaxThreads = 64, kMinBlocks = 1;     // one or two warps per CTA: registers are not the limit here
  int n_terms, n_pts; const double* xs; const double* val0; const Cpdf1dTerm* cache; double* out;   // out[k] = unnormalised sum of point k
  static MCE_HD size_t smem_bytes() { return sizeof(Cpdf1dTerm) * CPDF_TILE + sizeof(double) * CPDF_TILE; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    Cpdf1dTerm* st = (Cpdf1dTerm*)c.smem();
    double* sv = (double*)(st + CPDF_TILE);
    const int NT = c.nthreads();
    // running sums live in `out` between tiles (a thread only ever touches its own point)
    c.par([&](int tid) { const int k = c.block() * NT + tid; if (k < n_pts) out[k] = 0.0; });
    for (int base = 0; base < n_terms; base += CPDF_TILE) {
      const int nt = n_terms - base < CPDF_TILE ? n_terms - base : CPDF_TILE;
      c.par([&](int tid) { for (int i = tid; i < nt; i += NT) { st[i] = cache[base + i]; sv[i] = val0[base + i]; } });
      c.par([&](int tid) {
        const int k = c.block() * NT + tid;
        if (k >= n_pts) return;
        double fx = out[k];
        if (k == 0) {                                  // first point: the uncached per-term values (cpdf_ndim.hpp:1327)
          for (int i = 0; i < nt; i++) fx += sv[i];
        } else {                                       // cached points (cpdf_ndim.hpp:1269-1280)
          // The quotients of a batch are independent (eight division chains in flight per thread); only the additions
          // to fx are ordered -- they run in term order, one after the other, exactly like the reference's loop.
          const double x1 = xs[k];
          int i = 0;
          for (; i + CPDF_UNROLL <= nt; i += CPDF_UNROLL) {
            double q[CPDF_UNROLL]; bool all_ok = true;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int u = 0; u < CPDF_UNROLL; u++) {
              const Cpdf1dTerm t = st[i + u];
              double x1ms = x1 - t.s;
              x1ms *= x1ms;
              bool ok;
              q[u] = div_nobranch(t.a * x1 + t.b, t.w2 + x1ms, &ok);      // == cpdf1d_term when ok
              all_ok = all_ok && ok;
            }
            if (!all_ok) {                      // rare (zero / tiny numerators): redo the batch with the plain division
              for (int u = 0; u < CPDF_UNROLL; u++) q[u] = cpdf1d_term(st[i + u], x1);
            }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int u = 0; u < CPDF_UNROLL; u++) fx += q[u];
          }
          for (; i < nt; i++) fx += cpdf1d_term(st[i], x1);
        }
        out[k] = fx;
      });
    }
  }
};


// ---------------------------------------------------------------------------------------------------------------------
// Point-wise 2-D marginal cpdf (cpdf_ndim.hpp:1356-1455).  Three kernels:
//   * KCpdf2dTerms, thread / term: the per-term cache of marg2d_eval_term_for_cpdf (cpdf_ndim.hpp:1474-1658) -- the two
//     marginal columns (:31-40), zero rows removed and (anti)parallel rows merged (:42-139), cell-wall angles (:142-170),
//     one sign vector per cell (:176-201), the G value of every cell, the real parts of gamma_1 / gamma_2 and the sines and
//     cosines of the wall angles -- as one fixed-stride record per term;
//   * KCpdf2dValues, CTA = 64 grid points x 64 terms: the term's integral at the point, marg2d_cached_eval_term_for_cpdf
//     (:1660-1747), for every (term, point) pair of a chunk of terms, records broadcast from shared memory;
//   * KCpdf2dSum, thread / grid point: adds the chunk's values to the point's running sum in term order.
// Splitting evaluation from summation keeps the reference's summation order (the sums cancel heavily) while every SM
// works on the expensive part.  atan2 / sin / cos are CUDA's (<= 2 ulp), the reference uses glibc's: values agree to
// rounding noise, not bit for bit -- the tests use a relative tolerance (tests/test_gpu_cpdf.py).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int CPDF2_ZERO_HP = 32;            // ZERO_HP_MARKER_VALUE, cpdf_ndim.hpp:373
constexpr double COALIGN_MU_EPS = COALIGN_EPS;        // cauchy_constants.hpp:32
constexpr double INTEGRAL_GAMMA_EPS = 1e-8;          // cauchy_constants.hpp:76
MCE_HD int cpdf2_rec_doubles(int S) { return 4 + 2 * (S + 1) + 4 * S; }
// record: [0] m, [1] b0, [2] b1, [3] -, then (sin, cos)[S+1], then (gam1_real, gam2_real, g.re, g.im)[S]

struct KCpdf2dTerms {
  GenView gen; int d, idx1, idx2, S; double bar_nu[MAXD]; double* recs;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int r = c.block() * c.nthreads() + tid;
      if (r >= gen.n_alive) return;
      const int gid = gen.alive[r], m = gen_m(gen, gid), cells = gen.cells[gid];
      const double* A = gen_A(gen, gid, m, d); const double* p = gen_p(gen, gid, m); const double* b = gen_b(gen, gid, d);
      double wA[2 * MAXM], wA2[2 * MAXM], wp[MAXM], wp2[MAXM], thetas[2 * MAXM];
      int c_map[MAXM], cs_map[MAXM], F_idxs[MAXM]; bool F[MAXM]; unsigned svneg[MAXM];
      for (int i = 0; i < m; i++) { wp[i] = p[i]; wA[2 * i] = A[i * d + idx1]; wA[2 * i + 1] = A[i * d + idx2]; F[i] = true; }
      for (int i = 0; i < m; i++) {                                              // cpdf_ndim.hpp:57-80
        double* ai = wA + 2 * i; const double f0 = fabs(ai[0]), f1 = fabs(ai[1]);
        if (f0 < MU_EPS && f1 < MU_EPS) { c_map[i] = CPDF2_ZERO_HP; cs_map[i] = CPDF2_ZERO_HP; F_idxs[i] = CPDF2_ZERO_HP; F[i] = false; continue; }
        const double sum_a = f0 + f1; ai[0] /= sum_a; ai[1] /= sum_a; wp[i] *= sum_a;
      }
      int mn = 0;
      for (int i = 0; i < m; i++) {                                              // cpdf_ndim.hpp:83-122
        if (!F[i]) continue;
        c_map[i] = mn; cs_map[i] = 1; F_idxs[i] = i; wp2[mn] = wp[i];
        const double* ai = wA + 2 * i;
        for (int j = i + 1; j < m; j++) {
          if (!F[j]) continue;
          const double* aj = wA + 2 * j;
          if (fabs(ai[0] - aj[0]) < COALIGN_MU_EPS && fabs(ai[1] - aj[1]) < COALIGN_MU_EPS) { c_map[j] = mn; cs_map[j] = 1; F[j] = false; F_idxs[j] = i; wp2[mn] += wp[j]; continue; }
          if (fabs(ai[0] + aj[0]) < COALIGN_MU_EPS && fabs(ai[1] + aj[1]) < COALIGN_MU_EPS) { c_map[j] = mn; cs_map[j] = -1; F[j] = false; F_idxs[j] = i; wp2[mn] += wp[j]; continue; }
        }
        mn++;
      }
      if (mn < m) { mn = 0; for (int i = 0; i < m; i++) if (F_idxs[i] == i) { wA2[2 * mn] = wA[2 * i]; wA2[2 * mn + 1] = wA[2 * i + 1]; mn++; } }
      else for (int i = 0; i < 2 * m; i++) wA2[i] = wA[i];
      const bool use_maps = mn != m;
      for (int i = 0, k = 0; i < mn; i++) {                                      // cell-wall angles, cpdf_ndim.hpp:142-170
        const double* a = wA2 + 2 * i; double p0, p1;
        if (fabs(a[0]) < fabs(a[1])) { p0 = 1; p1 = -a[0] / a[1]; } else { p0 = -a[1] / a[0]; p1 = 1; }
        double t1 = atan2(p1, p0);
        if (t1 < 0) t1 += M_PI;
        thetas[k++] = t1; thetas[k++] = t1 + M_PI;
      }
      for (int i = 1; i < 2 * mn; i++) { const double v = thetas[i]; int j = i - 1; while (j >= 0 && thetas[j] > v) { thetas[j + 1] = thetas[j]; j--; } thetas[j + 1] = v; }
      for (int i = 0; i < mn; i++) {                                             // sign vector of cell i, cpdf_ndim.hpp:176-190
        const double tt = (thetas[i + 1] + thetas[i]) / 2.0, p0 = cos(tt), p1 = sin(tt);
        unsigned neg = 0;
        for (int j = 0; j < mn; j++) { const double sum = wA2[2 * j] * p0 + wA2[2 * j + 1] * p1; if (!(sum > 0)) neg |= 1u << j; }
        svneg[i] = neg;
      }
      double* rec = recs + (long long)r * cpdf2_rec_doubles(S);
      rec[0] = (double)mn; rec[1] = b[idx1]; rec[2] = b[idx2]; rec[3] = 0;
      double* sc = rec + 4; double* gq = rec + 4 + 2 * (S + 1);
      const unsigned* keys = gen_keys(gen, gid, m); const cplx* G = gen_G(gen, gid, m);
      const int top = 1 << (m - 1), rev = (1 << m) - 1;
      for (int i = 0; i < mn; i++) {                                             // cpdf_ndim.hpp:1548-1600
        const unsigned neg = svneg[i]; int enc = 0;
        if (!use_maps) enc = (int)neg;
        else for (int j = 0; j < m; j++) {
          if (c_map[j] == CPDF2_ZERO_HP) { if (dot_lr(A + j * d, bar_nu, d) < 0) enc |= 1 << j; }
          else { const int sv = ((neg >> c_map[j]) & 1u) ? -1 : 1; if (sv * cs_map[j] < 0) enc |= 1 << j; }
        }
        const cplx gv = cpdf_lookup(keys, G, cells, enc, top, rev);
        double g1 = 0, g2 = 0;
        for (int j = 0; j < mn; j++) {
          const double sv = ((neg >> j) & 1u) ? -1.0 : 1.0;
          g1 -= (wA2[2 * j] * wp2[j]) * sv; g2 -= (wA2[2 * j + 1] * wp2[j]) * sv;
        }
        sc[2 * i] = sin(thetas[i]); sc[2 * i + 1] = cos(thetas[i]);
        gq[4 * i] = g1; gq[4 * i + 1] = g2; gq[4 * i + 2] = gv.re; gq[4 * i + 3] = gv.im;
      }
      sc[2 * mn] = sin(thetas[mn]); sc[2 * mn + 1] = cos(thetas[mn]);
    });
  }
};

// complex * real and real / complex as GNU C evaluates them: componentwise product; the real numerator promoted to x + 0i
MCE_HD cplx cpdf2_lim_fast(double s, double cth, cplx g1sq, cplx g12) { return cdiv(make_cplx(s, 0.0), cadd(cscale(g1sq, cth), cscale(g12, s))); }

// marg2d_cached_eval_term_for_cpdf, cpdf_ndim.hpp:1660-1747; *bad is set on the reference's "possible singularity" exit
MCE_HD double cpdf2_term(const double* rec, int S, double x1, double x2, int* bad) {
  const int m = (int)rec[0];
  const double gam1_imag = rec[1] - x1, gam2_imag = rec[2] - x2;
  const double* sc = rec + 4; const double* gq = rec + 4 + 2 * (S + 1);
  const bool check_gamma1 = fabs(gam1_imag) < INTEGRAL_GAMMA_EPS;
  double term_integral = 0;
  for (int i = 0; i < m; i++) {
    const double sin_t1 = sc[2 * i], cos_t1 = sc[2 * i + 1], sin_t2 = sc[2 * i + 2], cos_t2 = sc[2 * i + 3];
    cplx gamma1 = make_cplx(gq[4 * i], gam1_imag), gamma2 = make_cplx(gq[4 * i + 1], gam2_imag);
    const cplx gv = make_cplx(gq[4 * i + 2], gq[4 * i + 3]);
    bool fast = true;
    if (check_gamma1 && fabs(gq[4 * i]) < INTEGRAL_GAMMA_EPS) {
      fast = false;
      if (fabs(gq[4 * i + 1]) < INTEGRAL_GAMMA_EPS && fabs(gam2_imag) < INTEGRAL_GAMMA_EPS) *bad = 1;
    }
    cplx cell;
    if (fast) {
      gamma2 = cmul(gamma2, gamma1); gamma1 = cmul(gamma1, gamma1);
      const cplx lo = cpdf2_lim_fast(sin_t1, cos_t1, gamma1, gamma2), hi = cpdf2_lim_fast(sin_t2, cos_t2, gamma1, gamma2);
      cell = cmul(csub(hi, lo), gv);
    } else {
      const cplx lo = cdiv(csub(cscale(gamma1, sin_t1), cscale(gamma2, cos_t1)), cadd(cscale(gamma1, cos_t1), cscale(gamma2, sin_t1)));
      const cplx hi = cdiv(csub(cscale(gamma1, sin_t2), cscale(gamma2, cos_t2)), cadd(cscale(gamma1, cos_t2), cscale(gamma2, sin_t2)));
      gamma1 = cmul(gamma1, gamma1); gamma2 = cmul(gamma2, gamma2);
      cell = cmul(csub(hi, lo), cdiv(gv, cadd(gamma1, gamma2)));
    }
    term_integral += cell.re;
  }
  return term_integral;
}

constexpr int CPDF2_PTS = 64, CPDF2_TERMS = 64;
struct KCpdf2dValues {          // grid = (point tiles, term tiles of the chunk)
  static constexpr int kMaxThreads = 64, kMinBlocks = 1;
  int S, t0, nt, n_pts, n_ptiles; const double* xs; const double* ys; const double* recs; double* V /*[nt][n_pts]*/; int* bad;
  static MCE_HD size_t smem_bytes(int S) { return sizeof(double) * (size_t)cpdf2_rec_doubles(S) * CPDF2_TERMS; }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    double* sr = (double*)c.smem();
    const int R = cpdf2_rec_doubles(S), ptile = c.block() % n_ptiles, ttile = c.block() / n_ptiles;
    const int tb = ttile * CPDF2_TERMS, ntile = nt - tb < CPDF2_TERMS ? nt - tb : CPDF2_TERMS;
    const double* src = recs + (long long)(t0 + tb) * R;
    c.par([&](int tid) { for (int i = tid; i < ntile * R; i += c.nthreads()) sr[i] = src[i]; });
    c.par([&](int tid) {
      const int k = ptile * CPDF2_PTS + tid;
      if (k >= n_pts) return;
      const double x1 = xs[k], x2 = ys[k];
      int b = 0;
      for (int t = 0; t < ntile; t++) V[(long long)(tb + t) * n_pts + k] = cpdf2_term(sr + t * R, S, x1, x2, &b);
      if (b) *bad = 1;
    });
  }
};

struct KCpdf2dSum {             // thread / grid point: running sum += values of the chunk, in term order
  int nt, n_pts, first; const double* V; double* out;
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      const int k = c.block() * c.nthreads() + tid;
      if (k >= n_pts) return;
      double fx = first ? 0.0 : out[k];
      for (int t = 0; t < nt; t++) fx += V[(long long)t * n_pts + k];
      out[k] = fx;
    });
  }
};

// Device self-test of div_nobranch (mce_math.h): pseudo-random operand pairs over the whole exponent range plus mantissa
// edge patterns; counts the pairs whose flag is set but whose value differs from `a / b`, and the pairs that were flagged.
struct KDivSelfTest {
  unsigned long long seed; long long n; unsigned long long* counters;   // [0] mismatches, [1] ok pairs
  static MCE_HD unsigned long long mix(unsigned long long k) {
    unsigned long long x = (k + 1) * 0x9E3779B97F4A7C15ULL; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32; x *= 0x94D049BB133111EBULL; x ^= x >> 29; return x;
  }
  static MCE_HD double make(unsigned long long bits, int mode) {
    // mode 0: any exponent; 1: exponents near 1; 2: mantissa of all ones / single bits
    unsigned long long man = bits & 0xFFFFFFFFFFFFFULL, sign = (bits >> 63) << 63; long long ex = (long long)((bits >> 52) & 0x7FF);
    if (mode == 1) ex = 1023 + (ex % 41) - 20;
    if (mode == 2) { const int k = (int)(man % 53); man = (bits & (1ULL << 60)) ? (0xFFFFFFFFFFFFFULL >> k) : ((1ULL << k) >> 1); ex = 1023 + (ex % 201) - 100; }
    union { unsigned long long u; double d; } v; v.u = sign | ((unsigned long long)ex << 52) | man; return v.d;
  }
  template <class Ctx> MCE_KERNEL_FN void run(Ctx& c) const {
    c.par([&](int tid) {
      unsigned long long bad = 0, okc = 0;
      const long long stride = (long long)c.nthreads() * c.nblocks();
      for (long long i = (long long)c.block() * c.nthreads() + tid; i < n; i += stride) {
        const unsigned long long ra = mix(seed + 2 * (unsigned long long)i), rb = mix(seed + 2 * (unsigned long long)i + 1);
        const int mode = (int)(i % 3);
        const double a = make(ra, mode), b = make(rb, mode);
        bool ok; const double q = div_nobranch(a, b, &ok), ref = a / b;
        union { double d; unsigned long long u; } x, y; x.d = q; y.d = ref;
        if (ok) { okc++; if (x.u != y.u) bad++; }
      }
      if (bad) c.atomic_add_u64(counters, bad);
      if (okc) c.atomic_add_u64(counters + 1, okc);
    });
  }
};

}  // namespace mce
#endif
