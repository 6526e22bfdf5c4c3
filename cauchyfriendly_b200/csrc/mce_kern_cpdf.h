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
