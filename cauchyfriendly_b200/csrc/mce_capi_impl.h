// mce_capi_impl.h -- extern "C" entry points of include/mce_b200.h over Engine<MCE_BACKEND>.
// Included exactly once by mce_capi.cu (CUDA backend -> libmce_b200.so).
#ifndef MCE_CAPI_IMPL_H_
#define MCE_CAPI_IMPL_H_

#include <new>
#include <string>

#include "../../include/mce_b200.h"
#include "mce_engine.h"

using EngineT = mce::Engine<MCE_BACKEND>;
struct mce_handle { EngineT* e; };
static thread_local std::string g_mce_error;

extern "C" {

void mce_default_options(mce_options* o) {
  memset(o, 0, sizeof(*o));
  o->device = -1;
  for (int i = 0; i < 12; i++) o->tr_search_order[i] = i;
}

mce_handle* mce_create(int d, int cmcc, int pncc, int p, int steps, const double* A0, const double* p0, const double* b0,
                       const double* root_point, const double* b_pert, const mce_options* opts) {
  mce_options def; mce_default_options(&def);
  if (!opts) opts = &def;
  if (!A0 || !p0 || !b0 || !root_point || !b_pert || p < 1 || steps < 1 || pncc < 0 || cmcc < 0) { g_mce_error = "mce_create: bad argument"; return nullptr; }
  const int max_shape = d > 1 ? (steps - 1) * pncc + d : d + pncc;
  std::string why;
  if (!EngineT::supported(d, max_shape, pncc, &why)) { g_mce_error = "mce_create: " + why; return nullptr; }
  std::string berr;
  if (!MCE_BACKEND::available(opts->device, &berr)) { g_mce_error = "mce_create: " + berr; return nullptr; }
  EngineT* e = new (std::nothrow) EngineT(d, cmcc, pncc, p, steps, A0, p0, b0, root_point, b_pert, opts->tr_search_order, opts->print_basic_info);
  if (!e) { g_mce_error = "mce_create: out of memory"; return nullptr; }
  e->fast_moments = opts->fast_moments != 0;
  e->phase_timing = opts->phase_timing != 0;
  e->lean_groups = opts->lean_group_kernel != 0;
  e->fast_moments_slots = opts->fast_moments_min_slots <= 0 ? 300000 : opts->fast_moments_min_slots;
  e->early_scale_slots = opts->early_scale_min_slots == 0 ? 400000 : (opts->early_scale_min_slots < 0 ? (1ll << 62) : opts->early_scale_min_slots);
  e->big_T = opts->group_split_threshold == 0 ? mce::BIG_T : (opts->group_split_threshold < 0 ? 0x7fffffff : opts->group_split_threshold);
  if (!e->be.init(opts->device, &berr)) { g_mce_error = "mce_create: " + berr; delete e; return nullptr; }
  mce_handle* h = new mce_handle; h->e = e;
  return h;
}

void mce_destroy(mce_handle* h) {
  if (!h) return;
  h->e->be.make_current();
  try { h->e->release_all(); } catch (...) {}
  h->e->be.shutdown();
  delete h->e; delete h;
}

int mce_step(mce_handle* h, double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H, double gamma,
             const double* B, const double* u) {
  if (!h || !H) { g_mce_error = "mce_step: bad argument"; return MCE_ERR_BAD_ARG; }
  EngineT* e = h->e;
  if (e->master_step > 0 && (e->master_step % e->p) == 0 && (!Phi || (e->pncc > 0 && (!Gamma || !beta)))) { g_mce_error = "mce_step: Phi/Gamma/beta required on a time-propagation step"; return MCE_ERR_BAD_ARG; }
  int rc;
  e->be.make_current();
  try { rc = e->step(msmt, Phi, Gamma, beta, H, gamma, B, u); }
  catch (const std::exception& ex) { g_mce_error = std::string("mce_step: ") + ex.what(); return MCE_ERR_CUDA; }
  if (rc < 0) g_mce_error = "mce_step: " + e->error;
  return rc;
}

int mce_get_moments(mce_handle* h, mce_moments* out) {
  if (!h || !out) return MCE_ERR_BAD_ARG;
  EngineT* e = h->e;
  memset(out, 0, sizeof(*out));
  out->fz[0] = e->fz.re; out->fz[1] = e->fz.im; out->fz_after_mu[0] = e->fz_mu.re; out->fz_after_mu[1] = e->fz_mu.im;
  for (int i = 0; i < e->d; i++) { out->mean[2 * i] = e->mean[i].re; out->mean[2 * i + 1] = e->mean[i].im; }
  for (int i = 0; i < e->d * e->d; i++) { out->cov[2 * i] = e->var[i].re; out->cov[2 * i + 1] = e->var[i].im; }
  for (int i = 0; i < e->d; i++) { out->mean_after_mu[2 * i] = e->mean_mu[i].re; out->mean_after_mu[2 * i + 1] = e->mean_mu[i].im; }
  for (int i = 0; i < e->d * e->d; i++) { out->cov_after_mu[2 * i] = e->var_mu[i].re; out->cov_after_mu[2 * i + 1] = e->var_mu[i].im; }
  out->g_scale_factor = e->G_SCALE_FACTOR; out->numeric_moment_errors = e->numeric_moment_errors;
  out->Nt = e->Nt; out->Nt_after_muc = e->Nt_muc; out->master_step = e->master_step; out->skip_post_mu = e->skip_post_mu;
  return 0;
}
int mce_shape_range(mce_handle* h) { return h ? h->e->shape_range : MCE_ERR_BAD_ARG; }
int mce_get_terms_per_shape(mce_handle* h, int* counts, int after_muc) {
  if (!h || !counts) return MCE_ERR_BAD_ARG;
  const std::vector<int>& v = after_muc ? h->e->muc_per_shape : h->e->terms_per_shape;
  for (int i = 0; i < h->e->shape_range; i++) counts[i] = v[i];
  return 0;
}
void mce_set_master_step(mce_handle* h, int master_step) { if (h) h->e->master_step = master_step; }
int mce_reset(mce_handle* h) { if (!h) return MCE_ERR_BAD_ARG; h->e->reset(); return 0; }
int mce_reinitialize_start_statistics(mce_handle* h, const double* A0, const double* p0, const double* b0) {
  if (!h || !A0 || !p0 || !b0) return MCE_ERR_BAD_ARG;
  EngineT* e = h->e;
  e->A0.assign(A0, A0 + e->d * e->d); e->p0.assign(p0, p0 + e->d); e->b0.assign(b0, b0 + e->d);
  e->A1 = e->A0; e->p1 = e->p0; e->b1 = e->b0;      // est:1306-1308: the first term is re-seeded at once
  return 0;
}
int mce_set_first_term(mce_handle* h, const double* A, const double* p, const double* b) {
  if (!h || !A || !p || !b) return MCE_ERR_BAD_ARG;
  EngineT* e = h->e;
  if (e->master_step != 0) { g_mce_error = "mce_set_first_term: only before the first step of a window (master_step == 0)"; return MCE_ERR_STATE; }
  e->A1.assign(A, A + e->d * e->d); e->p1.assign(p, p + e->d); e->b1.assign(b, b + e->d);
  return 0;
}
int mce_shift_b(mce_handle* h, const double* delta, double sign) {
  if (!h || !delta) return MCE_ERR_BAD_ARG;
  h->e->be.make_current();
  try { return h->e->shift_b(delta, sign); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
int mce_deterministic_time_prop(mce_handle* h, const double* Phi, const double* B, const double* u) {
  if (!h || !Phi) return MCE_ERR_BAD_ARG;
  if ((B == nullptr) != (u == nullptr)) { g_mce_error = "mce_deterministic_time_prop: set both B and u or neither (est:1336-1344)"; return MCE_ERR_BAD_ARG; }
  h->e->be.make_current();
  try { return h->e->det_time_prop(Phi, B, u); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
int mce_export_shape(mce_handle* h, int m, int* n_terms, long long* n_cells_total, double* A, double* p, double* b, int* cells, uint32_t* keys, double* G) {
  if (!h || !n_terms || !n_cells_total) return MCE_ERR_BAD_ARG;
  h->e->be.make_current();
  try { return h->e->export_shape(m, n_terms, n_cells_total, A, p, b, cells, keys, G); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
int mce_cpdf_grid_count(double grid_low, double grid_high, double grid_res) {
  if (!(grid_high > grid_low) || !(grid_res > 0)) return MCE_ERR_BAD_ARG;                     // asserts of cpdf_ndim.hpp:2057-2058
  return (int)((grid_high - grid_low + grid_res - 1e-15) / grid_res) + 1;                   // cpdf_ndim.hpp:2060
}
int mce_marginal_1d_points(mce_handle* h, int marg_idx, const double* bar_nu, int n, const double* xs, double* ys) {
  if (!h || !bar_nu || !xs || !ys || n < 1) { g_mce_error = "mce_marginal_1d_points: bad argument"; return MCE_ERR_BAD_ARG; }
  h->e->be.make_current();
  int rc;
  try { rc = h->e->marginal_1d_points(marg_idx, bar_nu, n, xs, ys); } catch (const std::exception& ex) { g_mce_error = std::string("mce_marginal_1d_points: ") + ex.what(); return MCE_ERR_CUDA; }
  if (rc < 0) { g_mce_error = "mce_marginal_1d_points: " + h->e->error; return MCE_ERR_BAD_ARG; }
  return rc;
}
int mce_marginal_1d_grid(mce_handle* h, int marg_idx, const double* bar_nu, double grid_low, double grid_high, double grid_res, double* xy, int n_cap) {
  const int n = mce_cpdf_grid_count(grid_low, grid_high, grid_res);
  if (!h || !xy || n < 1 || n > n_cap) { g_mce_error = "mce_marginal_1d_grid: bad grid or capacity"; return MCE_ERR_BAD_ARG; }
  std::vector<double> xs(n), ys(n);
  for (int i = 0; i < n; i++) { double g = grid_low + i * grid_res; if (g > grid_high) g = grid_high; xs[i] = g; }   // cpdf_ndim.hpp:2063-2070
  const int rc = mce_marginal_1d_points(h, marg_idx, bar_nu, n, xs.data(), ys.data());
  if (rc <= 0) return rc;
  for (int i = 0; i < n; i++) { xy[2 * i] = xs[i]; xy[2 * i + 1] = ys[i]; }
  return n;
}
int mce_marginal_2d_points(mce_handle* h, int marg_idx1, int marg_idx2, const double* bar_nu, int n, const double* xs, const double* ys, double* zs) {
  if (!h || !bar_nu || !xs || !ys || !zs || n < 1) { g_mce_error = "mce_marginal_2d_points: bad argument"; return MCE_ERR_BAD_ARG; }
  h->e->be.make_current();
  int rc;
  try { rc = h->e->marginal_2d_points(marg_idx1, marg_idx2, bar_nu, n, xs, ys, zs); } catch (const std::exception& ex) { g_mce_error = std::string("mce_marginal_2d_points: ") + ex.what(); return MCE_ERR_CUDA; }
  if (rc < 0) { g_mce_error = "mce_marginal_2d_points: " + h->e->error; return MCE_ERR_BAD_ARG; }
  return rc;
}
int mce_marginal_2d_grid(mce_handle* h, int marg_idx1, int marg_idx2, const double* bar_nu, double xlo, double xhi, double xres, double ylo, double yhi, double yres,
                         double* xyz, int n_cap, int* nx_out, int* ny_out) {
  const int nx = mce_cpdf_grid_count(xlo, xhi, xres), ny = mce_cpdf_grid_count(ylo, yhi, yres);
  if (!h || !xyz || nx < 1 || ny < 1 || (long long)nx * ny > n_cap) { g_mce_error = "mce_marginal_2d_grid: bad grid or capacity"; return MCE_ERR_BAD_ARG; }
  const int n = nx * ny;
  std::vector<double> xs(n), ys(n), zs(n);
  for (int i = 0; i < ny; i++) {                                    // cpdf_ndim.hpp:1830-1847: y-major
    double gy = ylo + i * yres; if (gy > yhi) gy = yhi;
    for (int j = 0; j < nx; j++) { double gx = xlo + j * xres; if (gx > xhi) gx = xhi; xs[i * nx + j] = gx; ys[i * nx + j] = gy; }
  }
  const int rc = mce_marginal_2d_points(h, marg_idx1, marg_idx2, bar_nu, n, xs.data(), ys.data(), zs.data());
  if (rc <= 0) return rc;
  for (int k = 0; k < n; k++) { xyz[3 * k] = xs[k]; xyz[3 * k + 1] = ys[k]; xyz[3 * k + 2] = zs[k]; }
  if (nx_out) *nx_out = nx;
  if (ny_out) *ny_out = ny;
  return n;
}
double mce_cpdf_last_ms(mce_handle* h) { return h ? h->e->cpdf_ms : 0.0; }
int mce_get_step_stats(mce_handle* h, mce_step_stats* out) {
  if (!h || !out) return MCE_ERR_BAD_ARG;
  const mce::StepStats& s = h->e->stats;
  memset(out, 0, sizeof(*out));
  out->ms_total = s.ms_total; out->ms_tp = s.ms_tp; out->ms_mu = s.ms_mu; out->ms_moments = s.ms_moments; out->ms_regroup = s.ms_regroup;
  out->ms_ftr = s.ms_ftr; out->ms_gtable = s.ms_gtable; out->ms_compact = s.ms_compact;
  out->parents = s.parents; out->slots = s.slots; out->terms_after_muc = s.terms_after_muc; out->groups = s.groups; out->survivors = s.survivors;
  out->bytes_gtable_algorithmic = s.bytes_gtable; out->bytes_step_algorithmic = s.bytes_step; out->kernel_launches = s.launches;
  out->ev_step_ms = s.ev_step_ms; out->ev_gtable_ms = s.ev_gtable_ms; out->gtable_launches = s.gtable_launches; out->gtable_lean_launches = s.gtable_lean_launches;
  out->cells_parents = s.cells_parents; out->cells_survivors = s.cells_survivors; out->split_groups = s.big_groups; out->ev_moments_ms = s.ev_moments_ms; out->ev_ftr_ms = s.ev_ftr_ms; out->ev_mu_ms = s.ev_mu_ms;
  out->ftr_rounds_max = s.ftr_rounds_max; out->diag_unmodelled_alias = s.diag_alias; out->diag_hash_overflow = s.diag_hash;
  return 0;
}
int mce_debug_div_selftest(mce_handle* h, long long n, unsigned long long seed, unsigned long long* out) {
  if (!h || !out || n < 1) return MCE_ERR_BAD_ARG;
  h->e->be.make_current();
  try { return h->e->div_selftest(n, seed, out); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
int mce_debug_moment_sums(mce_handle* h, long long n, int d, const double* g, const double* y, double* out) {
  if (!h || n < 0 || d < 0 || d > mce::MAXD || !out || (n > 0 && !g) || (n > 0 && d > 0 && !y)) return MCE_ERR_BAD_ARG;
  h->e->be.make_current();
  try { return h->e->debug_moment_sums(n, d, g, y, out); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
long long mce_debug_export_slots(mce_handle* h, long long cap, double* g, double* y) {
  if (!h || cap < 0) return MCE_ERR_BAD_ARG;
  h->e->be.make_current();
  try { return h->e->debug_export_slots(cap, g, y); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
int mce_debug_sum_scan(mce_handle* h, long long n, const double* g, double* out) {
  if (!h || n < 0 || !out || (n > 0 && !g)) return MCE_ERR_BAD_ARG;
  h->e->be.make_current();
  try { return h->e->debug_sum_scan(n, g, out); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
int mce_debug_capture(mce_handle* h, int enable) { if (!h) return MCE_ERR_BAD_ARG; h->e->capture = enable != 0; return 0; }
int mce_debug_muc_shape(mce_handle* h, int m, int* n_terms, double* A, double* p, double* q, double* b, double* cd, int* meta, uint8_t* cmap, int8_t* csmap, int* F) {
  if (!h || !n_terms) return MCE_ERR_BAD_ARG;
  *n_terms = 0;
  for (auto& cs : h->e->cap) {
    if (cs.m != m) continue;
    *n_terms = cs.n;
    if (A) {
      memcpy(A, cs.A.data(), cs.A.size() * 8); memcpy(p, cs.p.data(), cs.p.size() * 8); memcpy(q, cs.q.data(), cs.q.size() * 8);
      memcpy(b, cs.b.data(), cs.b.size() * 8); memcpy(cd, cs.cd.data(), cs.cd.size() * 8); memcpy(meta, cs.meta.data(), cs.meta.size() * 4);
      memcpy(cmap, cs.cmap.data(), cs.cmap.size()); memcpy(csmap, cs.csmap.data(), cs.csmap.size()); memcpy(F, cs.F.data(), cs.F.size() * 4);
    }
  }
  return 0;
}
int mce_shard_unique_id(int device, void* id128) {
  if (!id128) return MCE_ERR_BAD_ARG;
  MCE_BACKEND be; std::string why;
  (void)device;
  if (!be.shard_unique_id(id128, &why)) { g_mce_error = "mce_shard_unique_id: " + why; return MCE_ERR_CUDA; }
  return 0;
}
static int shard_check(mce_handle* h, int rank, int world) {
  if (!h || world < 1 || rank < 0 || rank >= world) { g_mce_error = "mce_shard_init: bad argument"; return MCE_ERR_BAD_ARG; }
  if (h->e->max_shape > 16) { g_mce_error = "mce_shard_init: term-level sharding needs at most 16 hyperplanes per term"; return MCE_ERR_BAD_ARG; }
  if (world > mce::PART_MAXW) { g_mce_error = "mce_shard_init: at most 16 ranks per estimator"; return MCE_ERR_BAD_ARG; }
  if (h->e->print_basic_info) { g_mce_error = "mce_shard_init: print_basic_info (post-reduction moment re-evaluation) is not available on a partitioned estimator"; return MCE_ERR_BAD_ARG; }
  if (h->e->master_step != 0) { g_mce_error = "mce_shard_init: call before the first step"; return MCE_ERR_BAD_ARG; }
  return 0;
}
int mce_shard_init(mce_handle* h, int rank, int world, const void* id128) {
  const int rc = shard_check(h, rank, world);
  if (rc) return rc;
  if (!id128) return MCE_ERR_BAD_ARG;
  std::string why;
  if (!h->e->be.shard_init_native(rank, world, id128, &why)) { g_mce_error = "mce_shard_init: " + why; return MCE_ERR_CUDA; }
  return 0;
}
int mce_shard_init_callback(mce_handle* h, int rank, int world, mce_exchange_fn fn, void* ctx) {
  const int rc = shard_check(h, rank, world);
  if (rc) return rc;
  if (!fn) return MCE_ERR_BAD_ARG;
  h->e->be.shard_init_callback(rank, world, (mce::mce_exchange_fn)fn, ctx);
  return 0;
}

int mce_shard_set_moments_mode(mce_handle* h, int mode) {
  if (!h || mode < 0 || mode > 2) return MCE_ERR_BAD_ARG;
  h->e->moments_mode = mode;
  return 0;
}
int mce_shard_export_gpos(mce_handle* h, int* out, int cap) {
  if (!h) return MCE_ERR_BAD_ARG;
  h->e->be.make_current();
  try { return h->e->export_gpos(out, cap); } catch (const std::exception& ex) { g_mce_error = ex.what(); return MCE_ERR_CUDA; }
}
int mce_shard_get_stats(mce_handle* h, mce_shard_stats* out) {
  if (!h || !out) return MCE_ERR_BAD_ARG;
  memset(out, 0, sizeof(*out));
  out->rank = h->e->be.shard.rank; out->world = h->e->be.shard.world;
  out->owned_terms = h->e->pstats.owned; out->imported_parents = h->e->pstats.imports; out->local_parents = h->e->gen[h->e->cur].v.n_alive;
  out->bytes_terms = h->e->pstats.bytes_terms; out->bytes_parents = h->e->pstats.bytes_parents; out->bytes_moments = h->e->pstats.bytes_moments; out->bytes_keys = h->e->pstats.bytes_keys;
  for (int i = 0; i < 8; i++) out->ms_stage[i] = h->e->pstats.ms[i];
  return 0;
}

const char* mce_last_error(void) { return g_mce_error.c_str(); }
const char* mce_version(void) { return MCE_VERSION_STRING; }

}  // extern "C"
#endif
