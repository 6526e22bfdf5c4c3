// emu_backend.h -- TEST INFRASTRUCTURE ONLY. A sequential stand-in for the CUDA backend that lets the
// kernel bodies of cauchyfriendly_b200/csrc/mce_kern_*.h be unit-tested on machines without a GPU:
// every `ctx.par(f)` runs f(0..nthreads-1) in a loop, blocks run one after another.  It is compiled into
// tests/emu/_build/libmce_emu.so by tests/emu/build.sh and loaded only by `-m "not gpu"` tests; the product
// library libmce_b200.so contains the CUDA backend alone and fails to create a handle without a GPU.
#ifndef MCE_EMU_BACKEND_H_
#define MCE_EMU_BACKEND_H_

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../cauchyfriendly_b200/csrc/mce_shard.h"

namespace mce {

struct EmuCtx {
  int block_, nblocks_, nthreads_;
  unsigned char* smem_;
  int block() const { return block_; }
  int nblocks() const { return nblocks_; }
  int nthreads() const { return nthreads_; }
  unsigned char* smem() const { return smem_; }
  template <class F> void par(F&& f) { for (int t = 0; t < nthreads_; t++) f(t); }
  template <class T> T uniform(const T& v) { return v; }
  int atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
  unsigned atomic_xor(unsigned* p, unsigned v) { unsigned o = *p; *p ^= v; return o; }
  unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
  unsigned atomic_or(unsigned* p, unsigned v) { unsigned o = *p; *p |= v; return o; }
  unsigned atomic_cas(unsigned* p, unsigned cmp, unsigned v) { unsigned o = *p; if (o == cmp) *p = v; return o; }
  int load_relaxed(const int* p) { return *p; }
  int atomic_min(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
  template <class T> struct Priv { std::vector<T> v; T& operator[](int t) { return v[t]; } };
  template <class T> Priv<T> priv() const { return Priv<T>{std::vector<T>(nthreads_)}; }
  template <class T, class Op> void block_scan(T* x, T*, Op op) { for (int k = 1; k < nthreads_; k++) x[k] = op(x[k - 1], x[k]); }
  void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
  void cp_async_wait() {}
  void cp_async_commit() {}
  void cp_async_wait_pending2() {}
};

struct EmuBackend {
  long long launch_count = 0;
  static bool available(int, std::string*) { return true; }
  bool init(int, std::string*) { return true; }
  void shutdown() {}
  void* alloc(size_t n) { void* p = malloc(n ? n : 1); memset(p, 0xCD, n); return p; }   // poison: catches reads of unwritten memory
  void free(void* p) { ::free(p); }
  void h2d(void* d, const void* s, size_t n) { memcpy(d, s, n); }
  void d2h(void* d, const void* s, size_t n) { memcpy(d, s, n); }
  void memset(void* p, int v, size_t n) { ::memset(p, v, n); }
  double tic() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  double toc(double t0) { return tic() - t0; }
  double ev_t[24] = {0};
  void ev_record(int i) { ev_t[i] = tic(); }
  void ev_record_side(int i) { ev_t[i] = tic(); }
  double ev_elapsed(int i0, int i1) { return ev_t[i1] - ev_t[i0]; }
  void ev_wait(int) {}
  template <class K> void launch(const K& k, int nblocks, int nthreads, size_t smem) {
    launch_count++;
    std::vector<unsigned char> sm(smem + 64, 0xCD);
    for (int b = 0; b < nblocks; b++) {
      EmuCtx c{b, nblocks, nthreads, sm.data()};
      k.run(c);
    }
  }
  void sync() {}
  // exchange layer: callback transport only (tests drive it over torch.distributed / gloo)
  ShardInfo shard;
  bool shard_unique_id(void*, std::string* why) { *why = "the emulation backend has no native transport"; return false; }
  bool shard_init_native(int, int, const void*, std::string* why) { *why = "the emulation backend has no native transport"; return false; }
  void shard_init_callback(int rank, int world, mce_exchange_fn fn, void* ctx) { shard.rank = rank; shard.world = world; shard.fn = fn; shard.fn_ctx = ctx; }
  void xchg_begin() {}
  void xchg_end() {}
  void xchg_allgather(void* base, size_t bytes_per_rank) { if (bytes_per_rank && shard.fn(shard.fn_ctx, MCE_XCHG_ALLGATHER, base, (long long)bytes_per_rank) != 0) throw std::runtime_error("exchange callback failed"); }
  void xchg_allreduce_u32(void* base, size_t n) { if (n && shard.fn(shard.fn_ctx, MCE_XCHG_ALLREDUCE_SUM_U32, base, (long long)n) != 0) throw std::runtime_error("exchange callback failed"); }
  void xchg_alltoallv(const void* send, const long long* soff, const long long* scnt, void* recv, const long long* roff, const long long* rcnt) {
    mce_alltoallv_args a{send, recv, soff, scnt, roff, rcnt};
    if (shard.fn(shard.fn_ctx, MCE_XCHG_ALLTOALLV, &a, (long long)shard.world) != 0) throw std::runtime_error("exchange callback failed");
  }
  void make_current() {}
  void side_begin() {}
  template <class K> void launch_side(const K& k, int nblocks, int nthreads, size_t smem) { launch(k, nblocks, nthreads, smem); }
  void side_join() {}
  void sort_pairs(const unsigned long long* kin, unsigned long long* kout, const int* vin, int* vout, int n) {
    std::vector<int> idx(n); std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return kin[a] < kin[b]; });
    for (int i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
  }
  void exclusive_scan(const int* in, int* out, int n) { int acc = 0; for (int i = 0; i < n; i++) { int v = in[i]; out[i] = acc; acc += v; } }
};

}  // namespace mce
#endif
