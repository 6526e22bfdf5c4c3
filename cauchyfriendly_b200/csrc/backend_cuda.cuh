// backend_cuda.cuh -- the CUDA (sm_100a) backend of the MCE engine: device memory, one stream, kernel launch
// of the functors in mce_kern_*.h, and the two library primitives the path uses (CUB radix sort of the
// FTR axis keys / reduction-group keys, CUB exclusive scan).  Everything else is hand-written kernels.
#ifndef MCE_BACKEND_CUDA_CUH_
#define MCE_BACKEND_CUDA_CUH_

#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <atomic>
#include <stdexcept>
#include <string>
#include <string.h>

#include <dlfcn.h>
#include "mce_exec.h"
#include "mce_shard.h"

namespace mce {

#define MCE_CUDA_CHECK(expr)                                                                              \
  do {                                                                                                    \
    cudaError_t e__ = (expr);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
      throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

// Kernels may declare `static constexpr int kMaxThreads, kMinBlocks` to bound their register allocation.
template <class K, class = void> struct launch_traits { static constexpr int max_threads = 1024, min_blocks = 1; };
template <class K> struct launch_traits<K, decltype((void)K::kMaxThreads)> { static constexpr int max_threads = K::kMaxThreads, min_blocks = K::kMinBlocks; };

template <class K>
__global__ void __launch_bounds__(launch_traits<K>::max_threads, launch_traits<K>::min_blocks) mce_kernel_entry(const __grid_constant__ K k) {
  extern __shared__ __align__(16) unsigned char mce_dyn_smem[];
  DevCtx c;
  c.smem_ = mce_dyn_smem;
  k.run(c);
}

constexpr int kMaxDynSmem = 227 * 1024;      // sm_100: 227 KB of dynamic shared memory per CTA

struct CudaBackend {
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;       // second stream: the serial moment sums overlap regroup + FTR
  cudaEvent_t side_ev = nullptr;
  int device = 0;
  long long launch_count = 0;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  size_t bytes_allocated = 0;

  static bool available(int dev, std::string* why) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { *why = std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); libmce_b200 has no CPU fallback"; return false; }
    if (dev >= n) { *why = "device ordinal out of range"; return false; }
    return true;
  }
  bool init(int dev, std::string* why) {
    try {
      if (dev >= 0) MCE_CUDA_CHECK(cudaSetDevice(dev));
      MCE_CUDA_CHECK(cudaGetDevice(&device));
      MCE_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      MCE_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
      MCE_CUDA_CHECK(cudaEventCreateWithFlags(&side_ev, cudaEventDisableTiming));
    } catch (const std::exception& ex) { *why = ex.what(); return false; }
    return true;
  }

  // ---- exchange layer (mce_shard.h): NCCL opened at run time, or a host callback ----
  struct NcclId { char b[128]; };       // ncclUniqueId (passed by value to ncclCommInitRank)
  struct Nccl {
    void* lib = nullptr; void* comm = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr; int (*GroupEnd)() = nullptr; int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
  } nccl;
  ShardInfo shard;
  bool nccl_load(std::string* why) {
    if (nccl.lib) return true;
    nccl.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!nccl.lib) { *why = std::string("cannot open libnccl.so.2: ") + dlerror(); return false; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(nccl.lib, n); if (!p) { ok = false; *why = std::string("libnccl.so.2 lacks ") + n; } return p; };
    nccl.GetUniqueId = (decltype(nccl.GetUniqueId))sym("ncclGetUniqueId");
    nccl.CommInitRank = (decltype(nccl.CommInitRank))sym("ncclCommInitRank");
    nccl.AllGather = (decltype(nccl.AllGather))sym("ncclAllGather");
    nccl.AllReduce = (decltype(nccl.AllReduce))sym("ncclAllReduce");
    nccl.Send = (decltype(nccl.Send))sym("ncclSend");
    nccl.Recv = (decltype(nccl.Recv))sym("ncclRecv");
    nccl.GroupStart = (decltype(nccl.GroupStart))sym("ncclGroupStart");
    nccl.GroupEnd = (decltype(nccl.GroupEnd))sym("ncclGroupEnd");
    nccl.CommDestroy = (decltype(nccl.CommDestroy))sym("ncclCommDestroy");
    nccl.GetErrorString = (decltype(nccl.GetErrorString))sym("ncclGetErrorString");
    return ok;
  }
  void nccl_check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + (nccl.GetErrorString ? nccl.GetErrorString(rc) : "NCCL error"));
  }
  bool shard_unique_id(void* id128, std::string* why) {
    if (!nccl_load(why)) return false;
    const int rc = nccl.GetUniqueId(id128);
    if (rc != 0) { *why = std::string("ncclGetUniqueId: ") + nccl.GetErrorString(rc); return false; }
    return true;
  }
  bool shard_init_native(int rank, int world, const void* id128, std::string* why) {
    if (!nccl_load(why)) return false;
    make_current();
    NcclId id; memcpy(id.b, id128, 128);
    const int rc = nccl.CommInitRank(&nccl.comm, world, id, rank);
    if (rc != 0) { *why = std::string("ncclCommInitRank: ") + nccl.GetErrorString(rc); return false; }
    shard.rank = rank; shard.world = world; shard.fn = nullptr; shard.fn_ctx = nullptr;
    return true;
  }
  void shard_init_callback(int rank, int world, mce_exchange_fn fn, void* ctx) { shard.rank = rank; shard.world = world; shard.fn = fn; shard.fn_ctx = ctx; }
  // grouped exchanges: with NCCL the calls between begin/end become one launch on the engine's stream
  void xchg_begin() { if (!shard.fn) nccl_check(nccl.GroupStart(), "ncclGroupStart"); else MCE_CUDA_CHECK(cudaStreamSynchronize(stream)); }
  void xchg_end() { if (!shard.fn) nccl_check(nccl.GroupEnd(), "ncclGroupEnd"); }
  void xchg_allgather(void* base, size_t bytes_per_rank) {
    if (bytes_per_rank == 0) return;
    if (shard.fn) { if (shard.fn(shard.fn_ctx, MCE_XCHG_ALLGATHER, base, (long long)bytes_per_rank) != 0) throw std::runtime_error("exchange callback failed"); return; }
    nccl_check(nccl.AllGather((const char*)base + (size_t)shard.rank * bytes_per_rank, base, bytes_per_rank, /*ncclInt8*/ 0, nccl.comm, stream), "ncclAllGather");
  }
  void xchg_allreduce_u32(void* base, size_t n) {
    if (n == 0) return;
    if (shard.fn) { if (shard.fn(shard.fn_ctx, MCE_XCHG_ALLREDUCE_SUM_U32, base, (long long)n) != 0) throw std::runtime_error("exchange callback failed"); return; }
    nccl_check(nccl.AllReduce(base, base, n, /*ncclUint32*/ 3, /*ncclSum*/ 0, nccl.comm, stream), "ncclAllReduce");
  }

  // Personalised exchange: `cnt[h]` bytes at `send + soff[h]` go to rank h, `rcnt[h]` bytes from rank h land at `recv + roff[h]`
  // (grouped ncclSend / ncclRecv over NVLink; the rank's own chunk is a device-to-device copy).  Call between xchg_begin / xchg_end.
  void xchg_alltoallv(const void* send, const long long* soff, const long long* scnt, void* recv, const long long* roff, const long long* rcnt) {
    if (shard.fn) {
      mce_alltoallv_args a{send, recv, soff, scnt, roff, rcnt};
      if (shard.fn(shard.fn_ctx, MCE_XCHG_ALLTOALLV, &a, (long long)shard.world) != 0) throw std::runtime_error("exchange callback failed");
      return;
    }
    for (int h = 0; h < shard.world; h++) {
      if (h == shard.rank) {
        if (scnt[h] > 0) MCE_CUDA_CHECK(cudaMemcpyAsync((char*)recv + roff[h], (const char*)send + soff[h], (size_t)scnt[h], cudaMemcpyDeviceToDevice, stream));
        continue;
      }
      if (scnt[h] > 0) nccl_check(nccl.Send((const char*)send + soff[h], (size_t)scnt[h], /*ncclInt8*/ 0, h, nccl.comm, stream), "ncclSend");
      if (rcnt[h] > 0) nccl_check(nccl.Recv((char*)recv + roff[h], (size_t)rcnt[h], /*ncclInt8*/ 0, h, nccl.comm, stream), "ncclRecv");
    }
  }

  // the calling host thread may be new (window banks step their estimators from a thread pool): bind it to this device
  void make_current() { cudaSetDevice(device); }
  void shutdown() {
    if (nccl.comm && nccl.CommDestroy) { nccl.CommDestroy(nccl.comm); nccl.comm = nullptr; }
    if (cub_tmp) cudaFree(cub_tmp);
    cub_tmp = nullptr; cub_tmp_bytes = 0;
    for (int i = 0; i < 24; i++) if (evs[i]) { cudaEventDestroy(evs[i]); evs[i] = nullptr; }
    if (side_ev) cudaEventDestroy(side_ev);
    if (side) cudaStreamDestroy(side);
    side = nullptr; side_ev = nullptr;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
  void* alloc(size_t n) {
    void* p = nullptr;
    MCE_CUDA_CHECK(cudaSetDevice(device));
    MCE_CUDA_CHECK(cudaMalloc(&p, n ? n : 1));
    bytes_allocated += n;
    return p;
  }
  void free(void* p) {
    cudaStreamSynchronize(stream);
    if (side) cudaStreamSynchronize(side);
    cudaFree(p);
  }
  void h2d(void* d, const void* s, size_t n) { MCE_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream)); MCE_CUDA_CHECK(cudaStreamSynchronize(stream)); }
  void d2h(void* d, const void* s, size_t n) { MCE_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream)); MCE_CUDA_CHECK(cudaStreamSynchronize(stream)); }
  void memset(void* p, int v, size_t n) { MCE_CUDA_CHECK(cudaMemsetAsync(p, v, n, stream)); }
  void sync() { MCE_CUDA_CHECK(cudaStreamSynchronize(stream)); }
  double tic() {
    MCE_CUDA_CHECK(cudaStreamSynchronize(stream));
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  double toc(double t0) { return tic() - t0; }
  // device-side timing with CUDA events on the launching stream (no host synchronisation until ev_elapsed)
  cudaEvent_t evs[24] = {};
  void ev_record(int i) {
    if (!evs[i]) MCE_CUDA_CHECK(cudaEventCreate(&evs[i]));
    MCE_CUDA_CHECK(cudaEventRecord(evs[i], stream));
  }
  void ev_record_side(int i) {
    if (!evs[i]) MCE_CUDA_CHECK(cudaEventCreate(&evs[i]));
    MCE_CUDA_CHECK(cudaEventRecord(evs[i], side));
  }
  void ev_wait(int i) { MCE_CUDA_CHECK(cudaEventSynchronize(evs[i])); }      // host waits for a recorded event
  double ev_elapsed(int i0, int i1) {
    float ms = 0;
    MCE_CUDA_CHECK(cudaEventSynchronize(evs[i1]));
    MCE_CUDA_CHECK(cudaEventElapsedTime(&ms, evs[i0], evs[i1]));
    return ms;
  }

  template <class K>
  void launch_on(cudaStream_t st, const K& k, int nblocks, int nthreads, size_t smem) {
    if (nblocks <= 0) return;
    static_assert(sizeof(K) <= 32000, "kernel functor exceeds the 32 KB parameter space (CUDA >= 12.1, sm_70+)");
    if (smem > 48 * 1024) {
      // The attribute belongs to (kernel, device), not to an engine: several estimators are stepped from a thread pool on one GPU
      // (SlidingWindowBank, concurrent = True), so every kernel template is opted in ONCE per device, to the device maximum --
      // setting the exact per-launch size would let another thread shrink it between this thread's set and its launch.
      static std::atomic<unsigned> opted{0};
      const unsigned bit = 1u << (device & 31);
      if (!(opted.load(std::memory_order_acquire) & bit)) {
        MCE_CUDA_CHECK(cudaFuncSetAttribute(mce_kernel_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        opted.fetch_or(bit, std::memory_order_release);
      }
      if (smem > (size_t)kMaxDynSmem) throw std::runtime_error("kernel needs " + std::to_string(smem) + " B of shared memory; the device limit is " + std::to_string(kMaxDynSmem));
    }
    mce_kernel_entry<K><<<nblocks, nthreads, smem, st>>>(k);
    MCE_CUDA_CHECK(cudaGetLastError());
    launch_count++;
  }
  template <class K> void launch(const K& k, int nblocks, int nthreads, size_t smem) { launch_on(stream, k, nblocks, nthreads, smem); }
  // side stream: starts after everything queued on the main stream so far; side_join() blocks the host until it is idle
  void side_begin() { MCE_CUDA_CHECK(cudaEventRecord(side_ev, stream)); MCE_CUDA_CHECK(cudaStreamWaitEvent(side, side_ev, 0)); }
  template <class K> void launch_side(const K& k, int nblocks, int nthreads, size_t smem) { launch_on(side, k, nblocks, nthreads, smem); }
  void side_join() { MCE_CUDA_CHECK(cudaStreamSynchronize(side)); }
  void ensure_tmp(size_t bytes) {
    if (bytes > cub_tmp_bytes) {
      if (cub_tmp) { cudaStreamSynchronize(stream); cudaFree(cub_tmp); }
      MCE_CUDA_CHECK(cudaMalloc(&cub_tmp, bytes + bytes / 2 + 256));
      cub_tmp_bytes = bytes + bytes / 2 + 256;
    }
  }
  void sort_pairs(const unsigned long long* kin, unsigned long long* kout, const int* vin, int* vout, int n) {
    size_t bytes = 0;
    MCE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, 64, stream));
    ensure_tmp(bytes);
    MCE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, bytes, kin, kout, vin, vout, n, 0, 64, stream));
    launch_count++;
  }
  void exclusive_scan(const int* in, int* out, int n) {
    size_t bytes = 0;
    MCE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, stream));
    ensure_tmp(bytes);
    MCE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(cub_tmp, bytes, in, out, n, stream));
    launch_count++;
  }
};

}  // namespace mce
#endif
