// mce_exec.h -- execution contexts for the MCE kernels.
//
// Every kernel body in mce_kernels.h is written against a small "Ctx" interface in bulk-synchronous
// style: block-uniform control flow outside `ctx.par(...)`, per-thread work inside it, one barrier at the
// end of every `par`.  DevCtx (below) is the CUDA implementation: par(f) runs f(threadIdx.x) followed by
// __syncthreads().  tests/emu/emu_backend.h provides a sequential implementation of the same interface
// that exists only to unit-test kernel logic on machines without a GPU; it is never part of libmce_b200.so.
#ifndef MCE_EXEC_H_
#define MCE_EXEC_H_

#include "mce_math.h"

namespace mce {

#if defined(__CUDACC__)
struct DevCtx {
  unsigned char* smem_;
  __device__ __forceinline__ int block() const { return (int)blockIdx.x; }
  __device__ __forceinline__ int nblocks() const { return (int)gridDim.x; }
  __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
  __device__ __forceinline__ unsigned char* smem() const { return smem_; }
  template <class F>
  __device__ __forceinline__ void par(F&& f) {
    f((int)threadIdx.x);
    __syncthreads();
  }
  // Reads a block-uniform value from shared memory for use in control flow; the trailing barrier keeps a
  // fast thread from overwriting it (in a later phase) before a slow thread has read it.
  template <class T>
  __device__ __forceinline__ T uniform(const T& v) {
    T r = v;
    __syncthreads();
    return r;
  }
  __device__ __forceinline__ int atomic_add(int* p, int v) { return atomicAdd(p, v); }
  __device__ __forceinline__ unsigned atomic_xor(unsigned* p, unsigned v) { return atomicXor(p, v); }
  __device__ __forceinline__ unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
  __device__ __forceinline__ unsigned atomic_or(unsigned* p, unsigned v) { return atomicOr(p, v); }
  __device__ __forceinline__ unsigned atomic_cas(unsigned* p, unsigned cmp, unsigned v) { return atomicCAS(p, cmp, v); }
  __device__ __forceinline__ int load_relaxed(const int* p) { return *(const volatile int*)p; }
  __device__ __forceinline__ int atomic_min(int* p, int v) { return atomicMin(p, v); }
  // Per-thread value that lives across phases: a register here, one element per emulated thread in tests/emu.
  template <class T> struct Priv { T v; __device__ __forceinline__ T& operator[](int) { return v; } };
  template <class T> __device__ __forceinline__ Priv<T> priv() const { return Priv<T>(); }
  // Block-wide inclusive scan of n == nthreads() elements held in shared memory, in place: x[k] <- op(x[0], ..., x[k]) folded left to right with an
  // ASSOCIATIVE op(earlier, later).  Warp shuffles inside a warp, one warp over the warp totals; `scratch` holds nthreads() / 32 elements.
  // Called from block-uniform code (outside par); ends with a barrier.
  template <class T, class Op>
  __device__ __forceinline__ void block_scan(T* x, T* scratch, Op op) {
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = ((int)blockDim.x + 31) >> 5;
    T v = x[tid];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      T o = v;
      unsigned long long* w = (unsigned long long*)&o;
#pragma unroll
      for (int i = 0; i < (int)(sizeof(T) / 8); i++) w[i] = __shfl_up_sync(0xffffffffu, w[i], off);
      if (lane >= off) v = op(o, v);
    }
    if (lane == 31) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
      T t = scratch[lane < nw ? lane : 0];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        T o = t;
        unsigned long long* w = (unsigned long long*)&o;
#pragma unroll
        for (int i = 0; i < (int)(sizeof(T) / 8); i++) w[i] = __shfl_up_sync(0xffffffffu, w[i], off);
        if (lane >= off && lane < nw) t = op(o, t);
      }
      if (lane < nw) scratch[lane] = t;
    }
    __syncthreads();
    if (warp > 0) v = op(scratch[warp - 1], v);
    x[tid] = v;
    __syncthreads();
  }
  // 16-byte asynchronous global -> shared copy (LDGSTS); the data is visible after cp_async_wait() + a barrier
  __device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
  }
  __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_all;" ::: "memory"); }
  // group-wise completion: commit closes the group of copies issued so far; wait_pending2 returns once at most the two
  // youngest groups of this thread are still in flight
  __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void cp_async_wait_pending2() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }
};
#endif

// Compiler-level fence: memory accesses are not moved across it (no instruction is emitted).
#if defined(__CUDA_ARCH__)
#define MCE_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define MCE_SCHED_FENCE() do {} while (0)
#endif

// Kernel functors expose: template <class Ctx> void run(Ctx&) const.
#if defined(__CUDACC__)
#define MCE_KERNEL_FN __device__
#else
#define MCE_KERNEL_FN
#endif

}  // namespace mce
#endif
