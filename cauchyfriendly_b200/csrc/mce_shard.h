// mce_shard.h -- exchange layer for term-level sharding of ONE estimator over several ranks.
//
// Scheme ("replicated state, sharded kernels"): every rank holds the whole term list and runs the cheap, global parts of a
// step (measurement update, term reduction, moments) redundantly -- they are deterministic, so all ranks stay bit-identical.
// The two heavy per-item kernels are split by item range: DCE-TP over the parents, the G-table kernel over the reduction
// groups of each (phase, shape).  After each of them the ranks all-gather what they produced, in place: the layouts give
// every rank an equal, contiguous chunk of every output array.  The only reduction is the per-parent re-orientation mask
// (at most one rank touches a parent, so a sum is the value).
//
// Transports: NCCL (CUDA backend; libnccl.so.2 is opened at run time, grouped calls on the engine's stream) or a host
// callback (any backend; used by the CPU tests over torch.distributed/gloo).
#ifndef MCE_SHARD_H_
#define MCE_SHARD_H_

#include <stddef.h>

namespace mce {

enum { MCE_XCHG_ALLGATHER = 0, MCE_XCHG_ALLREDUCE_SUM_U32 = 1, MCE_XCHG_ALLTOALLV = 2 };
// op ALLGATHER: `base` holds world chunks of n bytes, chunk `rank` is valid on entry, all are valid on return.
// op ALLREDUCE_SUM_U32: n 32-bit unsigned values at `base`, summed over the ranks in place.  Returns 0 on success.
// op ALLTOALLV: `base` points to a mce_alltoallv_args, n = world: scnt[h] bytes at send + soff[h] go to rank h, rcnt[h] bytes
// from rank h land at recv + roff[h] (the rank's own chunk included).
struct mce_alltoallv_args { const void* send; void* recv; const long long* soff; const long long* scnt; const long long* roff; const long long* rcnt; };
typedef int (*mce_exchange_fn)(void* ctx, int op, void* base, long long n);

struct ShardInfo {
  int rank = 0, world = 1;
  mce_exchange_fn fn = nullptr; void* fn_ctx = nullptr;   // callback transport (nullptr: the backend's native transport)
};

// equal split of n items over `world` ranks: chunk size and this rank's [lo, hi)
inline int shard_chunk(int n, int world) { return (n + world - 1) / world; }
inline void shard_range(int n, int rank, int world, int* lo, int* hi) {
  const int ch = shard_chunk(n, world);
  *lo = rank * ch < n ? rank * ch : n;
  *hi = (rank + 1) * ch < n ? (rank + 1) * ch : n;
}
inline int round_up(int n, int w) { return (n + w - 1) / w * w; }

}  // namespace mce
#endif
