"""ONE estimator partitioned over torch.distributed ranks (one process per GPU) -- host glue for the mce_shard_* entry
points of include/mce_b200.h (design: csrc/mce_kern_part.h).

Every rank creates the estimator with identical arguments, calls `init_term_sharding` once before the first step, and then
makes identical step calls.  Every term lives on one rank; moments and counts come back global on every rank.

transport="nccl": the library opens libnccl.so.2 itself and issues grouped send/recv, all-gathers and all-reduces on its own
CUDA stream; torch.distributed is only used to ship the 128-byte NCCL id from rank 0.  transport="callback": the library hands
every exchange to Python, which runs it over `dist` (used by the CPU tests with the gloo backend and host memory).
moments: "ordered" (every sum bit-identical to one GPU), "hybrid" (Re fz -- the sum that feeds back into the filter -- bit-identical through an exact
scan over all ranks' slots, the other sums per rank (two-level reductions) and added in rank order: every count, key and G stays bit-identical, mean / covariance move in the
last digits) or "allreduce" (all sums per rank; fz moves in the last bits and deep steps can lose or gain a term)."""
import ctypes as ct

import numpy as np

from . import _capi

_KEEP = []          # callback objects must outlive the handles that use them
EXCHANGES = [0, 0]   # callback transport: number of exchanges, bytes received per rank (diagnostics)


def init_term_sharding(handle, dist, lib=None, transport="nccl", device=-1, moments="ordered"):
    lib = lib or _capi.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    if lib.mce_shard_set_moments_mode(handle, {"ordered": 0, "allreduce": 1, "hybrid": 2}[moments]) != 0:
        raise RuntimeError("mce_shard_set_moments_mode failed")
    if transport == "nccl":
        buf = ct.create_string_buffer(128)
        if rank == 0 and lib.mce_shard_unique_id(device, buf) != 0:
            raise RuntimeError(lib.mce_last_error().decode())
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        ident = ct.create_string_buffer(box[0], 128)
        if lib.mce_shard_init(handle, rank, world, ident) != 0:
            raise RuntimeError(lib.mce_last_error().decode())
        return
    import torch

    def exchange(_ctx, op, base, n):
        try:
            EXCHANGES[0] += 1
            EXCHANGES[1] += n * (world - 1) if op == 0 else 4 * n
            if op == 2:          # personalised exchange (mce_alltoallv_args); gloo has no all_to_all: pairwise isend / irecv
                a = ct.cast(base, ct.POINTER(_capi.MceAllToAllV)).contents
                def view(addr, nbytes):
                    return torch.from_numpy(np.ctypeslib.as_array((ct.c_ubyte * nbytes).from_address(addr)))
                reqs, outs = [], []
                for h in range(world):
                    sc, rc = a.scnt[h], a.rcnt[h]
                    if h == rank:
                        if sc:
                            ct.memmove(a.recv + a.roff[h], a.send + a.soff[h], sc)
                        continue
                    if sc:
                        reqs.append(dist.isend(view(a.send + a.soff[h], sc).clone(), dst=h))
                    if rc:
                        buf = torch.empty(rc, dtype=torch.uint8)
                        reqs.append(dist.irecv(buf, src=h))
                        outs.append((buf, a.recv + a.roff[h], rc))
                    EXCHANGES[1] += rc
                for r in reqs:
                    r.wait()
                for buf, addr, rc in outs:
                    view(addr, rc)[:] = buf
            elif op == 0:          # in-place all-gather of `world` chunks of n bytes
                whole = torch.from_numpy(np.ctypeslib.as_array((ct.c_ubyte * (n * world)).from_address(base)))
                chunks = [torch.empty(n, dtype=torch.uint8) for _ in range(world)]
                dist.all_gather(chunks, whole[rank * n:(rank + 1) * n].clone())
                for r in range(world):
                    whole[r * n:(r + 1) * n] = chunks[r]
            else:                # in-place sum of n uint32 (two's complement: the int32 sum has the same bits)
                t = torch.from_numpy(np.ctypeslib.as_array((ct.c_int32 * n).from_address(base)))
                dist.all_reduce(t)
            return 0
        except Exception as e:      # noqa: BLE001 -- must not unwind through the C caller
            print("exchange callback failed:", e, flush=True)
            return 1

    fn = _capi.EXCHANGE_FN(exchange)
    _KEEP.append(fn)
    if lib.mce_shard_init_callback(handle, rank, world, fn, None) != 0:
        raise RuntimeError(lib.mce_last_error().decode())
