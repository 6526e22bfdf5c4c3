"""NVLink all-to-all probe (torchrun): bytes/s per rank of torch.distributed.all_to_all_single over NCCL for a few sizes,
and of plain peer-to-peer copies (cudaMemcpyPeer via tensor.copy_), as a yardstick for the partitioned estimator's exchanges."""
import os, time, torch, torch.distributed as dist
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W, R = dist.get_world_size(), dist.get_rank()
for mb in (64, 256, 1024):
    n = mb * (1 << 20) // 4 * W
    a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    for _ in range(2): dist.all_to_all_single(b, a)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(5): dist.all_to_all_single(b, a)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    if R == 0: print("all_to_all_single %5d MB per peer: %.2f ms, %.0f GB/s received per rank" % (mb, dt * 1e3, mb * (W - 1) / 1024 / dt), flush=True)
    del a, b
dist.barrier()
if R == 0:
    x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0"); y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:1")
    y.copy_(x); torch.cuda.synchronize(0); torch.cuda.synchronize(1); t0 = time.perf_counter()
    for _ in range(5): y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print("peer copy 1 GiB gpu0 -> gpu1: %.0f GB/s" % (5 * 1.0737 / (time.perf_counter() - t0)), flush=True)
dist.barrier(); dist.destroy_process_group()
