"""Host -> device copy ceiling with every rank copying at once (the end-to-end limiter of the N-GPU bench).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/scripts/h2d_bw.py
Each rank binds to its GPU's NUMA node (sysfs), allocates 192 MiB pinned, and copies it to its GPU 20 times
inside a barrier-to-barrier region; prints per-rank and aggregate GB/s and where each rank's buffer was bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import bench

rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
torch.cuda.set_device(lr)
for bind in (False, True):
    os.sched_setaffinity(0, range(os.cpu_count()))
    ncpu, node = bench.bind_to_gpu_numa(torch, lr) if bind else (len(os.sched_getaffinity(0)), -1)
    host = torch.empty(192 << 20, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty_like(host, device=f"cuda:{lr}")
    for _ in range(3):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dev.copy_(host, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lr}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    print(f"bind={bind} rank {rank}: {host.numel() / ms / 1e6:.1f} GB/s alone-in-parallel, numa node {node}, cpus bound {ncpu}", flush=True)
    if rank == 0:
        print(f"bind={bind} aggregate (slowest rank): {world * host.numel() / float(t[0]) / 1e6:.1f} GB/s over {world} GPUs", flush=True)
    del host, dev
if world > 1:
    dist.destroy_process_group()
