"""Concurrent pinned H2D/D2H bandwidth of all ranks, with and without binding each rank to its GPU's NUMA-local CPUs.
run under torchrun (one rank per GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import pynvml

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(local)
ncpu = os.cpu_count()
words = (ncpu + 63) // 64
try:
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
    cpus = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
except Exception as e:
    cpus = []
    print(rank, "affinity query failed", e)
print(f"rank {rank}: os.cpu_count {ncpu}, current affinity {sorted(os.sched_getaffinity(0))}, GPU-local cpus {cpus}", flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -24; lscpu | grep -i -E 'numa|socket|model name' ")


def host_alloc(n, write_combined):
    """n bytes of page-locked host memory as a uint8 tensor; write_combined: cudaHostAllocWriteCombined (not snooped, slow to read
    from the CPU, meant for buffers the CPU only writes and the GPU only reads)."""
    if not write_combined:
        return torch.empty(n, dtype=torch.uint8, pin_memory=True)
    import ctypes
    import numpy as np
    rt = ctypes.CDLL("libcudart.so")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(0x04))
    assert rc == 0, rc
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(p.value)))


def measure(tag, write_combined=False):
    n = 343_000_000
    host = host_alloc(n, write_combined)
    host[::4096] = 1                      # touch every page
    print(f"rank {rank}: {tag}: is_pinned {host.is_pinned()}", flush=True) if rank == 0 else None
    out = torch.empty(196_000_000, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    for both in (False, True):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            with torch.cuda.stream(up):
                dev.copy_(host, non_blocking=True)
            if both:
                with torch.cuda.stream(down):
                    out.copy_(dev[:196_000_000], non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        gb = 10 * (n + (196_000_000 if both else 0)) / 1e9
        if rank == 0:
            print(f"{tag} {'H2D+D2H' if both else 'H2D only'}: {gb / dt.item():.1f} GB/s per rank, {world * gb / dt.item():.1f} GB/s aggregate", flush=True)
    del host, out


measure("default placement")
measure("write-combined source", write_combined=True)
if cpus:
    os.sched_setaffinity(0, cpus)
    measure("bound to GPU-local CPUs")
dist.destroy_process_group()
