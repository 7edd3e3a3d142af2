"""What the host links deliver when every GPU of the box copies at once (the ceiling of bench.py's e2e at N > 1).

Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mb_pcie_multi.py
Every rank moves the e2e count step's bytes (800 MB in, 400 MB out) between pinned host memory and its own GPU:
H2D alone, D2H alone, both at once; all ranks start together (barrier), the time is the max over ranks.
Variants of the host side: plain pinned (cudaHostAlloc default, what torch pin_memory gives), write-combined query
buffers (cudaHostAllocWriteCombined), the H2D split over two streams, and each rank's pages first touched from a
thread pinned to a different core set (NUMA placement; all GPUs of this pool's boxes hang off NUMA node 0).
"""
import ctypes as C
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
H2D, D2H = 1, 2
NQ = 100_000_000
IN_B, OUT_B = NQ * 8, NQ * 4


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def maxr(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def host_alloc(nbytes, flags):
    p = C.c_void_p(0)
    rc = rt.cudaHostAlloc(C.byref(p), nbytes, flags)
    assert rc == 0, rc
    C.memset(p, 1, nbytes)  # first touch
    return p


d_in = torch.empty(IN_B, dtype=torch.uint8, device=dev)
d_out = torch.empty(OUT_B, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]


def copies(h_in, h_out, do_in, do_out, in_streams, piece):
    if do_in:
        k = 0
        for lo in range(0, IN_B, piece):
            m = min(piece, IN_B - lo)
            s = streams[k % in_streams]
            rt.cudaMemcpyAsync(C.c_void_p(d_in.data_ptr() + lo), C.c_void_p(h_in.value + lo), m, H2D, C.c_void_p(s.cuda_stream))
            k += 1
    if do_out:
        for lo in range(0, OUT_B, piece):
            m = min(piece, OUT_B - lo)
            rt.cudaMemcpyAsync(C.c_void_p(h_out.value + lo), C.c_void_p(d_out.data_ptr() + lo), m, D2H, C.c_void_p(streams[3].cuda_stream))


def measure(name, h_in, h_out, do_in, do_out, in_streams=1, piece=64 << 20, reps=4):
    for _ in range(2):
        copies(h_in, h_out, do_in, do_out, in_streams, piece)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        copies(h_in, h_out, do_in, do_out, in_streams, piece)
    torch.cuda.synchronize()
    dt = maxr(time.perf_counter() - t0) / reps
    barrier()
    if rank == 0:
        gin = world * IN_B / dt / 1e9 if do_in else 0.0
        gout = world * OUT_B / dt / 1e9 if do_out else 0.0
        print(f"N={world} {name}: {dt * 1e3:7.2f} ms per step; all ranks together H2D {gin:6.1f} GB/s, D2H {gout:6.1f} GB/s "
              f"-> {world * NQ / dt / 1e9:.2f} Gq/s if this were the count e2e step", flush=True)


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            if cmd[0] == "lscpu":
                out = "\n".join(l for l in out.splitlines() if "NUMA" in l or "Model name" in l or l.startswith("CPU(s)"))
            print(out, flush=True)
        except Exception as ex:
            print(cmd, "failed", ex)
    print(f"affinity of rank 0: {sorted(os.sched_getaffinity(0))}", flush=True)

plain_in, plain_out = host_alloc(IN_B, 1), host_alloc(OUT_B, 1)  # cudaHostAllocPortable
measure("pinned, H2D alone          ", plain_in, plain_out, True, False)
measure("pinned, D2H alone          ", plain_in, plain_out, False, True)
measure("pinned, H2D + D2H          ", plain_in, plain_out, True, True)
measure("pinned, H2D on 2 streams+D2H", plain_in, plain_out, True, True, in_streams=2)
measure("pinned, 8 MB pieces, both  ", plain_in, plain_out, True, True, piece=8 << 20)
wc_in = host_alloc(IN_B, 1 | 4)  # + cudaHostAllocWriteCombined
measure("write-combined in, H2D alone", wc_in, plain_out, True, False)
measure("write-combined in, both    ", wc_in, plain_out, True, True)
rt.cudaFreeHost(wc_in)
# NUMA placement: first touch from the cores of the other socket half (if the box has more than one node)
ncpu = os.cpu_count() or 1
try:
    nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
except Exception:
    nodes = []
if len(nodes) > 1:
    for nd in nodes:
        cpus = open(f"/sys/devices/system/node/{nd}/cpulist").read().strip()
        cpuset = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            cpuset.update(range(int(a), int(b or a) + 1))
        try:
            os.sched_setaffinity(0, cpuset & os.sched_getaffinity(0) or cpuset)
        except Exception as ex:
            if rank == 0:
                print(f"cannot bind to {nd}: {ex}")
            continue
        a_in, a_out = host_alloc(IN_B, 1), host_alloc(OUT_B, 1)
        measure(f"first touch on {nd}, both   ", a_in, a_out, True, True)
        rt.cudaFreeHost(a_in)
        rt.cudaFreeHost(a_out)
elif rank == 0:
    print(f"one NUMA node visible ({nodes}); no placement variant")
if world > 1:
    dist.destroy_process_group()
