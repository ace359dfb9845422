"""Probe (2+ GPUs, torchrun): is torch.distributed._symmetric_memory usable here without NVSHMEM, can OUR kernels read a peer's
buffer through the tensor `get_buffer` returns, how fast is a P2P pull of a gradient-sized buffer, what does a barrier cost."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm
    print(rank, "nvshmem:", symm.is_nvshmem_available(), flush=True)
    n = 22 * 1024 * 1024          # 88 MB of fp32: the size of the gradient set
    t = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok: world", hdl.world_size, "rank", hdl.rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (n,), torch.float32)
    print(rank, "peer value", float(pt[12345]), "device", pt.device, flush=True)
    # our own kernel reading peer memory
    from avid_cma_b200 import ops
    mine = torch.zeros(n, device=dev)
    ops.add_(mine, pt)
    torch.cuda.synchronize()
    assert float(mine[777]) == float(peer + 1), float(mine[777])
    # bandwidth of a P2P pull through a plain copy and through our kernel
    for name, fn in (("copy_", lambda: mine.copy_(pt)), ("avid_add_inplace", lambda: ops.add_(mine, pt))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(rank, name, "%.3f ms  %.1f GB/s over NVLink" % (ms, n * 4 / ms / 1e6), flush=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        hdl.barrier(channel=0)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "barrier %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), flush=True)
    # NCCL all-reduce of the same size for comparison
    x = torch.ones(n, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "nccl all_reduce 88 MB %.3f ms" % (e0.elapsed_time(e1) / 10), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
