#!/usr/bin/env python
"""Times the z-slab halo exchange (ny_halo_exchange: NCCL send/recv of 3-plane faces) under torchrun.
Development tool: python -m torch.distributed.run --nproc-per-node N tools/halo_bench.py [ny nx]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nyles_b200 import lib, comm
    import ctypes as C
    ny, nx = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 1024)
    nz = 32
    L = lib.load()
    ctx = lib.context()
    cm = comm.get()
    below = rank - 1 if rank > 0 else -1
    above = rank + 1 if rank < world - 1 else -1
    for nf in (1, 3, 4):
        xs = [torch.randn((nz, ny, nx), dtype=torch.float64, device="cuda") for _ in range(nf)]
        ptrs = (C.c_void_p * nf)(*[x.data_ptr() for x in xs])

        def go():
            lib.check(L.ny_halo_exchange(ctx, cm, ptrs, nf, lib.ext(xs[0]), 3, below, above, 0, 0, lib.stream()))
        for _ in range(3):
            go()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        face = 3 * ny * nx * 8 * nf
        nb = (below >= 0) + (above >= 0)
        if rank == min(1, world - 1):
            print("fields %d face %.1f MB: %.3f ms per exchange (max over ranks), %.0f GB/s out per GPU (rank with %d neighbours)"
                  % (nf, face / 1e6, t.item(), nb * face / t.item() / 1e6, nb), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
