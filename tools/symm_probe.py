#!/usr/bin/env python
"""Does torch's symmetric memory (peer-mapped buffers + signal pads) work on this box?"""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    n = 64 * 1024 * 1024
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; multicast:", hdl.has_multicast_support, hex(hdl.multicast_ptr or 0),
          "ptrs", [hex(p) for p in hdl.buffer_ptrs][:3], flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    torch.cuda.synchronize()
    out = torch.empty_like(t)
    t0 = time.perf_counter()
    for _ in range(10):
        out.copy_(peer)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(rank, "peer value", float(out[0]), float(out[-1]), "copy_ from peer: %.1f GB/s" % (n * 4 / dt / 1e9), flush=True)
    hdl.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
