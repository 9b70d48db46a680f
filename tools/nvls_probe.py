import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank=int(os.environ["RANK"]); dev=torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
t=symm_mem.empty((1024,1024), dtype=torch.float32, device=dev); t.fill_(rank+1)
h=symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "multicast_ptr", hex(getattr(h,"multicast_ptr",0) or 0), "has_multicast", getattr(h, "multicast_ptr", None) not in (None,0), flush=True)
dist.barrier(); dist.destroy_process_group()
