"""Probe: does torch symmetric memory (peer-mapped buffers over NVLink) rendezvous on this box?  torchrun --nproc-per-node 2"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
t = symm.empty(1024, dtype=torch.float64, device="cuda")
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal", [hex(p) for p in h.signal_pad_ptrs], "signal_pad_size", h.signal_pad_size, flush=True)
t.fill_(rank + 1)
dist.barrier()
peer = h.get_buffer((rank + 1) % world, (1024,), torch.float64)
print(rank, "peer value", float(peer[0]), "multicast", getattr(h, "multicast_ptr", None), flush=True)
dist.barrier()
dist.destroy_process_group()
