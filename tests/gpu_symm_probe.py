"""Probe (GPU box, torchrun): does torch's symmetric memory give peer pointers on this box, for the default group and
for sub-groups, and what does a barrier cost? Prints one line per rank."""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
grp = dist.new_group(list(range(world)))
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
h = symm_mem.rendezvous(t, grp)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "sig", [hex(p) for p in h.signal_pad_ptrs], "pad", h.signal_pad_size,
      "multicast", h.has_multicast_support, flush=True)
t.fill_(float(rank))
h.barrier(0)
peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
peer[:16].fill_(100.0 + rank)      # P2P store into the next rank's buffer
h.barrier(0)
torch.cuda.synchronize()
print(rank, "after peer write:", t[:2].tolist(), t[16:18].tolist(), flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    h.barrier(0)
torch.cuda.synchronize()
e0.record()
for _ in range(100):
    h.barrier(0)
e1.record()
torch.cuda.synchronize()
print(rank, "barrier us:", e0.elapsed_time(e1) * 10, flush=True)
# p2p copy bandwidth: 256 MB into the peer
big = symm_mem.empty(64 << 20, dtype=torch.float32, device=dev)
hb = symm_mem.rendezvous(big, grp)
src = torch.empty(64 << 20, dtype=torch.float32, device=dev)
pb = hb.get_buffer((rank + 1) % world, (64 << 20,), torch.float32)
pb.copy_(src)
torch.cuda.synchronize()
hb.barrier(0)
e0.record()
for _ in range(5):
    pb.copy_(src)
e1.record()
torch.cuda.synchronize()
print(rank, "p2p copy GB/s:", 5 * 256e6 / (e0.elapsed_time(e1) * 1e-3) / 1e9 * 1.048576, flush=True)
dist.barrier()
dist.destroy_process_group()
