"""Feasibility probe: torch symmetric memory (NVLink peer mapping) on this box."""
import os
import torch
import torch.distributed as dist

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty((4096,), dtype=torch.float64, device=torch.device('cuda', local))
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, 'rank/world', hdl.rank, hdl.world_size, 'buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs], 'signal pads',
      [hex(p) for p in hdl.signal_pad_ptrs][:2], 'multicast', hex(getattr(hdl, 'multicast_ptr', 0) or 0), flush=True)
t.fill_(rank + 1)
dist.barrier()
peer = hdl.get_buffer((rank + 1) % world, (4096,), torch.float64)
print(rank, 'peer value', float(peer[0]), flush=True)
dist.barrier()
dist.destroy_process_group()
