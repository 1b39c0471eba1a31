"""Does torch symmetric memory (peer pointers over NVLink, multicast) work on this box?"""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    s = peer[:4].tolist()
    print(f"rank {rank}: buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]} signal_pads {[hex(p) for p in hdl.signal_pad_ptrs]} "
          f"multicast {hex(hdl.multicast_ptr) if hdl.multicast_ptr else None} peer values {s} signal_pad_size {hdl.signal_pad_size}", flush=True)
    hdl.barrier()
except Exception as e:
    print(f"rank {rank}: symmetric memory FAILED: {type(e).__name__}: {e}", flush=True)
dist.barrier()
os._exit(0)
