"""Probe: is torch symmetric memory (peer-mapped buffers over NVLink) usable on this box?  torchrun --nproc-per-node 2 tools/probe_symm.py"""
import os
import sys
import time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as sm
    t0 = time.time()
    buf = sm.empty(4096, dtype=torch.float32, device=dev)
    hdl = sm.rendezvous(buf, dist.group.WORLD)
    print(f"[rank {rank}] rendezvous ok in {time.time() - t0:.2f}s: rank {hdl.rank}/{hdl.world_size} buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]} "
          f"signal_pad_ptrs {[hex(p) for p in hdl.signal_pad_ptrs]} signal_pad_size {hdl.signal_pad_size} multicast {getattr(hdl, 'multicast_ptr', None)}", flush=True)
    buf.fill_(float(rank + 1))
    hdl.barrier()
    peer = (rank + 1) % world
    remote = hdl.get_buffer(peer, (4096,), torch.float32)
    got = float(remote[:8].sum())
    print(f"[rank {rank}] read peer {peer} through its mapping: sum of 8 = {got} (expect {8.0 * (peer + 1)})", flush=True)
    hdl.barrier()
    remote[100 + rank] = 1000.0 + rank                 # peer store
    hdl.barrier()
    torch.cuda.synchronize()
    print(f"[rank {rank}] local buffer after peer stores: {buf[100:100 + world].tolist()}", flush=True)
    print(f"[rank {rank}] SYMM_OK", flush=True)
except Exception as ex:
    import traceback
    traceback.print_exc()
    print(f"[rank {rank}] SYMM_FAILED {type(ex).__name__}: {ex}", flush=True)
dist.barrier()
dist.destroy_process_group()
