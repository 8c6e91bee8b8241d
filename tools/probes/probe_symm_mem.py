"""Probe: does torch.distributed._symmetric_memory work on this box (peer-mapped buffers + device barrier)?"""
import os, time
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    n = 50 * 1000 * 1000 // 4
    t = symm_mem.empty((n,), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
    print(rank, "rendezvous ok", type(hdl).__name__, [a for a in dir(hdl) if not a.startswith("_")][:40], flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (n,), torch.float32)
    print(rank, "peer value", float(pb[0]), float(pb[-1]), "ptrs", [hex(p) for p in hdl.buffer_ptrs][:8], flush=True)
    out = torch.empty_like(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        out.copy_(pb)
    e0.record()
    for _ in range(10):
        out.copy_(pb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(rank, "peer read %.1f MB in %.3f ms = %.1f GB/s" % (n * 4 / 1e6, ms, n * 4 / ms / 1e6), flush=True)
    e0.record()
    for _ in range(100):
        hdl.barrier(channel=1)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "barrier %.1f us" % (e0.elapsed_time(e1) * 10), flush=True)
    print(rank, "SYMM_MEM OK", flush=True)
except Exception as exc:
    import traceback
    traceback.print_exc()
    print(rank, "SYMM_MEM FAILED", type(exc).__name__, str(exc)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
