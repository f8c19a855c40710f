"""Probe (2+ GPUs, torchrun): which cross-process peer-memory mechanisms work on this box.
 (1) torch.distributed._symmetric_memory rendezvous (buffer_ptrs / signal pads / multicast)
 (2) legacy CUDA IPC through torch.multiprocessing.reductions (cudaIpcGetMemHandle)"""
import os, sys, time, traceback
import torch, torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(float(rank))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    say("symm_mem ok: ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support(dev.type, lr) if hasattr(hdl, "has_multicast_support") else None,
        hex(hdl.multicast_ptr) if hdl.multicast_ptr else 0, "sigpad", hdl.signal_pad_size)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    say("peer value", float(peer[0]))
    peer[:4] = 100.0 + rank                    # P2P store
    torch.cuda.synchronize(); hdl.barrier(channel=0)
    say("after peer store my buf", t[:5].tolist())
except Exception:
    say("symm_mem FAILED"); traceback.print_exc()
try:
    from torch.multiprocessing.reductions import reduce_tensor
    x = torch.full((1 << 20,), float(rank), device=dev)
    fn, args = reduce_tensor(x)
    objs = [None] * world
    dist.all_gather_object(objs, (fn, args))
    nxt = (rank + 1) % world
    f, a = objs[nxt]
    a = list(a); a[6] = lr                      # storage_device: map into MY device context
    px = f(*a)
    say("ipc ok: peer tensor device", px.device, "value", float(px[0]))
    px[:4] = 200.0 + rank
    torch.cuda.synchronize(); dist.barrier()
    say("after ipc peer store my x", x[:5].tolist())
except Exception:
    say("ipc FAILED"); traceback.print_exc()
dist.barrier()
dist.destroy_process_group()
