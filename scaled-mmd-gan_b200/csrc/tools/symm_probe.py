"""Does torch symmetric memory (peer pointers over NVLink) work on this box?  torchrun --nproc-per-node 2 symm_probe.py"""
import os, sys, time, traceback
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl")
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((1024, 256), dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024, 256), torch.bfloat16)
    v = float(peer[5, 7])
    print("rank %d symm ok: peer value %.1f buffer_ptrs=%s signal_pad_ptrs=%s" % (rank, v, [hex(p) for p in hdl.buffer_ptrs], [hex(p) for p in hdl.signal_pad_ptrs]), flush=True)
    hdl.barrier()
except Exception:
    print("rank %d symm FAILED" % rank); traceback.print_exc()
# plain CUDA IPC through torch's storage sharing
try:
    x = torch.full((1024,), float(rank + 10), device=dev)
    h = x.untyped_storage()._share_cuda_()
    objs = [None] * world
    dist.all_gather_object(objs, h)
    print("rank %d ipc handle fields: %d" % (rank, len(h)), flush=True)
    other = objs[(rank + 1) % world]
    st = torch.UntypedStorage._new_shared_cuda(*other)
    y = torch.empty(0, dtype=torch.float32, device=torch.device("cuda", other[0])).set_(st)
    torch.cuda.synchronize(); dist.barrier()
    print("rank %d ipc ok: device %s value %.1f ptr %s" % (rank, y.device, float(y[3]), hex(y.data_ptr())), flush=True)
    dist.barrier()
except Exception:
    print("rank %d ipc FAILED" % rank); traceback.print_exc()
dist.destroy_process_group()
