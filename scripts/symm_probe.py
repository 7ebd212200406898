"""Probe torch symmetric memory (peer-mapped buffers over NVLink) under torchrun."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm
buf = symm.empty((world, 4, 1080, 1920), dtype=torch.float32, device=f"cuda:{local}")
hdl = symm.rendezvous(buf, dist.group.WORLD)
peer0 = hdl.get_buffer(0, buf.shape, buf.dtype)
print(rank, "peer0 ptr", hex(peer0.data_ptr()), "local ptr", hex(buf.data_ptr()), flush=True)
peer0[rank].fill_(float(rank + 1))
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print("rank0 sees", [float(buf[r, 0, 0, 0]) for r in range(world)], flush=True)
# bandwidth of plain stores into the peer
x = torch.randn((4, 1080, 1920), device=f"cuda:{local}")
torch.cuda.synchronize(); dist.barrier()
t = time.perf_counter()
for _ in range(20): peer0[rank].copy_(x)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print(rank, "peer write GB/s %.1f" % (20 * x.numel() * 4 / dt / 1e9), flush=True)
dist.barrier(); dist.destroy_process_group()
