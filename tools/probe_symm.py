"""Probe: torch symmetric memory (NVLink peer pointers) on this box.  torchrun --nproc-per-node 2 tools/probe_symm.py"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float64, device=dev)
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "dev ptr table", hex(hdl.buffer_ptrs_dev), flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float64)
print(rank, "peer sum", float(peer.sum().item()), "expected", float(((rank + 1) % world + 1) * (1 << 20)), flush=True)
# bandwidth of a peer read
big = symm_mem.empty(1 << 27, dtype=torch.float64, device=dev)   # 1 GiB
h2 = symm_mem.rendezvous(big, dist.group.WORLD)
pb = h2.get_buffer((rank + 1) % world, (1 << 27,), torch.float64)
loc = torch.empty(1 << 27, dtype=torch.float64, device=dev)
h2.barrier(); torch.cuda.synchronize()
for _ in range(2): loc.copy_(pb)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): loc.copy_(pb)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(rank, "peer read GB/s", (1 << 30) / dt / 1e9, flush=True)
# barrier latency
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): hdl.barrier()
torch.cuda.synchronize(); print(rank, "symm barrier us", (time.perf_counter() - t0) / 100 * 1e6, flush=True)
t0 = time.perf_counter()
for _ in range(100): dist.barrier()
torch.cuda.synchronize(); print(rank, "nccl barrier us", (time.perf_counter() - t0) / 100 * 1e6, flush=True)
dist.destroy_process_group()
