"""Multi-GPU plumbing for batches of independent environments (mode 1).

Environments share nothing (the reference's ``Env`` objects are independent), so a batch of B
environments is block-partitioned over the ranks and each rank runs the single-GPU kernels on its
shard: there is NO collective on the data path.  The only exchange is optional and off the hot
path: gathering the per-environment (reward, num_agents) pairs, and the max-over-ranks of a timing.
One process per GPU, ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the environments owned by `rank`: contiguous blocks, the first
    ``n_envs % world_size`` ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_envs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_stats(reward_local: torch.Tensor, alive_local: torch.Tensor, n_envs: int):
    """All ranks get the global per-environment arrays (reward[n_envs] float64, alive[n_envs] int64),
    ordered by environment index.  Shards may be ragged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return reward_local.clone(), alive_local.clone()
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_envs, world, r) for r in range(world)]
    longest = max(b - a for a, b in sizes)
    pad_r = torch.zeros(longest, dtype=torch.float64, device=reward_local.device)
    pad_a = torch.zeros(longest, dtype=torch.int64, device=alive_local.device)
    n = reward_local.numel()
    pad_r[:n] = reward_local
    pad_a[:n] = alive_local
    out_r = [torch.empty_like(pad_r) for _ in range(world)]
    out_a = [torch.empty_like(pad_a) for _ in range(world)]
    dist.all_gather(out_r, pad_r)
    dist.all_gather(out_a, pad_a)
    reward = torch.cat([out_r[r][:b - a] for r, (a, b) in enumerate(sizes)])
    alive = torch.cat([out_a[r][:b - a] for r, (a, b) in enumerate(sizes)])
    return reward, alive


def max_over_ranks(value: float, device=None) -> float:
    """Device-side timings are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------
# host-side placement: one process per GPU, each on the cores (and memory) next to its GPU
# --------------------------------------------------------------------------------------------
def parse_cpulist(text: str):
    """'0-15,32-47' -> [0, ..., 15, 32, ..., 47] (the format of sysfs cpulist files)."""
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device_index: int):
    """The cores of the NUMA node GPU `device_index` hangs off (sysfs local_cpulist of its PCI device), or None."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        return cpus or None
    except Exception:
        return None


def split_cpus(cpus, sharers: int, position: int):
    """The `position`-th of `sharers` near-equal contiguous shares of `cpus` (never empty)."""
    n = len(cpus)
    if sharers <= 1 or n < sharers:
        return list(cpus)
    lo, hi = position * n // sharers, (position + 1) * n // sharers
    return list(cpus[lo:hi])


def bind_rank_to_gpu_cores(local_rank: int, local_world: int):
    """Pin this process to its share of the cores next to its GPU, so that its pinned host buffers are first-touched on
    that NUMA node and its host-side copies do not cross the socket interconnect (round 1 ran all eight ranks of the
    host-buffer path on NUMA node 0: 1.95x on 8 GPUs).  Ranks whose GPUs share a node split that node's cores.
    -> the cores bound to, or None if the topology is not readable (nothing changed)."""
    import os
    mine = gpu_local_cpus(local_rank)
    if mine is None:
        return None
    sharers = [r for r in range(local_world) if gpu_local_cpus(r) == mine]
    cpus = split_cpus(mine, len(sharers), sharers.index(local_rank) if local_rank in sharers else 0)
    try:
        os.sched_setaffinity(0, cpus)
        torch.set_num_threads(max(1, len(cpus)))
    except Exception:
        return None
    return cpus
