"""Per-rank, per-phase timing of the slab-decomposed step (CUDA events).  torchrun --nproc-per-node G tools/slab_profile.py N"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import die_b200 as D
    from die_b200.slab import SlabEnv, SlabPhysarumAgent
    phys = dict(scale=1.785 / (N - 1), turn_angle=30, sense_offset=10.2 / (N - 1))
    env = SlabEnv((N, N), D.Dynamics(init_agent_ratio=0.1), seed=3)
    agent = SlabPhysarumAgent(env, seed=11, **phys)
    s = env.slab
    for _ in range(10):
        env.step_async(agent.forward())
    torch.cuda.synchronize(); dist.barrier()
    names = ["forward", "move", "bar1", "field", "bar2", "corners", "feed", "bar3", "allreduce"]
    acc = np.zeros(len(names))
    for it in range(steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record(); agent.forward()
        ev[1].record(); s.phase_move()
        ev[2].record(); env.peers.barrier()
        ev[3].record(); s.phase_field()
        ev[4].record(); env.peers.barrier()
        ev[5].record(); s.phase_corners()
        ev[6].record(); s.phase_feed()
        ev[7].record(); env.peers.barrier()
        ev[8].record(); dist.all_reduce(s.stats)
        ev[9].record()
        torch.cuda.synchronize()
        cur = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
        acc += cur
        if rank in (0, world // 2) and it % 5 == 0:
            print(f"rank {rank} step {it + 10}: total {cur.sum():.3f} ms  " + " ".join(f"{n}={v:.2f}" for n, v in zip(names, cur)), flush=True)
    acc /= steps
    out = [None] * world
    dist.all_gather_object(out, (rank, env.layout.local_slots(rank), env.layout.n0[rank], acc.round(3).tolist()))
    if rank == 0:
        print("phase ms per rank:", names)
        for r in out: print(r)
    dist.barrier(); dist.destroy_process_group()

main()
