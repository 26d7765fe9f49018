"""Multi-process check of the slab decomposition over NVLink peer pointers:
    torchrun --nproc-per-node G tools/slab_check.py [H W steps]
Rank 0 builds a global state, every rank takes its slab; the sharded run must reproduce a single-GPU Env
(run on rank 0) bit for bit.  Prints 'SLAB CHECK OK' on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import die_b200 as D
    from die_b200 import data_init
    from die_b200.slab import SlabEnv, SlabPhysarumAgent, split_global_state

    phys = dict(scale=0.02, turn_angle=30, sense_offset=0.08)
    np.random.seed(5)
    medium = data_init.init_medium((H, W), 0.15, noise_seed=5)
    agents = data_init.agents_from_medium(medium)
    M = agents.shape[1]
    rng = np.random.default_rng(5)
    tr = np.radians(30)
    theta = rng.integers(-6, 6, M) * tr
    coins = rng.integers(0, 2, (steps, M))
    layout, mediums, locals_ = split_global_state(medium, agents, world)
    env = SlabEnv((H, W), D.Dynamics(), init_state=(layout, mediums[rank], locals_[rank]))
    ids = layout.global_ids(rank)
    agent = SlabPhysarumAgent(env, theta=theta[ids], **phys)

    ref_env = ref_agent = None
    if rank == 0:
        ref_env = D.Env((H, W), D.Dynamics(), init_state=(medium, agents))
        ref_agent = D.PhysarumAgent(max_agents=M, **phys)
        ref_agent.set_state(theta=theta)
        robs = ref_env._get_current_obs
    ok = True
    obs = env._get_current_obs
    for it in range(steps):
        action = agent.forward(obs, coin=coins[it][ids])
        obs, reward, _, _, info = env.step(action)
        # gather the sharded state on rank 0
        parts = [None] * world
        dist.gather_object((env.medium.cpu().numpy(), env.agents.cpu().numpy()[:, :len(ids)],
                            env.slab.theta.cpu().numpy()[:len(ids)]), parts if rank == 0 else None, dst=0)
        if rank == 0:
            ract = ref_agent.forward(robs, coin=coins[it])
            robs, rr, _, _, rinfo = ref_env.step(ract)
            med, ag = ref_env.get_state()
            gmed = np.concatenate([p[0] for p in parts], axis=1)
            gag, gth = np.zeros_like(ag), np.zeros(M)
            for q, p in enumerate(parts):
                gag[:, layout.global_ids(q)] = p[1]
                gth[layout.global_ids(q)] = p[2]
            same = (np.array_equal(gmed, med) and np.array_equal(gag, ag) and
                    np.array_equal(gth, ref_agent.get_state()[0]) and info['num_agents'] == rinfo['num_agents'] and
                    abs(reward - rr) <= 1e-11 * max(1.0, abs(rr)))
            if not same:
                ok = False
                print(f"step {it}: MISMATCH medium {np.array_equal(gmed, med)} agents {np.array_equal(gag, ag)} "
                      f"theta {np.array_equal(gth, ref_agent.get_state()[0])} reward {reward} vs {rr}", flush=True)
                break
    if rank == 0:
        print("SLAB CHECK OK" if ok else "SLAB CHECK FAILED", f"(G={world}, field {H}x{W}, {steps} steps, "
              f"bit-exact medium/agents/theta vs single-GPU Env)", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
