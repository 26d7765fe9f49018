"""Host-side cost of the drop-in loop on ONE small environment (the reference's own use case, 256x256):
   python tools/latency_small_env.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import die_b200 as D

env = D.Env((256, 256), D.Dynamics(init_agent_ratio=0.1))
for name, ag in (("physarum", D.PhysarumAgent(max_agents=65536, scale=0.007, turn_angle=30, sense_offset=0.04)),
                 ("brownian", D.BrownianAgent(move_scale=0.01))):
    obs = env._get_current_obs
    for _ in range(50):
        obs, *_ = env.step(ag.forward(obs))
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(1000):
        obs, r, term, trunc, info = env.step(ag.forward(obs))
    torch.cuda.synchronize(); sync_us = (time.perf_counter() - t) * 1e3
    t = time.perf_counter()
    for _ in range(1000):
        obs, _, _ = env.step_async(ag.forward(obs))
    torch.cuda.synchronize(); async_us = (time.perf_counter() - t) * 1e3
    t = time.perf_counter()
    for _ in range(1000):
        ag.forward(obs)
    torch.cuda.synchronize(); fwd_us = (time.perf_counter() - t) * 1e3
    print(f"{name}: forward + step (reward read back) {sync_us:.1f} us/iter; forward + step_async {async_us:.1f} us/iter; "
          f"forward alone {fwd_us:.1f} us/call")
