"""A short tour of every kernel and option combination on small inputs (device init, batches, every agent, food flow +
sense mask + 'constant' diffusion, speculative move, render, chunked host path, the cluster-fused step with its
distributed-shared-memory claims and bulk copies, the 128-bit field pass, pair mode, agents_die, float32 fields, the
graphed loop, a population of convolution policies).
Written to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py      (shared-memory hazards inside a CTA)
Where the tool is not usable it still serves as a crash / sticky-error check."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import die_b200 as D
from die_b200 import _lib
from die_b200.render import EnvRenderer

lib = _lib.load()
field = (37, 53)
for mode, flow, mask in (("wrap", False, False), ("constant", True, True), ("reflect", False, False)):
    dyn = D.Dynamics(init_agent_ratio=0.2, diffuse_mode=mode, apply_sense_mask=mask)
    if flow:
        dyn.op_food_flow = D.WaveSequence(field, dt=0.5, t_bounds=(0, 2)).get_flow_operator(0.5, 0.5)
    env = D.Env(field, dyn, init='device', seed=1, batch=3)
    m = env.max_agents
    for agent in (D.PhysarumAgent(max_agents=m, scale=0.02, sense_offset=0.06), D.BrownianAgent(0.03),
                  D.GradientAgent(max_agents=m, scale=0.02, sense_offset=0.05), D.ConstAgent((0.01, -0.02), 0.3)):
        if hasattr(agent, "fuse_move"):
            agent.fuse_move = (mode == "reflect")
        obs = env._get_current_obs
        for _ in range(4):
            obs, r, *_ = env.step(agent.forward(obs))
    env.render(host=True)
    hobs = tuple(t.cpu().numpy() for t in env._get_current_obs)
    _lib.check(lib.die_set_tuning(b"host_chunk_min_kb", 0))
    _lib.check(lib.die_set_tuning(b"host_chunks", 2))
    big = D.Env(field, D.Dynamics(init_agent_ratio=0.2), init='device', seed=2, batch=5)
    ag = D.PhysarumAgent(max_agents=m, scale=0.02, sense_offset=0.06)
    hobs = tuple(t.cpu().numpy() for t in big._get_current_obs)
    for _ in range(3):
        hobs, *_ = big.step(ag.forward(hobs))
    _lib.check(lib.die_set_tuning(b"host_chunk_min_kb", 32 << 10))
    _lib.check(lib.die_set_tuning(b"host_chunks", 4))
for key, batch in ((b"step_impl", 3), (b"field_vec", None), (b"pair_mode", 2)):     # the cluster-fused step, the 128-bit
    _lib.check(lib.die_set_tuning(key, 2 if key == b"pair_mode" else 1))            # field pass, pair mode forced
    env = D.Env((64, 40), D.Dynamics(), init='device', seed=3, batch=batch)
    ag = D.PhysarumAgent(max_agents=env.max_agents, scale=0.02, sense_offset=0.06)
    obs = env._get_current_obs
    for _ in range(4):
        obs, *_ = env.step(ag.forward(obs))
    _lib.check(lib.die_set_tuning(key, 1 if key == b"pair_mode" else 0))
env = D.Env((48, 64), D.Dynamics(agents_die=True, rate_feed=0.02), init='device', seed=4)        # lifecycle
ag = D.BrownianAgent(0.03, 2.0)
obs = env._get_current_obs
for _ in range(4):
    obs, *_ = env.step(ag.forward(obs))
env = D.Env((48, 64), D.Dynamics(), init='device', seed=5, field_dtype=torch.float32)            # float32 fields
ag = D.PhysarumAgent(max_agents=env.max_agents, scale=0.02, sense_offset=0.06)
obs = env._get_current_obs
for _ in range(4):
    obs, *_ = env.step(ag.forward(obs))
loop = D.GraphedLoop(D.Env((32, 32), D.Dynamics(), init='device', seed=6), D.BrownianAgent(0.02))       # CUDA graph replay
loop.run(6)
penv = D.Env((40, 48), D.Dynamics(food_infinite=True), init='device', seed=7, batch=4)                  # one model per env
pag = D.NeuralAutomataAgent(kernel_sizes=[3, 5], scale=0.01, deposit=2.0)
ev = D.PopulationEvaluator(penv, pag)
ev.evaluate(np.random.default_rng(0).uniform(-0.5, 0.5, (4, pag.model.num_parameters)).astype(np.float32), 5, reset=True)
torch.cuda.synchronize()
print("sanitize tour done")
