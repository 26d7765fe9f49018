#!/usr/bin/env python
"""bench.py -- the hot path of gkirgizov/die on B200: agent.forward + env.step per iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl die_b200|reference]

Every N : 4096 independent 256x256 Physarum envs (README params) sharded over the ranks, no data-path
          collective (BASELINE.json configs[3], "sharded across 1/2/4/8 B200"); strong scaling.
          For N > 1 launch with torchrun, one rank per GPU.
N = 1   : additionally the single 4096x4096 Physarum field (BASELINE.json configs[2]) is measured in the
          same run and reported under "also" (its own value + roofline).
Rank 0 prints ONE JSON line.  `value` = cell-updates/s of the whole job with state resident in
HBM; `e2e` = the same loop through the host-buffer API (numpy obs/action cross PCIe every call);
`roofline` = the dominant kernel's algorithmic bytes / CUDA-event time against the measured HBM
peak; `cpu_baseline` = the numpy oracle (a port of the reference's CPU path) on this host.
`--impl reference` times that CPU path alone on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)      # reference README.md:45-48
AGENT_RATIO = 0.1
ARGS = None
METRIC = "env cell-updates/sec (agent.forward + env.step, Physarum)"
UNIT = "cell-updates/s"

# Algorithmic bytes per unit (float64 fields and agents, s_f = s_a = 8).  SURVEY.md 8d's model:
#   physarum_forward  R{x,y,theta} W{theta,dx,dy,dep} + gathers{4 chem, 1 food}        = 96 B / slot
#   move_claim        R{x,y,alive,dx,dy} W{x,y}                                          = 56 B / slot
#   agent_feed        R{agent_food,dep} W{agent_food} + gathers{food,occ}                = 40 B / slot
#   field_step        chem R+W, food R+W, occupancy R+W (+24 B per alive agent:         = 48 B / cell
#                     its deposit gather, chem add and occupancy write)
# = 240 B per cell-update (SURVEY_BYTES, reported as `survey_*`).  `roofline.achieved` is computed from the SMALLER
# figures the kernels that actually run need (round 1's verdict: 96 B/slot made the forward kernel look better than its
# DRAM utilisation):
#   physarum_forward  R{x,y,theta} W{theta,dx,dy,dep} + gathers{gradient pair, food}    = 72 B / slot (float32 pair; 80 B
#                     with the float64 cache): the gradient is published once per cell by the field pass
#   env_step_fused    the cluster-fused step: R{x,y,dx,dy,dep,agent_food} W{x,y,agent_food}, chem R+W, food R+W,
#                     occupancy W = 112 B / cell-update (alive from 1 bit, occupancy / claims / the food under a slot
#                     never leave the chip, dx / dy are read once from HBM)
# Implementation extras (4 B/slot cell cache, 8 B/cell published gradient, 4 B/cell claim table of the three-kernel path,
# the 8 B/slot cost hint the forward kernel leaves for the feed kernel -- which then reads it in place of `dep`, so
# agent_feed stays at 40 B) are NOT counted as achieved.
BYTES = {"physarum_forward": 72.0, "move_claim": 56.0, "agent_feed": 40.0, "field_step": 48.0,
         "env_step_fused": 112.0, "finalize_stats": 0.0}
SURVEY_BYTES = {"physarum_forward": 96.0, "move_claim": 56.0, "agent_feed": 40.0, "field_step": 48.0,
                "env_step_fused": 144.0, "finalize_stats": 0.0}
# fused step (default): the forward kernel also does move_claim's reads, cell resolution and claims; the
# positions (W{x,y}, 16 B) are committed by the feed kernel.  Same 240 B per cell-update in total.
BYTES_FUSED = {"physarum_forward": 72.0 + 40.0, "move_claim": 0.0, "agent_feed": 40.0 + 16.0, "field_step": 48.0,
               "finalize_stats": 0.0}
# committed move (`--fuse commit`, the run-loop contract of DIE_FWD_COMMIT_MOVE): the forward kernel is the move --
# R{x,y,theta} W{theta,dx,dy,dep,x,y} + gathers{gradient pair, food} = 88 B / slot; no move_claim launch; the feed
# kernel is the plain one.  72 + 16 + 40 + 48 = 176 B per cell-update: the 40 B that move_claim re-reads are gone.
BYTES_COMMIT = {"physarum_forward": 72.0 + 16.0, "move_claim": 0.0, "agent_feed": 40.0, "field_step": 48.0,
                "finalize_stats": 0.0}
ALIVE_EXTRA = {"field_step": 24.0, "env_step_fused": 24.0}
# float32 FIELD mode (Env(field_dtype=torch.float32): medium + consumed_field in float32, agents / actions float64), bytes
# counted per dtype (SURVEY 8d with s_f = 4, s_a = 8: 189 B per cell-update):
#   physarum_forward  R{x,y,theta} W{theta,dx,dy,dep} = 56 + gathers{gradient pair 8 (float32 x 2), food 4}   = 68 B / slot
#                     (survey: 56 + 5 gathers x 4 = 76)
#   move_claim        56 B / slot (no field access)
#   agent_feed        R{agent_food,dep} W{agent_food} = 24 + gathers{food, occ} 2 x 4                        = 32 B / slot
#   field_step        chem R+W, food R+W, occupancy R+W = 6 x 4                                             = 24 B / cell
BYTES_F32 = {"physarum_forward": 68.0, "move_claim": 56.0, "agent_feed": 32.0, "field_step": 24.0, "finalize_stats": 0.0}
SURVEY_BYTES_F32 = {"physarum_forward": 76.0, "move_claim": 56.0, "agent_feed": 32.0, "field_step": 24.0, "finalize_stats": 0.0}
ALIVE_EXTRA_F32 = {"field_step": 12.0}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except Exception:
        return None


def ncu_traffic_note():
    """Provenance remark of that capture (e.g. "taken before change X"), if the file carries one."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get("_note")
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def reader():
            for line in self.proc.stdout:
                self.samples.append(line.strip())
        self.thread = threading.Thread(target=reader, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------
def build_host_state(field, n_distinct, seed):
    """`n_distinct` seeded initial states (food = gradient noise with 256^2-like texture, i.e.
    8 lattice periods per 256 cells; occupancy Bernoulli(0.1); chem = 0)."""
    from die_b200 import data_init
    periods = max(8, 8 * field[0] // 256)
    mediums, agents = [], []
    for k in range(n_distinct):
        np.random.seed(seed + k)
        m = data_init.init_medium(field, AGENT_RATIO, noise_seed=seed + k, periods=periods)
        mediums.append(m)
        agents.append(data_init.agents_from_medium(m))
    return np.stack(mediums), np.stack(agents)


def lattice_theta_device(B, M, turn_angle, seed, device):
    """theta_0 uniform on the turn lattice (what discretize(angle(N(0,.4)^2)) gives), drawn on device."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    tr = float(np.radians(turn_angle))
    nlat = int(round(2 * np.pi / tr))
    k = torch.randint(-nlat // 2, nlat - nlat // 2, (B, M), generator=g, device=device)
    return k.to(torch.float64) * tr


def make_env_and_agent(D, torch, field, B_local, batched, device, rank, seed_base, field_dtype=None, fuse=None):
    """Seeded synthetic state: up to 16 distinct host-built environments, tiled on the device."""
    n_distinct = 1 if not batched else min(B_local, 16)
    med_h, ag_h = build_host_state(field, n_distinct, seed=seed_base + 1000 * rank)
    med = torch.from_numpy(med_h).to(device)
    ag = torch.from_numpy(ag_h).to(device)
    if batched and B_local > n_distinct:
        reps = (B_local + n_distinct - 1) // n_distinct
        med = med.repeat(reps, 1, 1, 1)[:B_local]
        ag = ag.repeat(reps, 1, 1)[:B_local]
    alive = int((ag[:, 2] > 0).sum().item())
    env = D.Env(field, D.Dynamics(init_agent_ratio=AGENT_RATIO), batch=(B_local if batched else None),
                init_state=(med, ag), device=device, field_dtype=field_dtype or torch.float64)
    del med, ag
    M = env.max_agents
    agent = D.PhysarumAgent(max_agents=M, rng="philox", seed=1234 + rank, **PHYS)
    agent._theta = lattice_theta_device(B_local, M, PHYS["turn_angle"], 77 + rank, device)
    agent.fuse_move = {'': False, 'spec': True, 'commit': 'commit'}[(ARGS.fuse if fuse is None else fuse) or '']
    return env, agent, alive


def measure(D, torch, dist, args, field, B_local, batched, device, rank, world, want_clocks, field_dtype=None, fuse=None):
    """Warm-up, the timed region (CUDA events, barrier + synchronize on both sides, max over
    ranks), then the per-kernel breakdown.  Returns a dict of raw measurements."""
    env, agent, alive_local = make_env_and_agent(D, torch, field, B_local, batched, device, rank, 0, field_dtype, fuse)
    M, C = env.max_agents, field[0] * field[1]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def loop(n, obs):
        for _ in range(n):
            action = agent.forward(obs)
            obs, _, _ = env.step_async(action)
        return obs

    obs = loop(args.warmup, env._get_current_obs)
    sync_all()
    sampler = ClockSampler(device.index)
    if want_clocks:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    obs = loop(args.steps, obs)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if want_clocks else None
    from die_b200.sharding import max_over_ranks
    ms = max_over_ranks(ms, device)                         # device time, max over ranks
    ms_per_step = ms / args.steps

    # per-kernel breakdown: CUDA events on the launching stream between the kernels
    from die_b200 import _lib as dlib
    fused0 = dlib.load().die_get_counter(b"step_fused")
    env.set_profiling(True)
    fwd_events = []
    n_prof = min(args.steps, 200)
    for _ in range(n_prof):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        action = agent.forward(obs)
        e1.record()
        obs, _, _ = env.step_async(action)
        fwd_events.append((e0, e1))
    torch.cuda.synchronize()
    kms, nprof = env.kernel_times()
    env.set_profiling(False)
    kernel_ms = {"physarum_forward": sum(a.elapsed_time(b) for a, b in fwd_events) / n_prof}
    if dlib.load().die_get_counter(b"step_fused") - fused0 >= n_prof:
        # the cluster-fused step: one launch, recorded as the first interval of the per-kernel breakdown
        kernel_ms["env_step_fused"] = sum(kms.values()) / max(nprof, 1)
    else:
        kernel_ms.update({k: v / max(nprof, 1) for k, v in kms.items()})
    grad_kind = dlib.load().die_env_gradient_kind(env._handle)
    return dict(env=env, agent=agent, obs=obs, steps_done=args.warmup + args.steps + n_prof,
                ms_per_step=ms_per_step, kernel_ms=kernel_ms, clocks=clocks,
                alive_local=alive_local, M=M, C=C, fused=bool(env.last_step_fused), grad_kind=int(grad_kind),
                committed=bool(getattr(agent, 'last_committed', False)),
                f32=(field_dtype is not None and field_dtype == torch.float32))


def steady_state_leg(torch, meas, checkpoints, window=40):
    """ms per step in a `window`-step window ending at each of `checkpoints` total steps (the ~90 % ghost slots start at
    (0, 0) and random-walk outwards: the gathers of the first few hundred steps hit L1 / L2 more often than later)."""
    env, agent, obs, done = meas["env"], meas["agent"], meas["obs"], meas["steps_done"]
    out = {}
    for cp in checkpoints:
        if cp < done + window:
            continue
        for _ in range(cp - window - done):
            obs, _, _ = env.step_async(agent.forward(obs))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(window):
            obs, _, _ = env.step_async(agent.forward(obs))
        e1.record()
        torch.cuda.synchronize()
        done = cp
        out[str(cp)] = round(e0.elapsed_time(e1) / window, 5)
    kernels = None
    if out:                                              # which kernel pays: the per-kernel breakdown at the last checkpoint
        env.set_profiling(True)
        fwd = []
        for _ in range(window):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            action = agent.forward(obs)
            a1.record()
            obs, _, _ = env.step_async(action)
            fwd.append((a0, a1))
        torch.cuda.synchronize()
        kms, nprof = env.kernel_times()
        env.set_profiling(False)
        done += window
        kernels = {"physarum_forward": round(sum(a.elapsed_time(b) for a, b in fwd) / window, 5)}
        kernels.update({k: round(v / max(nprof, 1), 5) for k, v in kms.items()})
    meas["obs"], meas["steps_done"] = obs, done
    return out, kernels


def small_env_leg(D, torch, device, agent_name, iters=300, warmup=20, cpu_iters=20):
    """BASELINE configs[0] / [1]: ONE 256x256 environment, the reference's own loop (examples/minimal_run.py:21-25,
    README.md:23-48) -- action = agent.forward(obs); obs, reward, _, _, info = env.step(action) with the reward read back
    every iteration -- timed by the wall clock around `iters` iterations; next to the oracle on one host core."""
    import die_b200 as Dm
    field = (256, 256)
    med_h, ag_h = build_host_state(field, 1, seed=4242)
    def make_agent(mod):
        if agent_name == "brownian":
            return mod.BrownianAgent(move_scale=0.01)
        return mod.PhysarumAgent(max_agents=field[0] * field[1], **PHYS)
    env = D.Env(field, D.Dynamics(init_agent_ratio=AGENT_RATIO), init_state=(med_h[0], ag_h[0]), device=device)
    agent = make_agent(D)
    obs = env._get_current_obs
    for _ in range(warmup):
        obs, reward, _, _, info = env.step(agent.forward(obs))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        obs, reward, _, _, info = env.step(agent.forward(obs))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"field": list(field), "agent": agent_name, "iters": iters, "ms_per_iter": dt / iters * 1e3,
           "value": field[0] * field[1] * iters / dt, "unit": UNIT, "num_agents": int(info["num_agents"]),
           "api": "agent.forward(obs) + env.step(action) -> (obs, reward: float, terminated, truncated, info), one "
                  "host synchronisation per iteration (the reward read-back)"}
    graph = getattr(Dm, "GraphedLoop", None)
    if graph is not None:
        # the same iteration captured in a CUDA graph (die_b200.GraphedLoop): forward + step_async, reward kept on device
        try:
            loop = graph(env, agent)
            loop.run(warmup)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loop.run(iters)
            torch.cuda.synchronize()
            dg = time.perf_counter() - t0
            out["cuda_graph"] = {"ms_per_iter": dg / iters * 1e3, "value": field[0] * field[1] * iters / dg, "unit": UNIT,
                                 "api": "die_b200.GraphedLoop(env, agent).run(n): forward + step_async replayed from one "
                                        "captured graph, rewards accumulated on the device"}
        except Exception as exc:                     # never lose the whole bench line to the optional leg
            out["cuda_graph"] = {"error": repr(exc)}
        if agent_name == "physarum" and "error" not in out["cuda_graph"]:
            # the captured loop IS the run-loop contract of the committed move: 4 launches per iteration instead of 5
            try:
                agent.fuse_move = "commit"
                loop = graph(env, agent)
                loop.run(warmup)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                loop.run(iters)
                torch.cuda.synchronize()
                dg = time.perf_counter() - t0
                out["cuda_graph_committed_move"] = {"ms_per_iter": dg / iters * 1e3, "value": field[0] * field[1] * iters / dg,
                                                    "unit": UNIT, "committed": bool(agent.last_committed),
                                                    "api": "the same with agent.fuse_move = 'commit'"}
            except Exception as exc:
                out["cuda_graph_committed_move"] = {"error": repr(exc)}
            agent.fuse_move = False
    if cpu_iters > 0:
        from oracle import die_ref as R
        np.random.seed(1)
        renv = R.Env(field, R.Dynamics(init_agent_ratio=AGENT_RATIO), noise_seed=1)
        ragent = R.BrownianAgent(0.01) if agent_name == "brownian" else R.PhysarumAgent(max_agents=field[0] * field[1], **PHYS)
        robs = renv._get_current_obs
        for _ in range(2):
            robs, *_ = renv.step(ragent.forward(robs))
        t0 = time.perf_counter()
        for _ in range(cpu_iters):
            robs, *_ = renv.step(ragent.forward(robs))
        dc = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": field[0] * field[1] * cpu_iters / dc, "unit": UNIT, "cores": 1, "kind": "port",
                               "ms_per_iter": dc / cpu_iters * 1e3,
                               "sample": f"{cpu_iters} iterations of the same loop by the numpy oracle on one host core"}
    return out


def host_pinned_budget_envs(field, M, frac=0.35):
    """How many environments' host buffers (pinned: obs x2 + action + agents) fit in `frac` of the available host memory."""
    C = field[0] * field[1]
    per_env = 8 * (2 * 3 * C + 4 * M + 3 * M) + 8 * (4 * M + 3 * C)      # env + agent pinned buffers, first host copy of obs
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    return max(1, int(avail * frac // per_env))


def e2e_leg(D, torch, dist, args, field, B_local, batched, device, rank, world, M, C, n_gpus):
    """The metric measured end to end through the public API with HOST buffers: numpy observation in, numpy action out of
    Agent.forward; numpy action in, numpy observation + reward out of Env.step; every byte crosses PCIe inside the timed
    region.  On the FULL per-GPU batch (bounded only by pinned host memory; `envs_per_gpu` says what ran).
    The two calls are synchronous, so one caller thread keeps only one PCIe direction busy at a time (forward uploads
    4 doubles per slot -- the channels the policy reads -- and downloads 3, step downloads 9 or 10): `value` is therefore measured with `workers` caller threads,
    each stepping its own share of the environments through the same two calls; the single-thread figure is reported
    beside it."""
    from die_b200.sharding import max_over_ranks
    B_e2e = B_local if batched else 1
    if args.e2e_envs > 0:
        B_e2e = min(B_e2e, args.e2e_envs)
    B_e2e = max(1, min(B_e2e, host_pinned_budget_envs(field, M)))
    k_e2e = max(3, min(args.steps, args.e2e_steps))
    W = max(1, min(args.e2e_workers, B_e2e))
    out = {"unit": UNIT, "steps": k_e2e, "envs_per_gpu": B_e2e,
           "api": "numpy obs/action across Agent.forward (die_gradient_forward_host) and Env.step ("
                  "die_env_step_host_flags: the action is not uploaded again when it is the array forward just returned; the "
                  "observation channels no kernel reads (forward) or writes (step: alive) stay where they are), "
                  "each call cut into chunks on two streams; pinned host memory"}
    # -- one caller thread
    env, agent, _ = make_env_and_agent(D, torch, field, B_e2e, batched, device, rank, 50)
    hobs = tuple(t.cpu().numpy() for t in env._get_current_obs)
    for _ in range(2):                                   # warm the pinned staging buffers
        hobs, *_ = env.step(agent.forward(hobs))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_start = time.perf_counter()
    for _ in range(k_e2e):
        hact = agent.forward(hobs)                       # H2D obs, kernel, D2H action
        hobs, hr, _, _, _ = env.step(hact)               # (H2D action,) kernels, D2H obs + reward
    torch.cuda.synchronize()
    t_one = max_over_ranks(time.perf_counter() - t_start, device)
    h2d_step, d2h_step = env.host_io_bytes_per_step()
    # forward uploads the channels the policy reads (x, y; env_food, chem1) and downloads the action
    fwd_h2d, fwd_d2h = agent.host_io_bytes_per_forward(B_e2e, M, C)
    h2d, d2h = (h2d_step + fwd_h2d) * n_gpus, (d2h_step + fwd_d2h) * n_gpus
    one = {"value": C * B_e2e * n_gpus * k_e2e / t_one, "ms_per_step": t_one / k_e2e * 1e3,
           "h2d_gbs_per_gpu": round(h2d / n_gpus / (t_one / k_e2e) / 1e9, 1),
           "d2h_gbs_per_gpu": round(d2h / n_gpus / (t_one / k_e2e) / 1e9, 1),
           "action_reupload_skipped": bool(getattr(env, "last_step_reused_device_action", False))}
    out.update({"h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h})
    del env, agent, hobs, hact
    torch.cuda.empty_cache()
    if W > 1 and batched:
        wk = e2e_workers_leg(D, torch, dist, args, field, B_e2e, device, rank, world, k_e2e, C, n_gpus, W)
        if "error" not in wk:
            t_w = wk["ms_per_step"] * 1e-3
            out.update({"value": wk["value"], "ms_per_step": wk["ms_per_step"], "workers": W,
                        "envs_per_gpu": wk["envs_per_worker"] * W,
                        "h2d_bytes_per_step": h2d * wk["envs_per_worker"] * W // B_e2e,
                        "d2h_bytes_per_step": d2h * wk["envs_per_worker"] * W // B_e2e,
                        "h2d_gbs_per_gpu": round(h2d / n_gpus * wk["envs_per_worker"] * W / B_e2e / t_w / 1e9, 1),
                        "d2h_gbs_per_gpu": round(d2h / n_gpus * wk["envs_per_worker"] * W / B_e2e / t_w / 1e9, 1),
                        "one_caller_thread": one})
            return out
        out["workers_error"] = wk["error"]
    out.update({"value": one["value"], "ms_per_step": one["ms_per_step"], "workers": 1,
                "h2d_gbs_per_gpu": one["h2d_gbs_per_gpu"], "d2h_gbs_per_gpu": one["d2h_gbs_per_gpu"],
                "action_reupload_skipped": one["action_reupload_skipped"]})
    return out


def e2e_workers_leg(D, torch, dist, args, field, B_e2e, device, rank, world, k_e2e, C, n_gpus, W):
    """The host-buffer loop driven by W host threads, each with its own Env / Agent over B_e2e / W environments and its
    own CUDA stream.  One thread's Agent.forward (upload-heavy: the observation) then overlaps another's Env.step
    (download-heavy), so both PCIe directions carry data all the time instead of taking turns.  Same public calls,
    same bytes per environment (the library calls release the GIL)."""
    import threading
    per = B_e2e // W
    pairs = [make_env_and_agent(D, torch, field, per, True, device, rank, 60 + w)[:2] for w in range(W)]
    streams = [torch.cuda.Stream(device=device) for _ in range(W)]
    start = threading.Barrier(W + 1)
    errors = []

    def worker(w):
        env, agent = pairs[w]
        try:
            torch.cuda.set_device(device)
            with torch.cuda.stream(streams[w]):
                hobs = tuple(t.cpu().numpy() for t in env._get_current_obs)
                for _ in range(2):
                    hobs, *_ = env.step(agent.forward(hobs))
                start.wait()
                for _ in range(k_e2e):
                    hobs, *_ = env.step(agent.forward(hobs))
                streams[w].synchronize()
        except Exception as exc:                # report, never hang the barrier
            errors.append(repr(exc))
            start.abort()

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(W)]
    for t in threads:
        t.start()
    try:
        start.wait()
    except threading.BrokenBarrierError:
        pass
    t0 = time.perf_counter()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    if errors:
        return {"error": errors[0]}
    from die_b200.sharding import max_over_ranks
    dt = max_over_ranks(dt, device)
    return {"workers": W, "envs_per_worker": per, "value": C * per * W * n_gpus * k_e2e / dt, "unit": UNIT,
            "ms_per_step": dt / k_e2e * 1e3}


def roofline_of(meas, B_local, wl_name):
    peak, peak_src = measured_hbm_peak()
    M, C, alive_local = meas["M"], meas["C"], meas["alive_local"]
    slots_local, cells_local = M * B_local, C * B_local
    kernels = {}
    BY = dict(BYTES_COMMIT if meas.get("committed") else (BYTES_FUSED if meas.get("fused") else BYTES))
    SV, AX = SURVEY_BYTES, ALIVE_EXTRA
    if meas.get("f32"):
        BY, SV, AX = dict(BYTES_F32), SURVEY_BYTES_F32, ALIVE_EXTRA_F32
    if meas.get("grad_kind") == 1:                 # float64 gradient cache: a 16-byte pair per slot
        BY["physarum_forward"] += 8.0
    for k, t_ms in meas["kernel_ms"].items():
        units = cells_local if k in ("field_step", "env_step_fused") else slots_local
        nbytes = BY[k] * units + AX.get(k, 0.0) * alive_local
        gbs = nbytes / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        kernels[k] = {"ms": round(t_ms, 5), "algorithmic_bytes": nbytes, "gbs": round(gbs, 1),
                      "frac": round(gbs / peak, 4),
                      "survey_bytes": SV[k] * units + AX.get(k, 0.0) * alive_local}
    dominant = max((k for k in kernels if BY[k] > 0), key=lambda k: kernels[k]["ms"])
    step_bytes = sum(v["algorithmic_bytes"] for v in kernels.values())
    survey_bytes = sum(v["survey_bytes"] for v in kernels.values())
    step_gbs = step_bytes / (meas["ms_per_step"] * 1e-3) / 1e9
    survey_gbs = survey_bytes / (meas["ms_per_step"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": dominant, "achieved": kernels[dominant]["gbs"], "peak": peak,
            "unit": "GB/s", "frac": kernels[dominant]["frac"], "peak_source": peak_src,
            "traffic": ncu_traffic(wl_name, dominant), "traffic_note": ncu_traffic_note(), "kernels": kernels,
            "step": {"algorithmic_bytes": step_bytes, "gbs": round(step_gbs, 1),
                     "frac": round(step_gbs / peak, 4), "frac_of_8TBs_nominal": round(step_gbs / 8000.0, 4),
                     "survey_bytes": survey_bytes, "survey_gbs": round(survey_gbs, 1),
                     "survey_frac": round(survey_gbs / peak, 4),
                     "survey_frac_of_8TBs_nominal": round(survey_gbs / 8000.0, 4)}}, step_bytes


def measure_slab(args, torch, dist, device, rank, world, N, steps, warmup):
    """BASELINE configs[4]: ONE field split into row slabs over the ranks, NVLink peer access in-kernel
    (die_b200/slab.py).  Agent parameters are the README ones expressed in cells: the
    256^2 geometry (1.785 cells per step, 10.2 cells look-ahead) at any field size.  -> the JSON line (rank 0) or None."""
    import die_b200 as D
    from die_b200.slab import SlabEnv, SlabPhysarumAgent
    from die_b200.sharding import max_over_ranks
    phys = dict(scale=1.785 / (N - 1), turn_angle=30, sense_offset=10.2 / (N - 1))
    t0 = time.time()
    env = SlabEnv((N, N), D.Dynamics(init_agent_ratio=AGENT_RATIO), seed=3, corner_r=args.corner_r)
    agent = SlabPhysarumAgent(env, seed=11, **phys)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    obs = env._get_current_obs
    for _ in range(warmup):
        obs, _ = env.step_async(agent.forward(obs))
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(device.index)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    mode = os.environ.get("DIE_SLAB_MODE", "reduce")          # diagnosis: reduce | noreduce | sync
    for _ in range(steps):
        obs, stats = env.step_async(agent.forward(obs), reduce_stats=(mode != "noreduce"))
        if mode == "sync":
            torch.cuda.synchronize()
    if mode == "noreduce":
        dist.all_reduce(stats)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(ev0.elapsed_time(ev1), device) / steps
    clocks = sampler.stop() if rank == 0 else None
    reward, alive = float(stats[0].item()), int(round(float(stats[1].item())))
    peak, peak_src = measured_hbm_peak()
    C = N * N
    step_bytes = 240.0 * C + 24.0 * alive
    del env, agent, obs
    torch.cuda.empty_cache()
    if rank == 0:
        gbs = step_bytes / (ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": C / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic (device-generated gradient-noise food, Bernoulli(0.1) agents)",
                "config": {"workload": f"physarum_single_field_{N}x{N}_slab", "field": [N, N], "agent": "PhysarumAgent",
                           **phys, "agent_ratio": AGENT_RATIO, "max_agents": C, "alive_agents": alive,
                           "decomposition": f"{world} row slabs, symmetric memory, in-kernel NVLink peer loads/atomics, "
                                            f"corner mirror of {args.corner_r} cells, "
                                            "3 barriers + one 2-double all-reduce per step",
                           "l2": "per-step working set >> 126 MB L2 (no flush)"},
                "agent_steps_per_s": C / (ms * 1e-3), "alive_agent_steps_per_s": alive / (ms * 1e-3),
                "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": round(gbs / world, 1), "peak": peak,
                             "unit": "GB/s per GPU", "frac": round(gbs / world / peak, 4), "peak_source": peak_src,
                             "traffic": None},
                "cpu_baseline": None, "e2e": None, "gpu_launches": steps * (10 if args.corner_r else 9), "launches_per_step": 10 if args.corner_r else 9,
                "clocks": clocks, "setup_s": round(setup_s, 1), "last_reward": reward}
        return line
    return None


def run_slab(args, torch, dist, device, rank, world):
    line = measure_slab(args, torch, dist, device, rank, world, args.field, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(line))


def run_die_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (die_b200 has no CPU fallback); "
                         "use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    bound_cpus = None
    if world > 1:
        from die_b200.sharding import bind_rank_to_gpu_cores
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        bound_cpus = bind_rank_to_gpu_cores(local_rank, local_world)     # cores (and first-touch memory) next to this GPU
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world
    if args.gpus != n_gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    import die_b200 as D
    from die_b200 import _lib as dlib
    for kv in args.tune:
        key, val = kv.split("=")
        dlib.check(dlib.load().die_set_tuning(key.encode(), int(val)))
    l2_prev = dlib.C.c_int32(0)
    dlib.check(dlib.load().die_device_l2_fetch_granularity(int(args.l2_fetch), dlib.C.byref(l2_prev)))

    if args.workload == "slab":
        if world < 2:
            raise SystemExit("--workload slab needs torchrun with >= 2 ranks")
        run_slab(args, torch, dist, device, rank, world)
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- headline workload: 4096 independent 256x256 Physarum envs sharded over the ranks ----------
    workload = args.workload
    if workload == "auto":
        workload = "batch256"
    t0 = time.time()
    if workload == "field4096":
        field, B_total, batched, scaling = (args.field, args.field), 1, False, "weak"
        B_local = 1
        wl_name = "physarum_single_field_%dx%d" % field
    else:
        field, B_total, batched, scaling = (256, 256), args.batch, True, "strong"
        from die_b200.sharding import shard_range
        lo, hi = shard_range(B_total, n_gpus, rank)          # block partition, no data-path collective
        B_local = hi - lo
        wl_name = f"physarum_batched_{B_total}x256x256"
    meas = measure(D, torch, dist, args, field, B_local, batched, device, rank, world, want_clocks=(rank == 0))
    setup_s = time.time() - t0
    M, C = meas["M"], meas["C"]
    ms_per_step = meas["ms_per_step"]
    cells_per_step = C * B_total
    value = cells_per_step / (ms_per_step * 1e-3)
    roofline, step_bytes = roofline_of(meas, B_local, wl_name)
    alive_local = meas["alive_local"]
    del meas["env"], meas["agent"], meas["obs"]
    torch.cuda.empty_cache()

    # ---- N = 1 only: the single 4096x4096 field (BASELINE.json configs[2]) in the same run, followed to its steady
    #      state, and the reference's own configurations (configs[0], [1]: ONE 256x256 environment, 300 iterations) ----
    also = None
    if n_gpus == 1 and workload == "batch256" and not args.no_single_field:
        f2 = (args.field, args.field)
        m2 = measure(D, torch, dist, args, f2, 1, False, device, rank, world, want_clocks=False)
        r2, sb2 = roofline_of(m2, 1, "physarum_single_field_%dx%d" % f2)
        steady, steady_kernels = steady_state_leg(torch, m2, [int(c) for c in args.steady.split(",") if c]) \
            if args.steady else ({}, None)
        C2 = f2[0] * f2[1]
        entry = {"value": C2 / (m2["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": m2["ms_per_step"],
                 "measured_at_steps": [args.warmup, args.warmup + args.steps],
                 "max_agents": m2["M"], "alive_agents": m2["alive_local"], "roofline": r2}
        if steady:
            last = steady[max(steady, key=int)]
            step_b = r2["step"]["algorithmic_bytes"]
            entry["steady_state"] = {
                "ms_per_step_in_the_40_steps_before_step": steady, "kernel_ms_after_last": steady_kernels,
                "value_at_last": C2 / (last * 1e-3), "unit": UNIT,
                "frac_at_last": round(step_b / (last * 1e-3) / 1e9 / r2["peak"], 4),
                "frac_of_8TBs_nominal_at_last": round(step_b / (last * 1e-3) / 1e9 / 8000.0, 4),
                "survey_frac_of_8TBs_nominal_at_last": round(r2["step"]["survey_bytes"] / (last * 1e-3) / 1e9 / 8000.0, 4)}
        also = {"physarum_single_field_%dx%d" % f2: entry}
        del m2
        torch.cuda.empty_cache()
        if not args.no_small_env:
            also["brownian_single_env_256x256_300_iters"] = small_env_leg(D, torch, device, "brownian",
                                                                         cpu_iters=0 if args.no_cpu else 20)
            also["physarum_single_env_256x256_300_iters"] = small_env_leg(D, torch, device, "physarum",
                                                                         cpu_iters=0 if args.no_cpu else 20)

    # ---- N = 1 only: the float32 FIELD mode of both workloads (fields in float32, agents float64; bytes counted per dtype;
    #      float64 stays the headline dtype: it is the reference's) ---------------------------------------------------
    if n_gpus == 1 and workload == "batch256" and not args.no_f32:
        for f3, B3, batched3, name3 in (((256, 256), B_local, True, wl_name), ((args.field, args.field), 1, False,
                                                                              "physarum_single_field_%dx%d" % (args.field, args.field))):
            if not batched3 and args.no_single_field:
                continue
            m3 = measure(D, torch, dist, args, f3, B3, batched3, device, rank, world, want_clocks=False,
                         field_dtype=torch.float32)
            r3, _ = roofline_of(m3, B3, name3 + "_f32_fields")
            also = dict(also or {})
            also[name3 + "_f32_fields"] = {
                "value": f3[0] * f3[1] * B3 / (m3["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": m3["ms_per_step"],
                "dtype": "f32 fields (medium, consumed_field), f64 agents / actions / headings", "roofline": r3}
            del m3
            torch.cuda.empty_cache()

    # ---- N = 1 only: the same workload under the run-loop contract (agent.fuse_move = 'commit', DIE_FWD_COMMIT_MOVE): the
    #      forward launch also moves the agents, Env.step is the field pass + the feed kernel -- bit-identical, opt-in,
    #      NOT the headline (the headline keeps forward and step independent calls, as any caller of the reference may) ---
    if n_gpus == 1 and workload == "batch256" and not ARGS.fuse and not args.no_commit:
        m4 = measure(D, torch, dist, args, field, B_local, batched, device, rank, world, want_clocks=False, fuse="commit")
        r4, _ = roofline_of(m4, B_local, wl_name + "_committed_move")
        also = dict(also or {})
        also[wl_name + "_committed_move"] = {
            "value": C * B_local / (m4["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": m4["ms_per_step"],
            "committed": m4["committed"], "launches_per_step": 4,
            "api": "agent.fuse_move = 'commit': `action = agent.forward(obs); obs, ... = env.step_async(action)` with nothing "
                   "in between (the loop of examples/minimal_run.py:21-25); any other continuation raises",
            "roofline": r4}
        del m4
        torch.cuda.empty_cache()

    # ---- N > 1: the single 32768x32768 field split into row slabs over the ranks (BASELINE.json configs[4]) ----------
    if n_gpus > 1 and workload == "batch256" and not args.no_slab:
        try:
            sl = measure_slab(args, torch, dist, device, rank, world, args.slab_field, args.slab_steps, 10)
        except Exception as exc:                     # (an unsupported peer topology must not lose the headline line)
            sl = {"error": repr(exc)}
        if rank == 0 and sl is not None:
            keep = ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "config", "roofline", "setup_s", "error")
            also = dict(also or {})
            also["physarum_single_field_%dx%d_slab" % (args.slab_field, args.slab_field)] = {k: sl[k] for k in keep if k in sl}

    # ---- e2e: the same loop through the host-buffer API (numpy obs/action cross PCIe every call) ----
    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(D, torch, dist, args, field, B_local, batched, device, rank, world, M, C, n_gpus)

    # ---- CPU baseline: the oracle on this host, rank 0, N = 1 only ------------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu = time_oracle(args.cpu_field, args.cpu_steps, warmup=2, procs=args.cpu_procs)

    # forward + the cluster-fused step, or forward, move_claim, field_step, agent_feed, finalize_stats (the speculative
    # "fused move" path has no move_claim)
    launches = 2 if "env_step_fused" in meas["kernel_ms"] else 5 - (1 if meas["fused"] else 0)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (gradient-noise food, Bernoulli(0.1) agents, seeded)",
            "config": {"workload": wl_name, "field": list(field), "envs": B_total, "envs_per_gpu": B_local,
                       "agent": "PhysarumAgent", **PHYS, "agent_ratio": AGENT_RATIO, "max_agents_per_env": M,
                       "alive_agents_per_gpu": alive_local, "rng": "philox (in-kernel)",
                       "l2": "per-step working set %.2f GB per GPU >> 126 MB L2 (inputs larger than L2, no flush)"
                             % (step_bytes / 1e9)},
            "agent_steps_per_s": M * B_total / (ms_per_step * 1e-3),
            "alive_agent_steps_per_s": alive_local * n_gpus / (ms_per_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "also": also,
            "gpu_launches": args.steps * launches, "launches_per_step": launches,
            "fused_move": ("commit" if meas.get("committed") else meas["fused"]), "tuning": ARGS.tune, "clocks": meas["clocks"],
            "host": {"cpus": os.cpu_count(), "rank0_bound_to_cpus": bound_cpus},
            "setup_s": round(setup_s, 1),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference)
# --------------------------------------------------------------------------------------------
_ORACLE_BARRIER = None


def _oracle_init(barrier):
    global _ORACLE_BARRIER
    _ORACLE_BARRIER = barrier


def _oracle_worker(job):
    """One process = one 256x256-style env stepped by the numpy oracle (imports no torch, no CUDA).  All workers
    finish their warm-up, meet at a barrier and only then start their timed loops."""
    field_n, steps, warmup, seed = job
    from oracle import die_ref as R
    field = (field_n, field_n)
    np.random.seed(seed)
    env = R.Env(field, R.Dynamics(init_agent_ratio=AGENT_RATIO), noise_seed=seed)
    m = env.agents.shape[-1]
    agent = R.PhysarumAgent(max_agents=m, **PHYS)
    obs = env._get_current_obs
    for _ in range(warmup):
        obs, *_ = env.step(agent.forward(obs))
    if _ORACLE_BARRIER is not None:
        _ORACLE_BARRIER.wait()
    t0 = time.time()
    for _ in range(steps):
        obs, *_ = env.step(agent.forward(obs))
    return t0, time.time()


def time_oracle(field_n, steps, warmup, procs=None):
    """The CPU path on this host: `procs` independent envs (one process each, the way the batched workload
    parallelises on a CPU; the reference itself is single-threaded per env), `steps` steps each.
    value = all cell-updates / (last finish - first start)."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    jobs = [(field_n, steps, warmup, k) for k in range(procs)]
    if procs == 1:
        spans = [_oracle_worker(jobs[0])]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Pool(procs, initializer=_oracle_init, initargs=(ctx.Barrier(procs),)) as pool:
            spans = pool.map(_oracle_worker, jobs, chunksize=1)
    wall = max(e for _, e in spans) - min(b for b, _ in spans)
    per_env_step = float(np.mean([(e - b) / steps for b, e in spans]))
    return {"value": field_n * field_n * steps * procs / wall, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{procs} processes x 1 env of {field_n}x{field_n} x {steps} steps of PhysarumAgent.forward + "
                      f"Env.step (numpy/scipy/pandas oracle port of the reference's CPU path; the reference runs one "
                      f"thread per env), after {warmup} warm-up steps",
            "ms_per_env_step": per_env_step * 1e3, "single_core_value": field_n * field_n / per_env_step,
            "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference
    itself cannot be imported here) on all the host cores: one 256x256 env per process, each bench step =
    one Physarum iteration of every process's env (a bounded sample of the 4096-env workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    field_n = args.cpu_field
    cpu = time_oracle(field_n, args.steps, args.warmup, procs=args.cpu_procs)
    value = cpu["value"]
    wl = "physarum_batched_4096x256x256"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": field_n * field_n * cpu["cores"] / value * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": wl, "sample_field": [field_n, field_n],
                                                            "sample_envs": cpu["cores"],
                                                            "agent": "PhysarumAgent", **PHYS,
                                                            "agent_ratio": AGENT_RATIO},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="die_b200", choices=["die_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "field4096", "batch256", "slab"])
    ap.add_argument("--field", type=int, default=4096, help="side of the single field (field4096 workload)")
    ap.add_argument("--batch", type=int, default=4096, help="total number of 256x256 envs (batch256 workload)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-envs", type=int, default=0,
                    help="cap on the envs per GPU of the host-buffer (e2e) leg (0 = the full per-GPU batch, bounded by "
                         "35 %% of the available host memory for pinned buffers)")
    ap.add_argument("--no-single-field", action="store_true")
    ap.add_argument("--no-small-env", action="store_true", help="skip the configs[0] / configs[1] legs (one 256x256 env)")
    ap.add_argument("--steady", default="300,3000", help="single field: also report ms/step in the 40 steps before these "
                                                          "total step counts ('' = off)")
    ap.add_argument("--no-f32", action="store_true", help="skip the float32-field legs")
    ap.add_argument("--no-commit", action="store_true", help="skip the committed-move leg (agent.fuse_move = 'commit')")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the configs[4] leg (one field over all ranks)")
    ap.add_argument("--slab-field", type=int, default=32768, help="side of the slab-decomposed field of the N > 1 run")
    ap.add_argument("--slab-steps", type=int, default=25)
    ap.add_argument("--cpu-field", type=int, default=256, help="side of each CPU sample env")
    ap.add_argument("--cpu-steps", type=int, default=100, help="steps per CPU env in the cpu_baseline leg")
    ap.add_argument("--cpu-procs", type=int, default=None, help="CPU processes (default: all host cores)")
    ap.add_argument("--corner-r", type=int, default=512, help="slab workload: side of the mirrored corner patches (0 = off)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-workers", type=int, default=4,
                    help="caller threads of the host-buffer (e2e) leg, each stepping its own share of the envs (1 = one thread)")
    ap.add_argument("--fuse", nargs="?", const="spec", default="", choices=["", "spec", "commit"],
                    help="agent.forward also evaluates the move of its action: 'spec' speculatively (the feed kernel commits "
                         "the positions once Env.step adopts it), 'commit' under the run-loop contract (DIE_FWD_COMMIT_MOVE: "
                         "forward stores the positions, Env.step must receive that very action)")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE",
                    help="die_set_tuning switch (result-neutral), e.g. fwd_min_blocks=5, turn_quick=0, pair_mode=0, step_impl=1")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (32 / 64 / 128; 0 = leave the default)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    global ARGS
    ARGS = args
    if args.impl == "reference":
        run_reference(args)
    else:
        run_die_b200(args)


if __name__ == "__main__":
    main()
