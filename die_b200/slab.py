"""One very large field split into row slabs over G GPUs (BASELINE configs[4]; SURVEY section 8e, mode 2).

Rank r owns rows ``[r*H/G, (r+1)*H/G)`` of the medium (and of the claim table, consumed_field and
gradient cache) plus a share of the agent slots; every per-cell array lives in symmetric memory so
the kernels read / atomically update remote cells directly over NVLink (``die_b200/csrc/die_slab.cuh``).
There is no agent migration and no packed halo exchange; ranks meet at three barriers per step and in
one 2-double all-reduce for (reward, num_agents).

Slot ownership keeps the reference's GLOBAL slot ids (they decide "last writer wins"): the reference
creates alive agents in row-major cell order in slots ``[0, A)``, so rank r takes the contiguous run of
alive slots that start on its rows (local at the start, and for as long as agents stay within their
slab) plus an even share of the ghost slots ``[A, M)``.

Two back-ends provide the peer pointer tables:
  * ``SymmetricPeers(group)``  -- one process per GPU, ``torch.distributed._symmetric_memory``;
  * ``EmulatedPeers(G)``       -- all G ranks in ONE process on one GPU (tests; kernels of different
    ranks never wait on each other, so running the phases rank after rank is exact).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .env import Dynamics, _dynamics_to_c


def validate_slab_dynamics(dynamics: Dynamics, cdyn=None) -> None:
    """Everything the slab kernels do NOT implement is refused here, up front and by name (the SLAB instantiations
    hard-code periodic diffusion and the identity food flow; nothing falls back silently)."""
    from .env import identity_food_flow
    cdyn = cdyn if cdyn is not None else _dynamics_to_c(dynamics)
    if dynamics.op_food_flow is not identity_food_flow and dynamics.op_food_flow is not None:
        raise NotImplementedError("slab mode runs the identity food flow only (Dynamics.op_food_flow)")
    if dynamics.apply_sense_mask:
        raise NotImplementedError("slab mode does not implement Dynamics.apply_sense_mask (the observation would be the "
                                  "unmasked slab)")
    if dynamics.diffuse_mode != 'wrap':
        raise NotImplementedError(f"slab mode diffuses periodically only (Dynamics.diffuse_mode={dynamics.diffuse_mode!r})")
    if not 1 <= cdyn.blur_radius <= 4:
        raise NotImplementedError(f"slab mode supports blur radius 1..4 (diffuse_sigma={dynamics.diffuse_sigma} gives "
                                  f"radius {cdyn.blur_radius})")
    if dynamics.agents_die:
        raise NotImplementedError("slab mode does not implement Dynamics.agents_die")


# --------------------------------------------------------------------------------------------
# layout
# --------------------------------------------------------------------------------------------
@dataclass
class SlabLayout:
    G: int
    H: int
    W: int
    M: int
    s0: List[int]
    n0: List[int]
    s1: List[int]
    n1: List[int]

    @property
    def rows_per(self) -> int:
        return self.H // self.G

    def local_slots(self, rank: int) -> int:
        return self.n0[rank] + self.n1[rank]

    def global_ids(self, rank: int) -> np.ndarray:
        return np.concatenate([np.arange(self.s0[rank], self.s0[rank] + self.n0[rank]),
                               np.arange(self.s1[rank], self.s1[rank] + self.n1[rank])])

    def to_c(self, rank: int) -> _lib.DieSlabGeom:
        g = _lib.DieSlabGeom()
        g.G, g.rank, g.H, g.W, g.M = self.G, rank, self.H, self.W, self.M
        for q in range(self.G):
            g.s0[q], g.n0[q], g.s1[q], g.n1[q] = self.s0[q], self.n0[q], self.s1[q], self.n1[q]
        return g


def make_layout(field_size: Tuple[int, int], G: int, M: int, alive_per_slab: Sequence[int]) -> SlabLayout:
    """alive_per_slab[q] = number of alive agents created on slab q's rows (they occupy consecutive
    global slots in rank order); ghosts are split evenly."""
    H, W = field_size
    if H % G != 0:
        raise ValueError(f"H={H} must be divisible by the number of ranks {G}")
    if G > _lib.DIE_MAX_RANKS:
        raise ValueError(f"at most {_lib.DIE_MAX_RANKS} ranks")
    A = int(sum(alive_per_slab))
    s0 = list(np.concatenate([[0], np.cumsum(alive_per_slab)[:-1]]).astype(int))
    n0 = [int(a) for a in alive_per_slab]
    ghosts = M - A
    base, extra = divmod(ghosts, G)
    n1 = [base + (1 if q < extra else 0) for q in range(G)]
    s1 = list((A + np.concatenate([[0], np.cumsum(n1)[:-1]])).astype(int))
    return SlabLayout(G, H, W, M, s0, n0, s1, n1)


def split_global_state(medium: np.ndarray, agents: np.ndarray, G: int):
    """Global (single-Env) state -> (layout, [medium slab per rank], [local agents per rank])."""
    _, H, W = medium.shape
    M = agents.shape[1]
    rows_per = H // G
    alive = agents[2] > 0
    A = int(alive.sum())
    if not alive[:A].all():
        raise ValueError("alive agents must occupy slots [0, A) (the reference's initial layout)")
    rows = np.rint(agents[0, :A] * (H - 1)).astype(np.int64)
    owner = np.minimum(rows // rows_per, G - 1)
    if (np.diff(owner) < 0).any():
        raise ValueError("alive agents must be in row-major order for the slab layout")
    layout = make_layout((H, W), G, M, [int((owner == q).sum()) for q in range(G)])
    mediums = [np.ascontiguousarray(medium[:, q * rows_per:(q + 1) * rows_per]) for q in range(G)]
    locals_ = [np.ascontiguousarray(agents[:, layout.global_ids(q)]) for q in range(G)]
    return layout, mediums, locals_


# --------------------------------------------------------------------------------------------
# peer back-ends
# --------------------------------------------------------------------------------------------
class EmulatedPeers:
    """All ranks in one process / one device."""

    def __init__(self, G: int, device):
        self.G, self.device = G, torch.device(device)
        self._keep = []

    def alloc(self, shapes: Sequence[Tuple[int, ...]], dtype, fill=None):
        """-> (list of per-rank tensors, device table of their base pointers)."""
        ts = [torch.empty(s, dtype=dtype, device=self.device) for s in shapes]
        if fill is not None:
            for t in ts:
                t.fill_(fill)
        table = torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=self.device)
        self._keep.append((ts, table))
        return ts, table


class SymmetricPeers:
    """One process per GPU; torch symmetric memory gives every rank the peers' base pointers."""

    def __init__(self, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.G = dist.get_world_size(self.group)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._handles = []

    def alloc(self, shape: Tuple[int, ...], dtype, fill=None):
        """Collective.  Every rank passes the SAME shape (pad to the largest shard).
        -> (local tensor, device pointer of the [G] table of peer base pointers)."""
        t = self._symm.empty(*shape, dtype=dtype, device=self.device)
        if fill is not None:
            t.fill_(fill)
        else:
            t.zero_()
        hdl = self._symm.rendezvous(t, self.group)
        self._handles.append((t, hdl))
        return t, hdl.buffer_ptrs_dev

    def barrier(self):
        self._handles[0][1].barrier()


# --------------------------------------------------------------------------------------------
# per-rank environment + agent
# --------------------------------------------------------------------------------------------
class SlabRank:
    """The slab of one rank: state tensors + the library handle.  Phases are separate calls; the owner
    (``SlabEnv`` or ``EmulatedSlabWorld``) places the cross-rank barriers between them."""

    def __init__(self, layout: SlabLayout, rank: int, dynamics: Dynamics, tensors: dict, tables: dict, device):
        self._lib = _lib.load()
        self.layout, self.rank, self.device = layout, rank, torch.device(device)
        self.medium = tensors['medium']            # [2] tensors [3, rows_per, W]
        self.action = tensors['action']            # [3, Ml] (symmetric: the field pass of other ranks gathers deposits)
        self.agents = tensors['agents']            # [4, Ml]
        self.theta = tensors['theta']              # [Ml]
        self.stats = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.cur = 0
        self.grad_valid = False
        self.cells_valid = False
        self._handle = _lib.C.c_void_p()
        geom = layout.to_c(rank)
        cdyn = _dynamics_to_c(dynamics)
        validate_slab_dynamics(dynamics, cdyn)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_slab_create(_lib.C.byref(geom), _lib.C.byref(cdyn), _lib.C.byref(self._handle)))
            _lib.check(self._lib.die_slab_bind(self._handle, *[int(tables[k]) for k in
                                                                   ('medium_a', 'medium_b', 'claim', 'consumed', 'grad', 'action')]))

    def __del__(self):
        try:
            if getattr(self, '_handle', None):
                self._lib.die_slab_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    @property
    def Ml(self) -> int:
        return self.layout.local_slots(self.rank)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self, params: _lib.DieGradientParams, coin: Optional[torch.Tensor], seed: int, step: int):
        hints = (1 if self.grad_valid else 0) | (2 if self.cells_valid else 0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_slab_forward(
                self._handle, _lib.C.byref(params), self.cur, self.agents.data_ptr(), self.theta.data_ptr(),
                self.action.data_ptr(), coin.data_ptr() if coin is not None else None, hints, seed, step,
                self._stream()))

    def phase_move(self):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_slab_move_claim(self._handle, self.agents.data_ptr(), self.action.data_ptr(),
                                                     self._stream()))
        self.cells_valid = True

    def phase_field(self):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_slab_field(self._handle, self.cur, 1, self._stream()))
        self.cur = 1 - self.cur
        self.grad_valid = True

    def set_corner_mirror(self, r: int):
        """Mirror the four r x r corner patches (gradient, food, consumed_field) locally, see include/die_b200.h."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_slab_set_corner_mirror(self._handle, int(r)))
        self.corner_r = int(r)

    def phase_corners(self):
        """After the barrier that follows phase_field (every slab's new rows are complete), before phase_feed."""
        if getattr(self, 'corner_r', 0) > 0:
            with torch.cuda.device(self.device):
                _lib.check(self._lib.die_slab_corner_refresh(self._handle, 1 - self.cur, 1, self._stream()))

    def phase_feed(self):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_slab_feed(self._handle, self.agents.data_ptr(), self.action.data_ptr(),
                                               self.stats.data_ptr(), self._stream()))

    def cells(self) -> torch.Tensor:
        from .env import _DevicePtrView
        view = _DevicePtrView(self._lib.die_slab_cells(self._handle), (max(self.Ml, 1),), '<i4')
        with torch.cuda.device(self.device):
            return torch.as_tensor(view, device=self.device)[:self.Ml].clone()


def _physarum_params(scale=0.005, deposit=4.0, sense_offset=0.03, normalized_grad=True, grad_clip=1e-5,
                     turn_angle=30, sense_angle=90, turn_tolerance=0.1) -> _lib.DieGradientParams:
    p = _lib.DieGradientParams()
    p.scale, p.deposit, p.inertia, p.sense_offset, p.noise_scale = scale, deposit, 0.0, sense_offset, 0.0
    p.grad_clip = 0.0 if grad_clip is None else float(grad_clip)
    p.turn_radians, p.sense_radians = float(np.radians(turn_angle)), float(np.radians(sense_angle))
    p.turn_tolerance = turn_tolerance
    p.normalized_grad, p.use_grad_clip, p.discrete_turn = int(normalized_grad), int(grad_clip is not None), 1
    return p


class EmulatedSlabWorld:
    """G ranks in one process on one GPU, stepped phase by phase: the multi-rank kernels exercised
    without a multi-GPU box.  Built from a GLOBAL state so results can be compared with ``Env``."""

    def __init__(self, medium: np.ndarray, agents: np.ndarray, theta: np.ndarray, G: int,
                 dynamics: Optional[Dynamics] = None, device=None, corner_r: int = 0, **physarum_kw):
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.dynamics = dynamics or Dynamics()
        validate_slab_dynamics(self.dynamics)
        self.layout, mediums, locals_ = split_global_state(medium, agents, G)
        L, rp, W = self.layout, self.layout.rows_per, self.layout.W
        peers = EmulatedPeers(G, device)
        self._peers = peers
        med_a, tbl_a = peers.alloc([(3, rp, W)] * G, torch.float64)
        med_b, tbl_b = peers.alloc([(3, rp, W)] * G, torch.float64)
        claim, tbl_c = peers.alloc([(rp * W,)] * G, torch.int32, fill=-1)
        cons, tbl_k = peers.alloc([(rp * W,)] * G, torch.float64, fill=0.0)
        grad, tbl_g = peers.alloc([(rp * W, 2)] * G, torch.float64, fill=0.0)
        act, tbl_act = peers.alloc([(3, max(L.local_slots(q), 1)) for q in range(G)], torch.float64, fill=0.0)
        tables = dict(medium_a=tbl_a.data_ptr(), medium_b=tbl_b.data_ptr(), claim=tbl_c.data_ptr(),
                      consumed=tbl_k.data_ptr(), grad=tbl_g.data_ptr(), action=tbl_act.data_ptr())
        self.params = _physarum_params(**physarum_kw)
        self.ranks: List[SlabRank] = []
        for q in range(G):
            med_a[q].copy_(torch.from_numpy(mediums[q]))
            ids = L.global_ids(q)
            tensors = dict(medium=[med_a[q], med_b[q]], action=act[q],
                           agents=torch.from_numpy(locals_[q]).to(device),
                           theta=torch.from_numpy(np.ascontiguousarray(theta[ids])).to(device))
            self.ranks.append(SlabRank(L, q, self.dynamics, tensors, tables, device))
            if corner_r:
                self.ranks[-1].set_corner_mirror(min(corner_r, L.H // 2, L.W // 2))
        self._step = 0

    def forward(self, coin_global: Optional[np.ndarray] = None):
        for q, r in enumerate(self.ranks):
            coin = None
            if coin_global is not None:
                coin = torch.from_numpy(np.ascontiguousarray(coin_global[self.layout.global_ids(q)]).astype(np.uint8)
                                        ).to(r.device)
            r.forward(self.params, coin, seed=q, step=self._step)

    def step(self):
        for r in self.ranks:
            r.phase_move()
        for r in self.ranks:            # (barrier)
            r.phase_field()
        for r in self.ranks:            # (barrier)
            r.phase_corners()
        for r in self.ranks:
            r.phase_feed()
        self._step += 1
        torch.cuda.synchronize()
        stats = torch.stack([r.stats for r in self.ranks]).cpu().numpy()
        return float(stats[:, 0].sum()), int(round(stats[:, 1].sum())), stats

    # -- gather back to the single-Env layout (tests) ----------------------------------------------
    def gather(self):
        L = self.layout
        medium = np.concatenate([r.medium[r.cur].cpu().numpy() for r in self.ranks], axis=1)
        agents = np.zeros((4, L.M))
        theta = np.zeros(L.M)
        action = np.zeros((3, L.M))
        cells = np.zeros(L.M, dtype=np.int32)
        for q, r in enumerate(self.ranks):
            ids = L.global_ids(q)
            n = len(ids)
            agents[:, ids] = r.agents.cpu().numpy()
            theta[ids] = r.theta.cpu().numpy()
            action[:, ids] = r.action.cpu().numpy()[:, :n]
            cells[ids] = r.cells().cpu().numpy()
        return medium, agents, theta, action, cells


# --------------------------------------------------------------------------------------------
# real multi-process environment (one process per GPU, torchrun)
# --------------------------------------------------------------------------------------------
from .device_init import grid_coords as _grid, device_gradient_noise      # noqa: E402  (device-side initial state)


class SlabEnv:
    """``Env`` for ONE field split over the ranks of a torch.distributed group (one process per GPU).

        env = SlabEnv((32768, 32768), Dynamics(init_agent_ratio=0.1))        # collective
        agent = SlabPhysarumAgent(env, scale=..., sense_offset=...)
        obs = env._get_current_obs
        for i in range(iters):
            action = agent.forward(obs)
            obs, reward, terminated, truncated, info = env.step(action)

    obs = (this rank's agents [4, Ml], this rank's medium slab [3, H/G, W]); reward / info are global
    (one 2-double all-reduce per step).  ``init_state=(layout, medium_slab, agents_local)`` injects a split
    of a global state (``split_global_state``); otherwise the state is generated on the device."""

    def __init__(self, field_size: Tuple[int, int], dynamics: Optional[Dynamics] = None, *, group=None,
                 init_state=None, seed: int = 0, noise_periods: Optional[int] = None, corner_r: int = 512):
        import torch.distributed as dist
        self._dist = dist
        self.dynamics = dynamics or Dynamics()
        validate_slab_dynamics(self.dynamics)          # before any allocation or collective
        self.peers = SymmetricPeers(group)
        self.rank, self.G, self.device = self.peers.rank, self.peers.G, self.peers.device
        H, W = int(field_size[0]), int(field_size[1])
        if H % self.G:
            raise ValueError(f"H={H} must be divisible by the world size {self.G}")
        rp = H // self.G
        M = H * W
        if init_state is not None:
            layout, medium_slab, agents_local = init_state
            medium_slab = torch.as_tensor(medium_slab, dtype=torch.float64).to(self.device)
            agents_local = torch.as_tensor(agents_local, dtype=torch.float64).to(self.device)
        else:
            medium_slab, alive_rc = self._device_init(H, W, rp, seed, noise_periods)
            counts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.G)]
            dist.all_gather(counts, torch.tensor([alive_rc.shape[0]], dtype=torch.int64, device=self.device),
                            group=self.peers.group)
            layout = make_layout((H, W), self.G, M, [int(c.item()) for c in counts])
            agents_local = self._device_agents(layout, alive_rc, H, W, rp, seed)
        self.layout = layout
        Ml = layout.local_slots(self.rank)
        Ml_max = max(layout.local_slots(q) for q in range(self.G))
        P = self.peers
        med_a, tbl_a = P.alloc((3, rp, W), torch.float64)
        med_b, tbl_b = P.alloc((3, rp, W), torch.float64)
        claim, tbl_c = P.alloc((rp * W,), torch.int32, fill=-1)
        cons, tbl_k = P.alloc((rp * W,), torch.float64)
        grad, tbl_g = P.alloc((rp * W, 2), torch.float64)
        act_full, tbl_act = P.alloc((3 * max(Ml_max, 1),), torch.float64)
        med_a.copy_(medium_slab)
        del medium_slab
        self._keep = (claim, cons, grad, act_full)
        action = act_full[:3 * max(Ml, 1)].view(3, max(Ml, 1))
        tensors = dict(medium=[med_a, med_b], action=action, agents=agents_local.contiguous(),
                       theta=torch.zeros(max(Ml, 1), dtype=torch.float64, device=self.device))
        tables = dict(medium_a=tbl_a, medium_b=tbl_b, claim=tbl_c, consumed=tbl_k, grad=tbl_g, action=tbl_act)
        self.slab = SlabRank(layout, self.rank, self.dynamics, tensors, tables, self.device)
        if corner_r and self.G > 1:
            self.slab.set_corner_mirror(min(int(corner_r), H // 2, W // 2))
        self._stats_host = torch.zeros(2, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize(self.device)
        P.barrier()

    # -- device-side initial state (core/env.py:74-86 in spirit; init-time only) -----------------------
    def _device_init(self, H, W, rp, seed, noise_periods):
        periods = noise_periods or max(8, 8 * H // 256)
        lo = self.rank * rp
        gen = torch.Generator(device=self.device)
        gen.manual_seed(seed * 1000003 + self.rank)
        medium = torch.zeros((3, rp, W), dtype=torch.float64, device=self.device)
        chunk = max(1, min(rp, (1 << 25) // W))              # bound the temporaries of the noise evaluation
        for r0 in range(0, rp, chunk):
            r1 = min(rp, r0 + chunk)
            food = device_gradient_noise(H, W, lo + r0, lo + r1, periods, seed, self.device)
            medium[1, r0:r1] = food * ((food >= 0.0) & (food <= 1.0))
            u = torch.round(torch.rand((r1 - r0, W), generator=gen, dtype=torch.float64, device=self.device) * 1000.0) / 1000.0
            medium[0, r0:r1] = ((u > 0.0) & (u <= self.dynamics.init_agent_ratio)).to(torch.float64)
            del food, u
        self._gen = gen
        return medium, torch.nonzero(medium[0])              # row-major (local row, col)

    def _device_agents(self, layout, alive_rc, H, W, rp, seed):
        Ml, n0 = layout.local_slots(self.rank), layout.n0[self.rank]
        agents = torch.zeros((4, max(Ml, 1)), dtype=torch.float64, device=self.device)
        gx = _grid(H, self.rank * rp, (self.rank + 1) * rp, self.device)
        gy = _grid(W, 0, W, self.device)
        agents[0, :n0] = gx[alive_rc[:, 0]]
        agents[1, :n0] = gy[alive_rc[:, 1]]
        agents[2, :n0] = 1.0
        u = torch.round(torch.rand(n0, generator=self._gen, dtype=torch.float64, device=self.device) * 1000.0) / 1000.0
        agents[3, :n0] = 0.9 * u + 0.1
        return agents

    # -- the reference protocol ---------------------------------------------------------------------------
    @property
    def agents(self):
        return self.slab.agents

    @property
    def medium(self):
        return self.slab.medium[self.slab.cur]

    @property
    def _get_current_obs(self):
        return self.agents, self.medium

    def step_async(self, action=None, reduce_stats: bool = True):
        """All kernels + barriers (+ the stats all-reduce unless ``reduce_stats=False``) enqueued; returns
        the device stats tensor [sum gained, num alive] (global, or this rank's share when not reduced)."""
        s = self.slab
        if action is not None and action.data_ptr() != s.action.data_ptr():
            s.action.copy_(action)
        s.phase_move()
        self.peers.barrier()          # every claim is in before any slab is blurred
        s.phase_field()
        self.peers.barrier()          # consumed_field / new medium complete before anybody gathers from it
        s.phase_corners()             # pull the corner patches (where the ghost slots live) into local memory
        s.phase_feed()
        self.peers.barrier()          # claim table clean before the next step's claims
        if reduce_stats:
            self._dist.all_reduce(s.stats, group=self.peers.group)
        return self._get_current_obs, s.stats

    def step(self, action=None):
        obs, stats = self.step_async(action)
        self._stats_host.copy_(stats, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        reward, n = float(self._stats_host[0]), int(round(float(self._stats_host[1])))
        mean = reward / n if n > 0 else 0.
        info = {'num_agents': n, 'reward': np.round(reward, 3), 'mean_reward': np.round(mean, 5)}
        return obs, reward, n == 0, False, info


class SlabPhysarumAgent:
    """PhysarumAgent (core/agent/gradient.py:138-219, default inertia = noise = 0) for a SlabEnv."""

    def __init__(self, env: SlabEnv, *, seed: int = 0, theta=None, **physarum_kw):
        self.env = env
        self.params = _physarum_params(**physarum_kw)
        self._seed, self._step = seed + 7919 * env.rank, 0
        Ml = env.layout.local_slots(env.rank)
        if theta is None:
            tr = self.params.turn_radians
            nlat = int(round(2 * np.pi / tr))
            gen = torch.Generator(device=env.device)
            gen.manual_seed(self._seed)
            k = torch.randint(-nlat // 2, nlat - nlat // 2, (max(Ml, 1),), generator=gen, device=env.device)
            theta = k.to(torch.float64) * tr
        env.slab.theta.copy_(torch.as_tensor(theta, dtype=torch.float64).to(env.device).reshape(-1)[:max(Ml, 1)])
        self._coin = None

    def forward(self, obs=None, coin=None):
        s = self.env.slab
        coin_t = None
        if coin is not None:
            coin_t = torch.as_tensor(np.ascontiguousarray(coin).astype(np.uint8)).to(s.device)
            self._coin = coin_t
        s.forward(self.params, coin_t, self._seed, self._step)
        self._step += 1
        return s.action
