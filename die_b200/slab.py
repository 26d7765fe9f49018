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
                 dynamics: Optional[Dynamics] = None, device=None, **physarum_kw):
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.dynamics = dynamics or Dynamics()
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
