"""TEST INFRASTRUCTURE: a numpy driver over the C ABI of libdie_hostsim.so (the kernel sources of die_b200/csrc
executed thread by thread on the CPU, tests/hostsim/cuda_runtime.h).  "Device pointers" are numpy arrays.

It mirrors what die_b200/env.py and die_b200/agent/*.py do with torch CUDA tensors -- the same C entry points in the
same order with the same flags -- so the `-m "not gpu"` suite can check the kernels' logic against the oracle
without a GPU.  The product never imports this module.
"""
import ctypes as C

import numpy as np

from die_b200 import _lib as L          # the prototypes (SIGNATURES) and the C structs only
from die_b200.env import Dynamics, _dynamics_to_c

from . import build as _build

_sim = None


def lib() -> C.CDLL:
    global _sim
    if _sim is None:
        so = C.CDLL(_build.build())
        for name, (res, args) in L.SIGNATURES.items():
            fn = getattr(so, name)
            fn.restype = res
            fn.argtypes = args
        _sim = so
    return _sim


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libdie_hostsim error {rc}: {lib().die_last_error().decode()}")


class _Fenced:
    """Keeps a fenced allocation alive as long as the numpy array built on it."""

    def __init__(self, nbytes):
        so = lib()
        so.hostsim_alloc.restype, so.hostsim_alloc.argtypes = C.c_void_p, [C.c_size_t]
        so.hostsim_free.restype, so.hostsim_free.argtypes = None, [C.c_void_p]
        self._so, self.nbytes = so, nbytes
        self.addr = so.hostsim_alloc(nbytes)
        if not self.addr:
            raise MemoryError(nbytes)

    def __del__(self):
        try:
            self._so.hostsim_free(self.addr)
        except Exception:
            pass


def fenced(shape, dtype=np.float64, fill=None):
    """A numpy array in "device" memory of the emulator: it ends at a page boundary followed by an inaccessible page
    (and an inaccessible page precedes it), so a kernel that runs off either end faults instead of corrupting a
    neighbour -- what compute-sanitizer's memcheck does on the GPU."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    nbytes = max(n * dtype.itemsize, 1)
    pad = (-nbytes) % 16                                   # hostsim_alloc rounds up to 16 bytes: keep OUR end on the fence
    blk = _Fenced(nbytes + pad)
    buf = (C.c_char * (nbytes + pad)).from_address(blk.addr)
    buf._fence = blk
    arr = np.frombuffer(buf, dtype=dtype, count=n, offset=pad).reshape(shape)
    if fill is not None:
        arr[...] = fill
    return arr


def fenced_copy(a, dtype=None):
    a = np.asarray(a, dtype=dtype)
    out = fenced(a.shape, a.dtype)
    out[...] = a
    return out


def ptr(a):
    return None if a is None else a.ctypes.data


def set_tuning(key: str, value: int) -> None:
    check(lib().die_set_tuning(key.encode(), int(value)))


def gradient_params(scale=0.005, deposit=4.0, inertia=0.0, sense_offset=0.03, noise_scale=0.0, normalized_grad=True,
                    grad_clip=1e-5, turn_angle=30, sense_angle=90, turn_tolerance=0.1, discrete_turn=True):
    """PhysarumAgent's constructor defaults (core/agent/gradient.py:139-151) -> die_gradient_params_t."""
    p = L.DieGradientParams()
    p.scale, p.deposit, p.inertia, p.sense_offset, p.noise_scale = scale, deposit, inertia, sense_offset, noise_scale
    p.grad_clip = 0.0 if grad_clip is None else float(grad_clip)
    p.turn_radians = float(np.radians(turn_angle)) if discrete_turn else 0.0
    p.sense_radians = float(np.radians(sense_angle)) if discrete_turn else 0.0
    p.turn_tolerance = float(turn_tolerance) if discrete_turn else 0.0
    p.normalized_grad = int(normalized_grad)
    p.use_grad_clip = int(grad_clip is not None)
    p.discrete_turn = int(discrete_turn)
    return p


class SimEnv:
    """die_b200.Env's call sequence on numpy arrays."""

    def __init__(self, field_size, medium, agents, dynamics: Dynamics = None, batch=None, field_dtype=np.float64):
        self.lib = lib()
        self.h, self.w = int(field_size[0]), int(field_size[1])
        self.B = int(batch) if batch is not None else 1
        self.dynamics = dynamics or Dynamics()
        self.agents = fenced_copy(np.asarray(agents, dtype=np.float64).reshape(self.B, 4, -1))
        self.M = self.agents.shape[-1]
        self.field_dtype = np.dtype(field_dtype)
        first = fenced_copy(np.asarray(medium, dtype=self.field_dtype).reshape(self.B, 3, self.h, self.w))
        self.medium_buf = [first, fenced(first.shape, self.field_dtype, fill=np.nan)]
        self.cur = 0
        self.reward = fenced((self.B,), fill=0.0)
        self.alive = fenced((self.B,), np.int64, fill=0)
        self.handle = C.c_void_p()
        cdyn = _dynamics_to_c(self.dynamics)
        check(self.lib.die_env_create(self.h, self.w, self.M, self.B, C.byref(cdyn), C.byref(self.handle)))
        if self.field_dtype == np.float32:
            check(self.lib.die_env_set_field_dtype(self.handle, L.FIELD_F32))
        self.publish_grad = False
        self.hints_valid = False
        self._flow = None

    def __del__(self):
        try:
            if self.handle:
                self.lib.die_env_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def medium(self):
        return self.medium_buf[self.cur]

    def set_food_frames(self, op):
        """Dynamics.op_food_flow = any other FieldSequence's flow operator: tabulated frames."""
        frames = fenced_copy(op.sequence.frames())
        self._flow = (frames,)
        check(self.lib.die_env_set_food_frames(self.handle, ptr(frames), frames.shape[0], op.calls % frames.shape[0],
                                               op.scale, op.decay))

    def set_food_flow(self, op):
        """Dynamics.op_food_flow = WaveSequence flow operator (die_b200/env.py:_install_food_flow)."""
        rwave, col, row = (np.ascontiguousarray(a) for a in op.sequence.device_tables())
        ts = np.ascontiguousarray(op.sequence.ts, dtype=np.float64)
        self._flow = (rwave, col, row, ts)
        check(self.lib.die_env_set_food_flow(self.handle, ptr(rwave), ptr(col), ptr(row), ptr(ts), len(ts),
                                             op.calls % len(ts), op.scale, op.decay))

    def want_gradient(self):
        if not self.publish_grad:
            check(self.lib.die_env_publish_gradient(self.handle, 1))
            self.publish_grad = True

    def step(self, action, flags=L.STEP_ALIVE_BITS):
        action = fenced_copy(np.asarray(action, dtype=np.float64).reshape(self.B, 3, self.M))
        nxt = 1 - self.cur
        check(self.lib.die_env_refresh_alive(self.handle, ptr(self.agents), None))
        check(self.lib.die_env_step_flags(self.handle, ptr(self.medium_buf[self.cur]), ptr(self.medium_buf[nxt]),
                                          ptr(self.agents), ptr(action), ptr(self.reward), ptr(self.alive), flags, None))
        self.cur = nxt
        self.hints_valid = True
        self.grad_valid = self.publish_grad
        return self.reward.copy(), self.alive.copy()

    def step_host(self, action, agents_host=None, flags=0):
        """die_env_step_host: action / observation through "host" buffers, chunked over the batch.  With `agents_host` (a
        buffer kept from an earlier call) and flags = HOST_KEEP_ALIVE_CHANNEL: die_env_step_host_flags."""
        action = np.ascontiguousarray(np.asarray(action, dtype=np.float64).reshape(self.B, 3, self.M))
        nxt = 1 - self.cur
        agents_h = np.empty_like(self.agents) if agents_host is None else agents_host
        medium_h = np.empty_like(self.medium_buf[0])
        if agents_host is not None or flags:
            check(self.lib.die_env_step_host_flags(self.handle, ptr(self.medium_buf[self.cur]), ptr(self.medium_buf[nxt]),
                                                   ptr(self.agents), ptr(action), None, ptr(agents_h), ptr(medium_h),
                                                   ptr(self.reward), ptr(self.alive), flags, None))
        else:
            check(self.lib.die_env_step_host(self.handle, ptr(self.medium_buf[self.cur]), ptr(self.medium_buf[nxt]),
                                             ptr(self.agents), ptr(action), ptr(agents_h), ptr(medium_h),
                                             ptr(self.reward), ptr(self.alive), None))
        self.cur = nxt
        self.hints_valid = True
        self.grad_valid = self.publish_grad
        return (agents_h, medium_h), self.reward.copy(), self.alive.copy()

    def cells(self):
        p = self.lib.die_env_cells(self.handle)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(self.B, self.M)).copy()

    def gradient(self):
        p = self.lib.die_env_gradient(self.handle)
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(self.B, self.h, self.w, 2)).copy()


class SimGradientAgent:
    """die_b200.GradientAgent / PhysarumAgent's call sequence on numpy arrays."""

    def __init__(self, M, B=1, seed=0, **params):
        self.lib = lib()
        self.p = gradient_params(**params)
        self.M, self.B = M, B
        self.theta = fenced((B, M), fill=0.0)
        needs_prev = (self.p.inertia != 0.0 or self.p.noise_scale != 0.0 or
                      not (self.p.discrete_turn and self.p.normalized_grad))       # die_b200/agent/gradient.py:_needs_prev
        self.prev_grad = fenced((B, 2, M), fill=0.0) if needs_prev else None
        self.action = fenced((B, 3, M), fill=np.nan)
        self.sense_cells = fenced((B, M), np.int32, fill=0)
        self.record_sense_cells = False
        self.seed, self.step_no = seed, 0
        self.fuse_move = False
        self.write_cost = False

    def forward(self, env: SimEnv, coin=None, noise=None, use_hints=True, obs=None):
        """obs = (agents, medium) arrays other than the env's own disable the hints, as die_b200/_hints.py does."""
        agents, medium = (env.agents, env.medium) if obs is None else obs
        coin_a = None if coin is None else fenced_copy(np.asarray(coin).astype(np.uint8).reshape(self.B, self.M))
        noise_a = None if noise is None else fenced_copy(np.asarray(noise, dtype=np.float64).reshape(self.B, 2, self.M))
        sc = self.sense_cells if self.record_sense_cells else None
        own = obs is None
        if own and use_hints:
            env.want_gradient()
            flags = 0
            if env.hints_valid:
                flags |= L.FWD_USE_CELLS
                if getattr(env, 'grad_valid', False):
                    flags |= L.FWD_USE_GRADIENT
            if self.fuse_move:
                check(self.lib.die_env_refresh_alive(env.handle, ptr(env.agents), None))
                flags |= L.FWD_SPECULATE_MOVE
                if self.fuse_move == 'commit':
                    flags |= L.FWD_COMMIT_MOVE
            if self.write_cost:
                flags |= L.FWD_WRITE_COST
            check(self.lib.die_env_forward_gradient(env.handle, C.byref(self.p), ptr(agents), ptr(medium), ptr(self.theta),
                                                    ptr(self.prev_grad), ptr(self.action), ptr(coin_a), ptr(noise_a),
                                                    ptr(sc), flags, self.seed, self.step_no, None))
            self.last_flags = flags
        else:
            check(self.lib.die_gradient_forward(C.byref(self.p), env.h, env.w, self.M, self.B, ptr(agents), ptr(medium),
                                                ptr(self.theta), ptr(self.prev_grad), ptr(self.action), ptr(coin_a),
                                                ptr(noise_a), ptr(sc), None, None, self.seed, self.step_no, None))
            self.last_flags = 0
        self.step_no += 1
        return self.action


class SimJonesAgent:
    """die_b200.JonesAgent's call on numpy arrays (die_jones_forward)."""

    def __init__(self, M, B=1, seed=0, scale=0.005, deposit=4.0, sense_offset=0.03, turn_angle=45, sense_angle=45):
        self.lib = lib()
        p = self.p = L.DieJonesParams()
        p.scale, p.deposit, p.sense_offset = scale, deposit, sense_offset
        p.sense_radians, p.turn_radians = float(np.radians(sense_angle)), float(np.radians(turn_angle))
        self.M, self.B = M, B
        self.theta = fenced((B, M), fill=0.0)
        self.action = fenced((B, 3, M), fill=np.nan)
        self.seed, self.step_no = seed, 0

    def forward(self, env, coin=None):
        coin_a = None if coin is None else fenced_copy(np.asarray(coin).astype(np.uint8).reshape(self.B, self.M))
        f32 = int(env.medium.dtype == np.float32)
        check(self.lib.die_jones_forward(C.byref(self.p), env.h, env.w, self.M, self.B, ptr(env.agents), ptr(env.medium), f32,
                                         ptr(self.theta), ptr(self.action), ptr(coin_a), self.seed, self.step_no, 0, None))
        self.step_no += 1
        return self.action


def brownian_forward(agents, move_scale=0.01, deposit_scale=0.5, u=None, seed=0, step=0):
    agents = np.ascontiguousarray(agents, dtype=np.float64)
    B = 1 if agents.ndim == 2 else agents.shape[0]
    M = agents.shape[-1]
    agents = fenced_copy(agents)
    action = fenced((B, 3, M), fill=np.nan)
    u_a = None if u is None else fenced_copy(np.asarray(u, dtype=np.float64).reshape(B, 3, M))
    check(lib().die_brownian_forward(ptr(agents), ptr(action), M, B, move_scale, deposit_scale, ptr(u_a), seed, step, None))
    return action if agents.ndim == 3 else action[0]


def conv_policy_forward(env, weights, coefs=(0.1, 0.1, 1.0), with_agent_channel=True, use_cells=False, population=None):
    """die_b200.NeuralAutomataAgent.forward's call on numpy arrays -> (action [B, 3, M] float64, model output [B, 3, H, W]).
    population: a list of B weight lists (one model per environment) instead of `weights`."""
    if population is not None:
        weights = population[0]
    ks = [int(w.shape[-1]) for w in weights]
    cin = int(weights[0].shape[1])
    sets = population if population is not None else [weights]
    rows = [np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in ws]) for ws in sets]
    stride = rows[0].size if population is not None else 0
    flat = fenced_copy(np.concatenate(rows), np.float32)
    ch = max(cin, 3)
    sa, sb = fenced((env.B, ch, env.h, env.w), np.float32, fill=np.nan), fenced((env.B, ch, env.h, env.w), np.float32, fill=np.nan)
    action = fenced((env.B, 3, env.M), fill=np.nan)
    ksa = (C.c_int32 * len(ks))(*ks)
    cf = (C.c_float * 3)(*coefs)
    final = C.c_int32(0)
    dtype = L.FIELD_F32 if env.medium.dtype == np.float32 else L.FIELD_F64
    cells = lib().die_env_cells(env.handle) if use_cells else None
    check(lib().die_conv_policy_forward_population(
        env.h, env.w, env.M, env.B, dtype, ptr(env.medium), 3, 0 if with_agent_channel else 1,
        cin, 3, len(ks), ksa, ptr(flat), stride, ptr(sa), ptr(sb), ptr(env.agents), cells, cf,
        ptr(action), C.byref(final), None))
    out = (sa, sb)[final.value]
    return action, out[:, :3].copy()


class SimSlabWorld:
    """die_b200.slab.EmulatedSlabWorld's call sequence on numpy arrays: one field split into row slabs over G
    "ranks", all in this process; the peer tables are arrays of numpy base addresses."""

    def __init__(self, medium, agents, theta, G, dynamics: Dynamics = None, corner_r=0, **physarum_kw):
        from die_b200.slab import split_global_state, _physarum_params
        self.lib = lib()
        self.dynamics = dynamics or Dynamics()
        self.layout, mediums, locals_ = split_global_state(medium, agents, G)
        Lo, rp, W = self.layout, self.layout.rows_per, self.layout.W
        self.G = G

        def alloc(shapes, dtype, fill):
            arrs = [fenced(s, dtype, fill=fill) for s in shapes]
            return arrs, np.array([a.ctypes.data for a in arrs], dtype=np.int64)

        self.med_a, self.tbl_a = alloc([(3, rp, W)] * G, np.float64, np.nan)
        self.med_b, self.tbl_b = alloc([(3, rp, W)] * G, np.float64, np.nan)
        self.claim, self.tbl_c = alloc([(rp * W,)] * G, np.int32, -1)
        self.cons, self.tbl_k = alloc([(rp * W,)] * G, np.float64, 0.0)
        self.grad, self.tbl_g = alloc([(rp * W, 2)] * G, np.float64, 0.0)
        self.act, self.tbl_act = alloc([(3, max(Lo.local_slots(q), 1)) for q in range(G)], np.float64, 0.0)
        self.params = _physarum_params(**physarum_kw)
        self.handles, self.agents, self.theta, self.stats = [], [], [], []
        self.cur = 0
        self.grad_valid = self.cells_valid = False
        self.corner_r = 0
        cdyn = _dynamics_to_c(self.dynamics)
        for q in range(G):
            self.med_a[q][...] = mediums[q]
            self.agents.append(fenced_copy(locals_[q]))
            self.theta.append(fenced_copy(theta[Lo.global_ids(q)]))
            self.stats.append(np.zeros(2))
            h = C.c_void_p()
            geom = Lo.to_c(q)
            check(self.lib.die_slab_create(C.byref(geom), C.byref(cdyn), C.byref(h)))
            check(self.lib.die_slab_bind(h, ptr(self.tbl_a), ptr(self.tbl_b), ptr(self.tbl_c), ptr(self.tbl_k),
                                         ptr(self.tbl_g), ptr(self.tbl_act)))
            if corner_r:
                self.corner_r = min(corner_r, Lo.H // 2, Lo.W // 2)
                check(self.lib.die_slab_set_corner_mirror(h, self.corner_r))
            self.handles.append(h)
        self._step = 0

    def __del__(self):
        try:
            for h in self.handles:
                self.lib.die_slab_destroy(h)
            self.handles = []
        except Exception:
            pass

    def forward(self, coin_global=None):
        hints = (1 if self.grad_valid else 0) | (2 if self.cells_valid else 0)
        for q, h in enumerate(self.handles):
            coin = None
            if coin_global is not None:
                coin = fenced_copy(coin_global[self.layout.global_ids(q)].astype(np.uint8))
            check(self.lib.die_slab_forward(h, C.byref(self.params), self.cur, ptr(self.agents[q]), ptr(self.theta[q]),
                                            ptr(self.act[q]), ptr(coin), hints, q, self._step, None))

    def step(self):
        for q, h in enumerate(self.handles):
            check(self.lib.die_slab_move_claim(h, ptr(self.agents[q]), ptr(self.act[q]), None))
        self.cells_valid = True
        for h in self.handles:                      # (barrier)
            check(self.lib.die_slab_field(h, self.cur, 1, None))
        self.cur = 1 - self.cur
        self.grad_valid = True
        if self.corner_r > 0:
            for h in self.handles:                  # (barrier)
                check(self.lib.die_slab_corner_refresh(h, 1 - self.cur, 1, None))
        for q, h in enumerate(self.handles):
            check(self.lib.die_slab_feed(h, ptr(self.agents[q]), ptr(self.act[q]), ptr(self.stats[q]), None))
        self._step += 1
        stats = np.stack(self.stats)
        return float(stats[:, 0].sum()), int(round(stats[:, 1].sum()))

    def gather(self):
        Lo = self.layout
        med = self.med_a if self.cur == 0 else self.med_b
        medium = np.concatenate(med, axis=1)
        agents, theta = np.zeros((4, Lo.M)), np.zeros(Lo.M)
        action, cells = np.zeros((3, Lo.M)), np.zeros(Lo.M, dtype=np.int32)
        for q, h in enumerate(self.handles):
            ids = Lo.global_ids(q)
            n = len(ids)
            agents[:, ids] = self.agents[q]
            theta[ids] = self.theta[q]
            action[:, ids] = self.act[q][:, :n]
            p = self.lib.die_slab_cells(h)
            cells[ids] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n]
        return medium, agents, theta, action, cells
