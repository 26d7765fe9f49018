"""CPU-side checks: the C-ABI library loads and exports every symbol include/die_b200.h
declares (no compute calls without a GPU); host-side logic of the Python mirror."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "die_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(die_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for must in ("die_env_create", "die_env_destroy", "die_env_step", "die_env_step_host",
                 "die_brownian_forward", "die_gradient_forward", "die_const_forward", "die_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from die_b200._build import LIB_PATH
    assert os.path.exists(LIB_PATH), "build with `python die_b200/_build.py`"
    lib = ctypes.CDLL(LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/die_b200.h but not exported"


def test_ctypes_signatures_cover_the_header():
    from die_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_functions()
    lib = _lib.load()
    assert b"sm_100a" in lib.die_version()


def test_struct_layouts_match_header():
    from die_b200 import _lib
    assert ctypes.sizeof(_lib.DieDynamics) == 8 * 4 + 8 * 17 + 4 * 5 + 4          # five int32 + tail padding to 8
    assert ctypes.sizeof(_lib.DieGradientParams) == 8 * 9 + 4 * 4


def test_argument_errors_are_codes_not_crashes():
    from die_b200 import _lib
    lib = _lib.load()
    out = ctypes.c_void_p()
    dyn = _lib.DieDynamics()
    assert lib.die_env_create(1, 1, 1, 1, ctypes.byref(dyn), ctypes.byref(out)) == 1      # DIE_E_INVALID
    assert b"invalid argument" in lib.die_last_error()
    dyn.blur_radius = 99
    assert lib.die_env_create(8, 8, 64, 1, ctypes.byref(dyn), ctypes.byref(out)) == 1
    assert lib.die_env_destroy(None) == 0
    assert lib.die_brownian_forward(None, None, 1, 1, 0.1, 0.1, None, 0, 0, None) == 1


def test_dynamics_to_c_defaults():
    from die_b200 import env as E
    c = E._dynamics_to_c(E.Dynamics())
    assert (c.rate_feed, c.rate_decay_chem) == (0.1, 0.1)
    assert (c.cost_w_deposit, c.cost_w_dist) == (0.02, 0.01)
    assert c.blur_radius == 2 and c.boundary == 0 and c.food_infinite == 0
    import scipy.ndimage
    imp = np.zeros(9)
    imp[4] = 1.0
    w = scipy.ndimage.gaussian_filter1d(imp, 0.5, mode='constant')[2:7]
    assert np.array_equal(np.array(list(c.blur_w)[:5]), w)


def test_dynamics_rejects_what_is_not_on_the_gpu_path():
    from die_b200 import env as E
    with pytest.raises(NotImplementedError):
        E._dynamics_to_c(E.Dynamics(op_action_cost=lambda a: 0))
    with pytest.raises(ValueError):
        E._dynamics_to_c(E.Dynamics(diffuse_mode='periodic'))
    assert [E._dynamics_to_c(E.Dynamics(diffuse_mode=m)).diffuse_mode
            for m in ('wrap', 'reflect', 'nearest', 'mirror', 'constant')] == [0, 1, 2, 3, 4]
    with pytest.raises(NotImplementedError):
        E._dynamics_to_c(E.Dynamics(diffuse_sigma=5.0))
    assert E._dynamics_to_c(E.Dynamics(op_action_cost=E.zero_cost)).cost_w_dist == 0.0
    assert E._dynamics_to_c(E.Dynamics(boundary=E.BoundaryCondition.limit)).boundary == 1


def test_env_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has CUDA")
    import die_b200 as D
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D.Env((16, 16))


def test_host_init_matches_oracle_init():
    """die_b200.data_init (product) and oracle.die_ref (checker) are separate implementations of
    core/data_init.py; on the same seeds they must build the same state."""
    from die_b200 import data_init
    from oracle import die_ref as R
    np.random.seed(3)
    m1 = data_init.init_medium((40, 56), 0.1, noise_seed=5)
    a1 = data_init.agents_from_medium(m1)
    np.random.seed(3)
    m2 = R.init_medium((40, 56), 0.1, noise_seed=5)
    a2 = R.agents_from_medium(m2)
    assert np.array_equal(m1, m2) and np.array_equal(a1, a2)
    n = int(a1[2].sum())
    assert 0.05 < n / (40 * 56) < 0.15
    assert (a1[:, n:] == 0).all()                       # ghosts: all-zero slots at (0, 0)
    assert ((a1[3, :n] >= 0.1) & (a1[3, :n] <= 1.0)).all()
    ix = np.rint(a1[0, :n] * 39).astype(int) * 56 + np.rint(a1[1, :n] * 55).astype(int)
    assert (np.diff(ix) > 0).all()                      # row-major nonzero order


def test_agent_params_roundtrip(tmp_path):
    import die_b200 as D
    a = D.PhysarumAgent(max_agents=64, scale=0.007, turn_angle=30, sense_offset=0.04)
    f = tmp_path / "agent.json"
    a.save(str(f))
    b = D.PhysarumAgent.load(str(f))
    assert b.init_params() == a.init_params()
    assert D.BrownianAgent.load.__self__ is D.BrownianAgent


def test_wave_sequence_host_mirror_and_device_tables():
    """die_b200.WaveSequence (host mirror of core/data_init.py:71-89) equals the oracle's restatement, and
    the tables handed to the field kernel recombine to the same field bit for bit (so the only arithmetic
    the kernel adds is one cosine per cell)."""
    from oracle import die_ref as R
    from die_b200 import data_init as DI
    for field in ((20, 33), (64, 64), (5, 9)):
        ws, wd = R.WaveSequence(field), DI.WaveSequence(field)
        assert len(ws) == len(wd) == 1000 and np.array_equal(wd.ts, np.arange(0, 10, 0.01))
        rwave, col, row = wd.device_tables()
        assert rwave.shape == field and col.shape == (1000, field[1]) and row.shape == (1000, field[0])
        for k in (0, 37, 999):
            t = wd.ts[k]
            assert np.array_equal(ws[t], wd[t])
            z = 0.75 * np.cos(np.pi * (rwave + t)) + 0.25 * (col[k][None, :] + row[k][:, None])
            assert np.array_equal(z, wd[t])
    op, oo = DI.WaveSequence((12, 10)).get_flow_operator(0.5, 0.5), R.WaveSequence((12, 10)).get_flow_operator(0.5, 0.5)
    f = np.random.default_rng(0).random((12, 10))
    for _ in range(3):
        assert np.array_equal(op(f), oo(f))
    assert op.calls == 3


def test_generic_field_sequences_on_the_host():
    """TabulatedSequence / PerlinNoiseSequence (core/data_init.py:16-68): frames, the flow operator evaluated on the
    host equals the oracle's, the iterator cycles; an arbitrary Python operator is still refused by the env."""
    from oracle import die_ref as R
    import die_b200 as D
    from die_b200.env import _dynamics_to_c
    rng = np.random.default_rng(0)
    frames = rng.normal(size=(3, 6, 7)).round(3)
    seq = D.TabulatedSequence(frames)
    assert len(seq) == 3 and seq.frames().shape == (3, 6, 7) and np.array_equal(seq[seq.ts[2]], frames[2])
    op, oo = seq.get_flow_operator(0.5, 0.25), R.FrameSequence(frames).get_flow_operator(0.5, 0.25)
    f = rng.random((6, 7))
    for _ in range(7):
        assert np.array_equal(op(f), oo(f))
    assert op.calls == 7
    _dynamics_to_c(D.Dynamics(op_food_flow=op))                       # accepted
    with pytest.raises(NotImplementedError):
        _dynamics_to_c(D.Dynamics(op_food_flow=lambda food: food * 0.5))
    # the reference's PerlinNoiseSequence with an injected noise callable (its default needs the perlin_noise package)
    pn = D.PerlinNoiseSequence((4, 5), dt=0.25, t_bounds=(0, 1), noise=lambda p: p[0] - 2 * p[1] + 0.1234567 * p[2])
    fr = pn.frames()
    assert fr.shape == (4, 4, 5)
    xs, ys = np.linspace(0, 1, 4), np.linspace(0, 1, 5)
    assert np.array_equal(fr[2], (xs[:, None] - 2 * ys[None, :] + 0.1234567 * 0.5).round(3))


def test_agent_postprocess_action_helpers():
    """core/agent/base.py:45-62: mask the action by alive-ness, rescale = identity."""
    import torch
    from die_b200 import Agent
    agents = torch.tensor([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6], [1.0, 0.0, 1.0], [0.5, 0.5, 0.5]], dtype=torch.float64)
    action = torch.arange(9, dtype=torch.float64).reshape(3, 3) + 1
    out = Agent.postprocess_action(agents, action)
    assert torch.equal(out, action * torch.tensor([1.0, 0.0, 1.0], dtype=torch.float64))
    batched = Agent._masked_alive(agents.expand(2, 4, 3), action.expand(2, 3, 3))
    assert batched.shape == (2, 3, 3) and torch.equal(batched[1], out)
    assert Agent._rescale_outputs(action) is action


def test_import_fails_loudly_without_the_cuda_library(tmp_path):
    """No CPU / PyTorch fallback: with libdie_sm100a.so missing, `import die_b200` itself raises (in a child process,
    with the library path pointed at a file that does not exist)."""
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f'''
        import importlib.util, sys
        sys.path.insert(0, {root!r})
        spec = importlib.util.spec_from_file_location("die_b200._build", {os.path.join(root, "die_b200", "_build.py")!r})
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.LIB_PATH = {str(tmp_path / "libdie_sm100a.so")!r}
        sys.modules["die_b200._build"] = mod
        try:
            import die_b200
        except ImportError as exc:
            print("ImportError:", exc)
            sys.exit(0)
        sys.exit(1)
    ''')
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CPU / PyTorch fallback" in out.stdout


def test_examples_and_tools_compile(tmp_path):
    """The example scripts and tools are not imported by any CPU test (they need a GPU to run): at least they parse."""
    import glob
    import py_compile
    files = glob.glob(os.path.join(ROOT, "examples", "*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py")) \
        + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) >= 10
    for f in files:
        py_compile.compile(f, doraise=True, cfile=str(tmp_path / (os.path.basename(f) + 'c')))
