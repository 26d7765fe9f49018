"""NeuralAutomataAgent (core/agent/evo.py:117-209): the convolution-stack policy.  CPU part: the CUDA kernel sources
(die_conv_kernels.cuh) under the emulator against the oracle, whose restatement equals the reference's own code executed
live (tests/test_golden_oracle.py).  float32 arithmetic whose accumulation order differs from torch's: the tolerance is
1e-5 relative + 5e-6 absolute (sums of up to 147 float32 products of size ~1, layer after layer) on the model output and
the action; the gather (cells) is exact."""
import numpy as np
import pytest

from oracle import die_ref as R

S = pytest.importorskip("tests.hostsim.sim")
import die_b200 as D                                    # noqa: E402
from tests.test_hostsim_kernels import make_pair       # noqa: E402


def _weights(kernel_sizes, cin, seed):
    rng = np.random.default_rng(seed)
    outs = [cin] * (len(kernel_sizes) - 1) + [3]
    return [rng.uniform(-0.3, 0.3, size=(co, cin, k, k)).astype(np.float32) for k, co in zip(kernel_sizes, outs)]


@pytest.mark.parametrize("field,kernel_sizes,with_agent_channel,batch", [((24, 40), (3,), True, None), ((33, 37), (3, 5), True, 2),
                                                                        ((20, 64), (5, 3, 3), False, None), ((12, 12), (7,), True, 3),
                                                                        ((5, 9), (7, 7), True, None)])
def test_conv_policy_kernels_against_the_oracle(field, kernel_sizes, with_agent_channel, batch):
    refs, env = make_pair(field, seed=4, ratio=0.2, batch=batch)
    rng = np.random.default_rng(1)
    for r in refs:                                       # a chem1 field to look at
        r.medium[2] = rng.random(field) * 3.0
    env.medium[...] = np.stack([r.medium for r in refs])
    cin = 3 if with_agent_channel else 2
    weights = _weights(kernel_sizes, cin, 3)
    coefs = (0.05, 0.05, 0.7)
    for use_cells in (False, True):
        if use_cells:                                    # after a step the env's cell cache is valid for its agents
            act0 = S.brownian_forward(env.agents, move_scale=0.02, seed=4, step=0)
            env.step(act0)
            for b, r in enumerate(refs):
                r.step(np.asarray(act0).reshape(env.B, 3, env.M)[b])
        action, sense = S.conv_policy_forward(env, weights, coefs, with_agent_channel, use_cells=use_cells)
        for b, r in enumerate(refs):
            oa = R.NeuralAutomataAgent(weights, scale=coefs[0], deposit=coefs[2], with_agent_channel=with_agent_channel)
            oact = oa.forward(r._get_current_obs)
            np.testing.assert_allclose(sense[b], oa.sense_output.numpy()[0], rtol=1e-5, atol=5e-6)
            np.testing.assert_allclose(action[b], oact.astype(np.float64), rtol=1e-5, atol=5e-6)
            assert np.array_equal(action[b].astype(np.float32).astype(np.float64), action[b]), "float32 values in the float64 action"
