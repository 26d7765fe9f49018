"""die_b200.NeuralAutomataAgent / ConvolutionModel on the GPU (core/agent/evo.py:45-209).

Parity: against the oracle's restatement (torch's own conv2d on the CPU; equal to the reference's code executed live,
tests/test_golden_oracle.py) on the same weights and observation, per step, with the env stepped by the GPU's action on
both sides: model output and action within 1e-5 relative + 5e-6 absolute (float32 sums in another order), everything
Env.step computes from the SAME action bit-exact except agent_food / reward (the reference evaluates the action cost in
float32 when the action is float32: 1e-6).  The reference's own unit tests (test/unit/agent.py) are mirrored at the end."""
import io

import numpy as np
import pytest

from tests._parity import make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field,kernel_sizes,with_agent_channel,batch", [((256, 256), (3,), True, None), ((96, 130), (3, 5), True, 3),
                                                                        ((64, 64), (5, 3, 3), False, None), ((40, 72), (7,), True, 2)])
def test_neural_automata_agent_against_the_oracle(field, kernel_sizes, with_agent_channel, batch):
    import torch
    import die_b200 as D
    from oracle import die_ref as R
    torch.manual_seed(3)
    refs, env = make_pair(field, seed=6, ratio=0.2, batch=batch)
    B, m = env.batch, env.max_agents
    ga = D.NeuralAutomataAgent(scale=0.05, deposit=0.7, with_agent_channel=with_agent_channel, kernel_sizes=kernel_sizes)
    ga.model.init_weights()
    weights = [w.numpy().copy() for w in ga.model.kernels]
    oas = [R.NeuralAutomataAgent(weights, scale=0.05, deposit=0.7, with_agent_channel=with_agent_channel) for _ in range(B)]
    gobs = env._get_current_obs
    for it in range(8):
        gact = ga.forward(gobs)
        act = gact.cpu().numpy().reshape(B, 3, m)
        sense = ga._sense_output.cpu().numpy()
        for b in range(B):
            oact = oas[b].forward(refs[b]._get_current_obs)
            np.testing.assert_allclose(sense[b], oas[b].sense_output.numpy()[0], rtol=1e-5, atol=5e-6)
            np.testing.assert_allclose(act[b], oact.astype(np.float64), rtol=1e-5, atol=5e-6)
            assert np.array_equal(act[b].astype(np.float32).astype(np.float64), act[b])
            refs[b].step(act[b].astype(np.float32))                # the GPU's action, float32 as the reference's
        gobs, gr, *_ = env.step(gact)
        med, ag = env.get_state()
        med, ag = med.reshape(B, 3, *field), ag.reshape(B, 4, m)
        for b in range(B):
            assert np.array_equal(med[b], refs[b].medium), f"medium differs at step {it}"
            assert np.array_equal(ag[b, :3], refs[b].agents[:3])
            np.testing.assert_allclose(ag[b, 3], refs[b].agents[3], rtol=1e-6, atol=1e-7)
    assert it == 7 and (it == 0 or True)


def test_hints_and_host_path_give_the_same_action():
    import torch
    import die_b200 as D
    torch.manual_seed(1)
    _, env = make_pair((48, 64), seed=2, ratio=0.2)
    ga = D.NeuralAutomataAgent(kernel_sizes=(3, 3))
    ga.model.init_weights()
    obs, *_ = env.step(D.BrownianAgent(0.02).forward(env._get_current_obs))
    a_hint = ga.forward(obs).clone()                               # the env's cell cache
    ga.use_env_hints = False
    a_plain = ga.forward(obs).clone()                              # cells resolved from the positions
    a_host = ga.forward(tuple(t.cpu().numpy() for t in obs))       # numpy in, numpy out
    assert torch.equal(a_hint, a_plain) and np.array_equal(a_host, a_plain.cpu().numpy())


# ---- the reference's own unit tests (test/unit/agent.py), on this implementation -------------------------------------
kernel_sizes_test = ((3,), (5,), (3, 3), (3, 5), (3, 5, 3),)


def test_convolution_model_init():
    import die_b200 as D
    model = D.ConvolutionModel(num_act_channels=3, num_obs_channels=3, kernel_sizes=(3, 5), p_agent_dropout=0.)
    assert not all(bool((k == 0).all()) for k in model.kernels)


@pytest.mark.parametrize('field_size', [(12, 12), (96, 96), (12, 8)])
@pytest.mark.parametrize('kernel_sizes', kernel_sizes_test)
def test_convolution_model_apply(field_size, kernel_sizes):
    import torch as th
    import die_b200 as D
    input_data = th.rand((1, 3, *field_size))
    model = D.ConvolutionModel(num_act_channels=2, num_obs_channels=3, kernel_sizes=kernel_sizes, p_agent_dropout=0.)
    model.init_weights()
    output = model.forward(input_data.cuda()).cpu()
    assert th.any(input_data[0, 0, :, :] != output[0, 0, :, :])
    assert 2 in output.shape
    # and the numbers are torch's: the same weights through torch.nn.functional.conv2d on the CPU
    x = input_data
    for w in model.kernels:
        p = w.shape[-1] // 2
        x = th.nn.functional.conv2d(th.nn.functional.pad(x, (p, p, p, p), mode='circular'), w)
    assert th.allclose(output, th.tanh(x), rtol=1e-5, atol=5e-6)


def test_serialize():
    import torch as th
    import die_b200 as D
    agent = D.NeuralAutomataAgent(kernel_sizes=(3, 5))
    buf = io.BytesIO()
    agent.save(buf)
    buf.seek(0)
    agent2 = D.NeuralAutomataAgent.load(buf)
    assert agent.init_params == agent2.init_params
    input_data = th.rand((1, 3, 16, 12)).cuda()
    assert th.allclose(agent.model.forward(input_data), agent2.model.forward(input_data))
    # the file is the reference's format: params_dict + model_state with its state_dict keys
    buf.seek(0)
    loaded = th.load(buf, weights_only=False)
    assert set(loaded) == {'params_dict', 'model_state'} and set(loaded['model_state']) == {'kernels.0.weight', 'kernels.1.weight'}


def test_population_evaluator_equals_one_candidate_at_a_time():
    """die_b200.PopulationEvaluator: P candidates scored at once on a batched env (one model per environment, CUDA-graph
    replay, rewards summed on the device) == the eager batched loop == every candidate alone on a single env started
    from the same state -- bit for bit; and the oracle's NeuralAutomataAgent + Env within the float32 tolerance."""
    import torch
    import die_b200 as D
    from oracle import die_ref as R
    from tests._parity import make_pair
    field, P, iters = (48, 40), 4, 9
    kw = dict(kernel_sizes=[3, 3], scale=0.01, deposit=2.0)
    rng = np.random.default_rng(3)
    agent = D.NeuralAutomataAgent(**kw)
    cands = rng.uniform(-0.5, 0.5, size=(P, agent.model.num_parameters)).astype(np.float32)
    dyn = dict(food_infinite=True)
    scores = []
    for graph in (True, False):
        refs, env = make_pair(field, seed=31, ratio=0.15, dynamics_kw=dyn, batch=P)
        ev = D.PopulationEvaluator(env, D.NeuralAutomataAgent(**kw), graph=graph)
        if not graph:
            ev.evaluate(cands, 2)                                    # the capture's two eager warm-up iterations
        s1 = ev.evaluate(cands, iters)
        s2 = ev.evaluate(cands[::-1].copy(), iters + 1)              # continues where the first left off; odd count
        scores.append((s1, s2, *env.get_state()))
    for a, b in zip(scores[0], scores[1]):
        assert np.array_equal(a, b)
    # every candidate alone, from the same initial state as its environment of the batch:
    refs, env = make_pair(field, seed=31, ratio=0.15, dynamics_kw=dyn, batch=P)
    singles, osum = [], []
    for p in range(P):
        (ref,), one = make_pair(field, seed=31 + p, ratio=0.15, dynamics_kw=dyn)
        ag = D.NeuralAutomataAgent(**kw)
        ag.model.set_parameters_vector(cands[p])
        oa = R.NeuralAutomataAgent([w.numpy() for w in ag.model.kernels], scale=0.01, deposit=2.0)
        obs, total, ototal = one._get_current_obs, 0.0, 0.0
        robs = ref._get_current_obs
        for _ in range(iters):
            obs, r, *_ = one.step(ag.forward(obs))
            total += r
            robs, rr, *_ = ref.step(oa.forward(robs))
            ototal += rr
        singles.append(total)
        osum.append(ototal)
    ev = D.PopulationEvaluator(env, D.NeuralAutomataAgent(**kw), graph=False)
    batch_scores = ev.evaluate(cands, iters)
    assert np.array_equal(batch_scores, np.array(singles))
    np.testing.assert_allclose(batch_scores, np.array(osum), rtol=2e-4)
    with pytest.raises(ValueError):
        ev.evaluate(cands[:2], 3)


def test_pgpe_improves_a_population_on_the_gpu():
    """A few generations of the search the reference's examples/learning_agents.py configures (PGPE, popsize 10, infinite
    food): the best score of the last generations exceeds the first generation's mean."""
    import die_b200 as D
    env = D.Env((64, 64), D.Dynamics(init_agent_ratio=0.15, food_infinite=True), init='device', seed=5, batch=10)
    agent = D.NeuralAutomataAgent(kernel_sizes=[3, 3], scale=0.01, deposit=2.0)
    ev = D.PopulationEvaluator(env, agent)
    es = D.PGPE(agent.model.num_parameters, popsize=10, radius_init=1.5, seed=1)
    history = []
    for gen in range(12):
        cand = es.ask()
        fit = ev.evaluate(cand, 20)
        assert np.isfinite(fit).all()
        es.tell(fit)
        history.append(fit)
    assert es.best[1] is not None and ev.evaluations == 12
    assert max(h.max() for h in history[-4:]) > history[0].mean()
