"""FlattenWrapper mirror (abmarl_b200/sim/flatten.py) against golden vectors of the reference's own flatten_wrapper.py
(tests/golden/flatten_golden.json, made by tests/golden/make_flatten_golden.py from the unmodified reference), and the
batched FlattenView against the per-point functions."""
import json
import os

import numpy as np
import pytest

from abmarl_b200.sim import Agent
from abmarl_b200.sim.flatten import FlattenWrapper, FlattenActionWrapper, flatten, unflatten, flatdim, flatten_space
from abmarl_b200.spaces import Box, Discrete, MultiDiscrete, Dict
from tests import scenarios

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flatten_golden.json')))


def _point(space, x):
    """JSON value -> a point of `space` (arrays for Box / MultiDiscrete, int keys for the encoding-keyed attack Dict)."""
    if isinstance(space, Dict):
        return {k: _point(s, x[str(k)]) for k, s in space.spaces.items()}
    if isinstance(space, (Box, MultiDiscrete)):
        return np.asarray(x)
    return x


def _same(a, b):
    if isinstance(a, dict):
        return set(map(str, a)) == set(map(str, b)) and all(_same(v, b[k] if k in b else b[str(k)]) for k, v in a.items())
    return np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize('name', sorted(GOLDEN))
def test_flatten_wrapper_matches_reference(mirror, name):
    sim = scenarios.SCENARIOS[name][0](mirror)
    # The reference builds its observers from a SET of classes (smart.py:53-62), so the key order of a two-observer
    # observation Dict (grid + ammo) is the set's iteration order: unspecified.  Put the mirror's Dict in the order the
    # golden run happened to have before wrapping.
    for agent_id, rec in GOLDEN[name].items():
        order = list(rec['points'][0]['obs'].keys())
        space = sim.agents[agent_id].observation_space
        if list(space.spaces.keys()) != order:
            assert sorted(space.spaces.keys()) == sorted(order)
            sim.agents[agent_id].observation_space = Dict({k: space[k] for k in order})
    wrapped = FlattenWrapper(sim)
    for agent_id, rec in GOLDEN[name].items():
        agent, w = sim.agents[agent_id], wrapped.agents[agent_id]
        assert isinstance(agent, Agent)
        assert flatdim(agent.observation_space) == rec['obs_dim'] and flatdim(agent.action_space) == rec['act_dim']
        for sp, lo, hi, dt in ((w.observation_space, rec['obs_low'], rec['obs_high'], rec['obs_dtype']),
                               (w.action_space, rec['act_low'], rec['act_high'], rec['act_dtype'])):
            assert isinstance(sp, Box) and np.array_equal(sp.low, lo) and np.array_equal(sp.high, hi)
            assert np.dtype(sp.dtype) == np.dtype(dt)
        if rec['null_observation'] is not None:
            assert np.array_equal(w.null_observation, rec['null_observation'])
        for p in rec['points']:
            obs, act = _point(agent.observation_space, p['obs']), _point(agent.action_space, p['action'])
            assert np.array_equal(wrapped.wrap_observation(agent, obs), p['flat_obs'])
            assert np.array_equal(wrapped.unwrap_action(agent, act), p['flat_action'])
            assert _same(wrapped.wrap_action(agent, np.asarray(p['flat_action'])), p['unflat_action'])
            assert _same(wrapped.unwrap_observation(agent, np.asarray(p['flat_obs'])), obs)
        assert sim.agents[agent_id].action_space is not w.action_space            # the wrapped sim keeps its own spaces
    aw = FlattenActionWrapper(sim)
    for agent_id, rec in GOLDEN[name].items():
        assert np.array_equal(aw.agents[agent_id].action_space.high, rec['act_high'])
        assert aw.agents[agent_id].observation_space == sim.agents[agent_id].observation_space


def test_flatten_functions_on_plain_spaces():
    d = Dict({'a': Discrete(4), 'b': Box(-1.5, 2.5, (2, 2), float), 'c': MultiDiscrete([3, 5])})
    fs = flatten_space(d)
    assert fs.dtype == float and fs.shape == (7,)
    assert np.array_equal(fs.low, [0, -1.5, -1.5, -1.5, -1.5, 0, 0]) and np.array_equal(fs.high, [3, 2.5, 2.5, 2.5, 2.5, 2, 4])
    p = {'a': 2, 'b': np.array([[0.5, -1.0], [2.0, 1.25]]), 'c': np.array([1, 4])}
    f = flatten(d, p)
    assert np.array_equal(f, [2, 0.5, -1.0, 2.0, 1.25, 1, 4])
    back = unflatten(d, f)
    assert back['a'] == 2 and np.array_equal(back['b'], p['b']) and np.array_equal(back['c'], p['c'])
    assert flatten_space(Discrete(5)).dtype == int


def test_every_space_samples_points_it_contains():
    for sp in (Box(-2, 3, (4, 4), int), Box(0.0, 1.0, (3,), float), Discrete(5), MultiDiscrete([26, 26, 3]),
               Dict({'move': Box(-1, 1, (2,), int), 'attack': MultiDiscrete([4, 4])})):
        sp.seed(3)
        for _ in range(5):
            assert sp.contains(sp.sample()), sp


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['tb_c2', 'tb_encoding', 'tb_restricted', 'tb_ammo_selective', 'mm_allstep'])
def test_flatten_view_on_the_engine(mirror, name):
    """Flat actions scatter into the rows encode_actions builds from the reference-style dicts; flat observations are
    flatten() of the dict observations as_dicts returns."""
    import torch
    from abmarl_b200.sim.flatten import FlattenView
    builder, manager, _ = scenarios.SCENARIOS[name]
    sim = builder(mirror)
    cls = mirror.managers.AllStepManager if manager == 'all_step' else mirror.managers.TurnBasedManager
    E = 5
    mgr = cls(sim, n_envs=E, seed=3, horizon=30, auto_reset=True, device='cuda:0')
    view = FlattenView(mgr)
    mgr.reset()
    rng = np.random.default_rng(5)
    for t in range(6):
        dicts, flat = [], np.zeros((E, mgr.n_learners, view.act_dim), dtype=np.int64)
        for e in range(E):
            d = {}
            for l, agent_id in enumerate(mgr.learner_ids):
                agent = sim.agents[agent_id]
                agent.action_space.seed(int(rng.integers(1 << 30)))
                a = agent.action_space.sample()
                d[agent_id] = a
                fa = flatten(agent.action_space, a)
                flat[e, l, :len(fa)] = fa
            dicts.append(d)
        rows = view.encode_actions(flat)
        assert torch.equal(rows, mgr.encode_actions(dicts)), (name, t)
        mgr.step(rows)
        fo = view.observations().cpu().numpy()
        for e in range(E):
            obs, _, _, _ = mgr.as_dicts(e)
            for l, agent_id in enumerate(mgr.learner_ids):
                if agent_id in obs:
                    want = flatten(sim.agents[agent_id].observation_space, obs[agent_id])
                    assert np.array_equal(fo[e, l, :len(want)], want), (name, t, e, agent_id)
