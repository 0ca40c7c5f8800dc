"""The reference-facing managers (abmarl_b200.managers) on the GPU: replay the reference transcripts through
AllStepManager / TurnBasedManager and read the results back in the reference's dict-of-dicts form."""
import os

import numpy as np
import pytest
import torch

from abmarl_b200 import _capi as K
from tests import scenarios

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', ['tb_c2', 'tb_blocking', 'maze_c1', 'mm_tiny', 'mm_c4', 'mm_tiny_allstep', 'tb_blocking+specialize', 'mm_c4+specialize'])
def test_manager_replays_reference_transcript(mirror, name):
    name, _, specialize = name.partition('+')          # specialize=True: the step kernel compiled for this sim alone (bgw_specialize)
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    builder, manager, _ = scenarios.SCENARIOS[name]
    cls = {'all_step': mirror.managers.AllStepManager, 'turn_based': mirror.managers.TurnBasedManager}[manager]
    mgr = cls(builder(mirror), n_envs=1, seed=int(g['seed']), device='cuda:0', specialize=bool(specialize))
    ids = mgr.learner_ids
    n_checked = 0
    for t in range(len(g['kind'])):
        present = g['obs_present'][t]
        if g['kind'][t] == 0:
            mgr.reset()
            obs = mgr.as_dicts(0, after_reset=True)
        else:
            mgr.step(torch.from_numpy(g['actions'][t][None].copy()).cuda())
            obs, rew, done, info = mgr.as_dicts(0)
            valid = (g['done'][t] & K.OUT_VALID) != 0
            assert set(rew) == {ids[l] for l in np.nonzero(valid)[0]}          # exactly the agents the reference reported
            for l in np.nonzero(valid)[0]:
                assert abs(rew[ids[l]] - g['reward'][t][l]) <= 1e-6
                assert done[ids[l]] == bool(g['done'][t][l] & K.OUT_DONE)
            assert done['__all__'] == bool(g['all_done'][t])
        for l in np.nonzero(present)[0]:
            (key, arr), = obs[ids[l]].items()
            np.testing.assert_array_equal(arr.ravel(), g['obs'][t][l][:arr.size])
            n_checked += 1
    assert n_checked > 0


@pytest.mark.parametrize('name', ['tb_c2', 'tb_blocking', 'mm_c4'])
def test_manager_get_obs_repeats_the_reported_observations(mirror, name):
    """manager.get_obs(agent_id) = the reference's sim.get_obs(agent_id) (smart.py:93-99): for every learner the
    transcript reports at a step, the same observation the reference returned there; nothing is stepped by asking."""
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    builder, manager, _ = scenarios.SCENARIOS[name]
    cls = {'all_step': mirror.managers.AllStepManager, 'turn_based': mirror.managers.TurnBasedManager}[manager]
    mgr = cls(builder(mirror), n_envs=1, seed=int(g['seed']), device='cuda:0')
    ids = mgr.learner_ids
    n_checked = 0
    for t in range(min(len(g['kind']), 40)):
        if g['kind'][t] == 0:
            mgr.reset()
        else:
            mgr.step(torch.from_numpy(g['actions'][t][None].copy()).cuda())
        steps_before = int(mgr.engine.state['step'][0].item())
        for l in np.nonzero(g['obs_present'][t])[0][:4]:
            got = mgr.get_obs(ids[l])
            (key, arr), = got.items()
            np.testing.assert_array_equal(arr.ravel(), g['obs'][t][l][:arr.size])
            n_checked += 1
        assert int(mgr.engine.state['step'][0].item()) == steps_before
    assert n_checked > 20


def test_encode_actions_matches_reference_dicts(mirror):
    mgr = mirror.managers.AllStepManager(scenarios.build_tb_c2(mirror), n_envs=2, seed=1, device='cuda:0')
    act = mgr.encode_actions([{'agent0': {'move': np.array([1, -1]), 'attack': 1}}, {'agent3': {'move': np.array([0, 1]), 'attack': 0}}])
    a = act.cpu().numpy()
    assert list(a[0, 0]) == [1, -1, 1, 0] and list(a[1, 3]) == [0, 1, 0, 0] and a.sum() == 2
    mgr.reset()
    obs, rew, done, all_done = mgr.step(act)
    assert obs.shape == (2, 24, 7, 7) and rew.shape == (2, 24) and done.shape == (2, 24) and all_done.shape == (2,)
