"""CPU parity: the C oracle replays the transcripts recorded from the UNMODIFIED reference
(tests/golden/*.npz, made by tests/golden/make_golden.py with the reference's numpy draws replaced by the keyed
Philox stream) and must reproduce every array; the mirror API must compile to the spec the reference compiled to."""
import os

import numpy as np
import pytest

from abmarl_b200 import _capi as K
from abmarl_b200.spec import compile_sim, CompiledSpec
from oracle.oracle import OracleEnv
from tests import scenarios

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', list(scenarios.SCENARIOS))
def test_mirror_spec_equals_reference_spec(mirror, name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    builder, manager, _ = scenarios.SCENARIOS[name]
    spec = compile_sim(builder(mirror), manager=manager, n_envs=1, seed=int(g['seed']))
    for s in CompiledSpec.SCALARS:
        want = int(g['spec_' + s]) if 'spec_' + s in g.files else 0      # (fields newer than the fixture are 0 in it)
        assert int(getattr(spec, s)) == want, s
    for t, _ in CompiledSpec.TABLES:
        np.testing.assert_array_equal(getattr(spec, t), g['spec_' + t], err_msg=t)
    for t in ('overlap', 'attack_map', 'reward'):
        np.testing.assert_array_equal(getattr(spec, t), g['spec_' + t], err_msg=t)
    assert spec.agent_ids == list(g['agent_ids'])


@pytest.mark.parametrize('name', list(scenarios.SCENARIOS))
def test_oracle_reproduces_reference_transcript(mirror, name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    builder, manager, _ = scenarios.SCENARIOS[name]
    spec = compile_sim(builder(mirror), manager=manager, n_envs=1, seed=int(g['seed']), auto_reset=False)
    ora = OracleEnv(spec)
    n_valid, episode = 0, -1
    for t in range(len(g['kind'])):
        present = g['obs_present'][t]
        if g['kind'][t] == 0:
            episode += 1
            if spec.layout_generator is not None:
                from abmarl_b200.layouts import layouts_for
                ora.set_layout(layouts_for(spec, [0], [episode]))
            ora.reset()
        else:
            ora.step(g['actions'][t][None])
            np.testing.assert_array_equal(ora.done[0], g['done'][t], err_msg=f'call {t} done')
            np.testing.assert_array_equal(ora.reward64[0], g['reward'][t], err_msg=f'call {t} reward (float64, exact)')
            np.testing.assert_allclose(ora.reward[0], g['reward'][t], rtol=0, atol=1e-6)
            assert int(ora.all_done[0] & K.ENV_ALL_DONE) == int(g['all_done'][t])
            n_valid += int(present.sum())
        np.testing.assert_array_equal(ora.obs[0][present], g['obs'][t][present], err_msg=f'call {t} obs')
        # bgwo_observe (the checker of bgw_observe) = the reference's sim.get_obs(agent_id) on the standing state (smart.py:93-99):
        # for every learner the transcript reports at this call it is the observation the reference returned
        np.testing.assert_array_equal(ora.observe(0)[present], g['obs'][t][present], err_msg=f'call {t} observe')
        st = ora.state
        np.testing.assert_array_equal(st['flags'][0], g['flags'][t], err_msg=f'call {t} flags')
        np.testing.assert_array_equal(st['cell'][0], g['cell'][t], err_msg=f'call {t} cell')
        np.testing.assert_array_equal(st['health'][0], g['health'][t], err_msg=f'call {t} health')
        np.testing.assert_array_equal(st['ammo'][0], g['ammo'][t], err_msg=f'call {t} ammo')
        in_grid = (g['flags'][t] & K.ST_IN_GRID) != 0
        np.testing.assert_array_equal(st['next'][0][in_grid], g['next'][t][in_grid], err_msg=f'call {t} next')
    assert n_valid > 0


def test_c1_first_observation_known_answer(mirror):
    """maze.txt, navigator at (1,4), view 2 -- out-of-bounds cells behind a wall are -2."""
    spec = compile_sim(scenarios.build_maze_c1(mirror), n_envs=1)
    ora = OracleEnv(spec)
    ora.reset()
    want = np.array([[-1, -2, -2, -2, -1], [0, 0, 2, 0, 2], [2, 0, 1, 0, 0], [-2, 2, 0, 2, -2], [-2, -2, 0, -2, -2]])
    np.testing.assert_array_equal(ora.obs_view()[0, 0], want)
