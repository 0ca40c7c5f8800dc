"""The trainer-facing adapters (abmarl_b200/external: the RLlib MultiAgentEnv / gym Env protocols of
abmarl/external/rllib_multiagentenv_wrapper.py:9-51 and gym_env_wrapper.py:6-70) driven with reference-style action dicts:
against the transcripts recorded from the unmodified reference, and against the oracle for batches."""
import os

import numpy as np
import pytest
import torch

from abmarl_b200 import _capi as K
from tests import scenarios

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def action_dict_of(spec, act_rows, acting):
    """int8 action rows [L, stride] -> {agent_id: {'move': ..., 'attack': ...}} for the learners in `acting`, in the
    reference's formats (actor.py:63-65 Box move, wrapper.py:180 ravelled move, actor.py:452 binary attack)."""
    out = {}
    for l in acting:
        a = spec.learner_agents[l]
        d = {}
        if spec.klass[a] & K.AG_MOVING:
            if spec.move_actor == K.MOVE_BOX and not spec.ravel_actions:
                d['move'] = np.array([int(act_rows[l, 0]), int(act_rows[l, 1])])
            else:
                d['move'] = int(np.uint8(act_rows[l, 0]))
        if spec.klass[a] & K.AG_ATTACKING and spec.attack_actor == K.ATTACK_BINARY:
            d['attack'] = int(act_rows[l, 2])
        out[spec.learner_ids[l]] = d
    return out


def rows_of(spec, obs_dict, stride):
    rows, present = np.zeros((spec.n_learners, stride), np.int8), np.zeros(spec.n_learners, bool)
    for l, aid in enumerate(spec.learner_ids):
        if aid in obs_dict:
            (key, arr), = obs_dict[aid].items()
            flat = np.asarray(arr).astype(np.int8).ravel()
            rows[l, :flat.size] = flat
            present[l] = True
    return rows, present


@pytest.mark.parametrize('name', ['tb_c2', 'tb_dense', 'mm_allstep', 'mm_tiny', 'mm_dynamic'])
def test_multi_agent_wrapper_replays_reference_transcript(mirror, name):
    """MultiAgentWrapper.reset / step with the reference's action dicts returns the dicts the reference's manager returned."""
    from abmarl_b200.external import MultiAgentWrapper
    from abmarl_b200 import managers
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    builder, manager, _ = scenarios.SCENARIOS[name]
    cls = {'all_step': managers.AllStepManager, 'turn_based': managers.TurnBasedManager, 'dynamic_order': managers.DynamicOrderManager}[manager]
    mgr = cls(builder(mirror), n_envs=1, seed=int(g['seed']), auto_reset=False)
    if mgr.spec.layout_generator is not None and not mgr.engine.device_layouts:
        pytest.skip('host-side layouts')
    env = MultiAgentWrapper(mgr)
    spec, stride = mgr.spec, mgr.engine.dims.obs_stride
    assert set(env.observation_space.keys()) == set(spec.learner_ids) == env._agent_ids
    reported = set()
    for t in range(len(g['kind'])):
        if g['kind'][t] == 0:
            obs = env.reset()
            reported = set()
        else:
            if manager == 'all_step':
                acting = [l for l in range(spec.n_learners) if l not in reported]
            else:
                acting = [int(mgr.turn[0].item())]
            obs, rew, done, info = env.step(action_dict_of(spec, g['actions'][t], acting))
            for l, aid in enumerate(spec.learner_ids):
                if g['done'][t][l] & K.OUT_VALID:
                    assert abs(rew[aid] - g['reward'][t][l]) <= 1e-6 and done[aid] == bool(g['done'][t][l] & K.OUT_DONE), (t, aid)
                    assert info[aid] == {}
                    if done[aid]:
                        reported.add(l)
                else:
                    assert aid not in rew and aid not in done
            assert done['__all__'] == bool(g['all_done'][t])
        rows, present = rows_of(spec, obs, stride)
        np.testing.assert_array_equal(present, g['obs_present'][t], err_msg=f'{name} call {t}: who observes')
        np.testing.assert_array_equal(rows[present], g['obs'][t][present], err_msg=f'{name} call {t} obs')


def test_vector_multi_agent_env_against_the_oracle(mirror):
    """VectorMultiAgentEnv: E envs stepped with one action dict per env, reset_at for finished envs."""
    from abmarl_b200.external import VectorMultiAgentEnv
    from abmarl_b200 import managers
    from abmarl_b200.spec import compile_sim
    from oracle.oracle import OracleEnv
    E = 6
    mgr = managers.AllStepManager(scenarios.build_tb_dense(mirror), n_envs=E, seed=5, env_offset=2, horizon=12, auto_reset=False)
    ora = OracleEnv(compile_sim(scenarios.build_tb_dense(mirror), n_envs=E, seed=5, env_offset=2, horizon=12, auto_reset=False))
    vec = VectorMultiAgentEnv(mgr)
    spec, stride = mgr.spec, mgr.engine.dims.obs_stride
    assert len(vec.get_sub_environments()) == E
    obs = vec.vector_reset()
    ora.reset()
    reported = [set() for _ in range(E)]
    for e in range(E):
        rows, present = rows_of(spec, obs[e], stride)
        assert present.all()
        np.testing.assert_array_equal(rows, ora.obs[e])
    for t in range(40):
        act = ora.sample_actions()
        dicts = [action_dict_of(spec, act[e], [l for l in range(spec.n_learners) if l not in reported[e]]) for e in range(E)]
        finished = (ora.all_done & K.ENV_ALL_DONE) != 0
        if finished.any():                                   # explicit per-env resets, as RLlib's sampler does
            for e in np.nonzero(finished)[0]:
                o = vec.reset_at(int(e))
                reported[e] = set()
            ora.reset(finished.astype(np.uint8))
            for e in np.nonzero(finished)[0]:
                rows, present = rows_of(spec, vec.sim.dicts_from(vec._snap, int(e), after_reset=True), stride)
                np.testing.assert_array_equal(rows, ora.obs[e])
            act = ora.sample_actions()
            dicts = [action_dict_of(spec, act[e], [l for l in range(spec.n_learners) if l not in reported[e]]) for e in range(E)]
        obs, rew, done, info = vec.vector_step(dicts)
        ora.step(act)
        for e in range(E):
            rows, present = rows_of(spec, obs[e], stride)
            valid = (ora.done[e] & K.OUT_VALID) != 0
            np.testing.assert_array_equal(present, valid)
            np.testing.assert_array_equal(rows[present], ora.obs[e][valid])
            for l, aid in enumerate(spec.learner_ids):
                if valid[l]:
                    assert abs(rew[e][aid] - ora.reward64[e][l]) <= 1e-6
                    assert done[e][aid] == bool(ora.done[e][l] & K.OUT_DONE)
                    if done[e][aid]:
                        reported[e].add(l)
            assert done[e]['__all__'] == bool(ora.all_done[e] & K.ENV_ALL_DONE)


def test_gym_wrappers_on_the_maze(mirror):
    """GymWrapper replays the reference transcript of BASELINE configs[0] (one learning agent); GymVectorEnv steps a batch
    with arrays and resets finished envs in place."""
    from abmarl_b200.external import GymWrapper, GymVectorEnv
    from abmarl_b200 import managers
    from abmarl_b200.spec import compile_sim
    from oracle.oracle import OracleEnv
    g = np.load(os.path.join(GOLDEN, 'maze_c1.npz'))
    mgr = managers.AllStepManager(scenarios.build_maze_c1(mirror), n_envs=1, seed=int(g['seed']), auto_reset=False)
    env = GymWrapper(mgr)
    assert env.agent_id == 'navigator' and env.action_space is env.agent.action_space
    for t in range(len(g['kind'])):
        if g['kind'][t] == 0:
            obs = env.reset()
        else:
            obs, reward, done, info = env.step({'move': np.array([int(g['actions'][t][0, 0]), int(g['actions'][t][0, 1])])})
            assert abs(reward - g['reward'][t][0]) <= 1e-6 and done == bool(g['done'][t][0] & K.OUT_DONE) and info == {}
        (key, arr), = obs.items()
        np.testing.assert_array_equal(np.asarray(arr).astype(np.int8).ravel(), g['obs'][t][0][:arr.size])

    E = 8
    vmgr = managers.AllStepManager(scenarios.build_maze_c1(mirror), n_envs=E, seed=3, horizon=9, auto_reset=False)
    ora = OracleEnv(compile_sim(scenarios.build_maze_c1(mirror), n_envs=E, seed=3, horizon=9, auto_reset=False))
    venv = GymVectorEnv(vmgr)
    obs = venv.reset()
    ora.reset()
    np.testing.assert_array_equal(obs.reshape(E, -1), ora.obs[:, 0, :obs[0].size])
    for t in range(30):
        act = ora.sample_actions()
        obs, reward, done, infos = venv.step(act[:, 0])
        ora.step(act)
        fin = (ora.all_done & K.ENV_ALL_DONE) != 0
        np.testing.assert_array_equal(done, fin)
        np.testing.assert_allclose(reward, ora.reward64[:, 0], atol=1e-6)
        if fin.any():
            ora.reset(fin.astype(np.uint8))
        np.testing.assert_array_equal(obs.reshape(E, -1), ora.obs[:, 0, :obs[0].size])
