"""SuperAgentWrapper (abmarl/sim/wrappers/super_agent_wrapper.py) as a view over the batched outputs: replay the
transcripts recorded from the UNMODIFIED reference wrapper (tests/golden/super_*.npz, make_super_golden.py) through the
oracle (CPU) or the engine (GPU) and abmarl_b200.sim.wrappers.SuperAgentView."""
import os

import numpy as np
import pytest
import torch

from abmarl_b200 import _capi as K
from abmarl_b200.spec import compile_sim
from abmarl_b200.sim.wrappers import SuperAgentWrapper, SuperAgentView
from tests import scenarios
from tests.golden.make_super_golden import mapping_for

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = {'super_tb_dense': scenarios.build_tb_dense, 'super_tb_c2': scenarios.build_tb_c2}


def _replay(name, backend, mirror):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    sim = CASES[name](mirror)
    wrapped = SuperAgentWrapper(sim, super_agent_mapping=mapping_for(sim))
    assert [gid for gid in wrapped.agents] == list(g['group_ids'])                 # super agents first, then the uncovered
    spec = compile_sim(sim, manager='all_step', n_envs=1, seed=int(g['seed']), auto_reset=False)
    if backend == 'oracle':
        from oracle.oracle import OracleEnv
        env, dev = OracleEnv(spec), 'cpu'
        tens = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    else:
        from abmarl_b200.engine import BatchedGridWorld
        env, dev = BatchedGridWorld(spec, device='cuda:0'), 'cuda:0'
        tens = lambda a: a
    view = SuperAgentView(spec, wrapped.super_agent_mapping, 1, device=dev)
    assert [gid for gid, _ in view.groups] == list(g['group_ids'])
    checked = 0
    for t in range(len(g['kind'])):
        present = g['obs_present'][t]
        if g['kind'][t] == 0:
            env.reset()
            view.reset()
            obs = tens(env.obs).cpu().numpy()[0]
            np.testing.assert_array_equal(obs[present], g['obs'][t][present], err_msg=f'{name} reset {t}')
            assert g['mask'][t][present].all()
            continue
        act = g['actions'][t][None].copy()
        env.step(act if backend == 'oracle' else torch.from_numpy(act).to(dev), view.order.cpu().numpy() if backend == 'oracle' else view.order)
        obs, mask, reward, done, valid = view.update(tens(env.obs).to(dev), tens(env.reward).to(dev), tens(env.done).to(dev),
                                                    tens(env.all_done).to(dev))
        np.testing.assert_array_equal(valid.cpu().numpy()[0], g['valid'][t], err_msg=f'{name} call {t} which groups report')
        v = g['valid'][t]
        np.testing.assert_array_equal(done.cpu().numpy()[0][v], g['done'][t][v], err_msg=f'{name} call {t} done')
        np.testing.assert_allclose(reward.cpu().numpy()[0][v], g['reward'][t][v], rtol=0, atol=1e-6, err_msg=f'{name} call {t} reward')
        np.testing.assert_array_equal(obs.cpu().numpy()[0][present], g['obs'][t][present], err_msg=f'{name} call {t} obs')
        np.testing.assert_array_equal(mask.cpu().numpy()[0][present], g['mask'][t][present], err_msg=f'{name} call {t} mask')
        checked += int(v.sum())
    assert checked > 100


@pytest.mark.parametrize('name', list(CASES))
def test_super_agent_view_on_the_oracle(mirror, name):
    _replay(name, 'oracle', mirror)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(CASES))
def test_super_agent_view_on_the_engine(mirror, name):
    _replay(name, 'engine', mirror)


@pytest.mark.gpu
def test_manager_over_a_super_agent_wrapper(mirror):
    """AllStepManager(SuperAgentWrapper(sim, mapping), n_envs=...): the inner sim is compiled, super_outputs() regroups."""
    from abmarl_b200.managers import AllStepManager
    sim = scenarios.build_tb_c2(mirror)
    mgr = AllStepManager(SuperAgentWrapper(sim, super_agent_mapping=mapping_for(sim)), n_envs=32, seed=3, horizon=40, auto_reset=True,
                         device='cuda:0')
    mgr.reset()
    for _ in range(60):
        _, reward, done, all_done = mgr.step(mgr.sample_actions())
        obs, mask, r, d, valid = mgr.super_outputs()
        got = r.sum(dim=1).cpu().numpy()
        want = torch.where((done & K.OUT_VALID) != 0, reward, torch.zeros_like(reward)).double().sum(dim=1).cpu().numpy()
        np.testing.assert_allclose(got, want, atol=1e-9)                           # regrouping conserves the reward
        assert (mask | (obs[..., 0] == -2) | ((done & K.OUT_DONE) != 0)).all()     # a masked row is null or just finished
        assert mask[:, mgr.super_view.uncovered].all()
