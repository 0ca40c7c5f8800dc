"""tests/golden/super_*.npz: the UNMODIFIED reference's SuperAgentWrapper (abmarl/sim/wrappers/super_agent_wrapper.py)
over a team battle under AllStepManager, draws replayed from the keyed Philox stream (build container only).

    python tests/golden/make_super_golden.py

Per manager call the transcript holds, for every learner of the inner sim, the observation row its super agent (or the
agent itself when uncovered) reported for it, the mask bit, and per group (super agents first, then the uncovered
learners) reward, done and whether the group was reported at all.  tests/test_super_agent.py replays the actions
through the oracle (CPU) / the engine (GPU) + abmarl_b200.sim.wrappers.SuperAgentView and compares.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from abmarl_b200 import _capi as K                      # noqa: E402
from abmarl_b200.spec import compile_sim                # noqa: E402
from oracle.oracle import OracleEnv                     # noqa: E402
from oracle.refshim import PhiloxReplay                 # noqa: E402
from tests import scenarios                             # noqa: E402
from tests.golden.make_golden import action_dict        # noqa: E402

SEED, OUT = 0xB200, os.path.dirname(os.path.abspath(__file__))


def mapping_for(sim):
    """teams 1 and 2 become super agents, team 3 stays uncovered (exercises both branches of the wrapper)"""
    return {f'team{t}': [a.id for a in sim.agents.values() if a.encoding == t] for t in (1, 2)}


def record(name, builder, n_steps):
    api = scenarios.reference_api()
    from abmarl.sim.wrappers import SuperAgentWrapper
    sim = builder(api)
    mapping = mapping_for(sim)
    mgr = api.managers.AllStepManager(SuperAgentWrapper(sim, super_agent_mapping=mapping))
    spec = compile_sim(sim, manager='all_step', n_envs=1, seed=SEED, auto_reset=False)
    ora = OracleEnv(spec)
    L, stride = spec.n_learners, ora.dims.obs_stride
    lid = {aid: l for l, aid in enumerate(spec.learner_ids)}
    covered = {m for ms in mapping.values() for m in ms}
    groups = list(mapping.items()) + [(aid, [aid]) for aid in spec.learner_ids if aid not in covered]
    rec = {k: [] for k in ('kind', 'actions', 'obs', 'obs_present', 'mask', 'reward', 'done', 'valid', 'all_done')}

    def rows_of(ref_obs):
        rows = np.zeros((L, stride), np.int8)
        present, mask = np.zeros(L, bool), np.zeros(L, bool)
        for gid, members in groups:
            if gid not in ref_obs:
                continue
            for m in members:
                o = ref_obs[gid][m] if gid in mapping else ref_obs[gid]
                (key, arr), = o.items()
                flat = np.asarray(arr).astype(np.int8).ravel()
                rows[lid[m], :flat.size] = flat
                present[lid[m]] = True
                mask[lid[m]] = bool(ref_obs[gid]['mask'][m][0]) if gid in mapping else True
        return rows, present, mask

    with PhiloxReplay(sim, SEED) as rp:
        t, need_reset = 0, True
        while t < n_steps:
            if need_reset:
                rp.episode += 1
                rp.step = 0
                rows, present, mask = rows_of(mgr.reset())
                ora.reset()
                for k, v in (('kind', 0), ('actions', np.zeros((L, ora.dims.action_stride), np.int8)), ('obs', rows), ('obs_present', present),
                             ('mask', mask), ('reward', np.zeros(len(groups))), ('done', np.zeros(len(groups), bool)),
                             ('valid', np.zeros(len(groups), bool)), ('all_done', 0)):
                    rec[k].append(v)
                need_reset = False
                continue
            act = ora.sample_actions()[0]
            rp.step += 1
            per_agent = action_dict(spec, sim, set(), act)                  # every learner; the wrapper filters the done ones
            actions = {}
            for gid, members in groups:
                if gid in mgr.done_agents:
                    continue
                actions[gid] = {m: per_agent[m] for m in members} if gid in mapping else per_agent[gid]
            ref_obs, ref_rew, ref_done, _ = mgr.step(actions)
            ora.step(act[None])
            rows, present, mask = rows_of(ref_obs)
            rec['kind'].append(1); rec['actions'].append(act); rec['obs'].append(rows); rec['obs_present'].append(present)
            rec['mask'].append(mask)
            rec['reward'].append(np.array([ref_rew.get(g, 0.0) for g, _ in groups], dtype=np.float64))
            rec['done'].append(np.array([bool(ref_done.get(g, False)) for g, _ in groups]))
            rec['valid'].append(np.array([g in ref_rew for g, _ in groups]))
            rec['all_done'].append(int(bool(ref_done['__all__'])))
            t += 1
            need_reset = bool(ref_done['__all__'])
    out = {k: np.stack(v) for k, v in rec.items()}
    out['seed'] = np.uint64(SEED)
    out['group_ids'] = np.array([g for g, _ in groups])
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    print(f"{name}: {n_steps} steps, {int((out['kind'] == 0).sum())} episodes, {len(groups)} groups, "
          f"{int(out['valid'].sum())} group reports, {int((~out['mask'] & out['obs_present']).sum())} masked rows -> "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == '__main__':
    record('super_tb_dense', scenarios.build_tb_dense, 60)
    record('super_tb_c2', scenarios.build_tb_c2, 60)
