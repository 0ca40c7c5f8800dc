"""Generate tests/golden/flatten_golden.json from the UNMODIFIED reference (build container only).

    python tests/golden/make_flatten_golden.py

For a set of scenarios of tests/scenarios.py the reference sim is built from the reference's own classes and wrapped in
the reference's FlattenWrapper (abmarl/sim/wrappers/flatten_wrapper.py:156-204).  Stored per scenario and learning agent:
the flattened spaces (low / high / dtype), a few sampled points of the original observation and action spaces together
with the reference's flatten() of them, and the reference's unflatten() of the flattened action -- which must give the
point back.  tests/test_flatten.py rebuilds the same sims from the mirror classes and must reproduce every array.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import scenarios                             # noqa: E402

NAMES = ['tb_c2', 'tb_c5_small', 'tb_encoding', 'tb_restricted', 'tb_selective', 'tb_ammo', 'tb_ammo_selective', 'maze_c1',
         'reach_target', 'pacman_simple', 'mm_allstep']
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'flatten_golden.json')


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, np.generic):
        return x.item()
    return x


def main():
    api = scenarios.reference_api()
    from abmarl.sim.wrappers import FlattenWrapper
    from abmarl.sim.wrappers.flatten_wrapper import flatten, unflatten, flatdim
    from abmarl.sim import Agent
    out = {}
    for name in NAMES:
        builder = scenarios.SCENARIOS[name][0]
        sim = builder(api)
        wrapped = FlattenWrapper(sim)
        rec = {}
        for agent_id, agent in sim.agents.items():
            if not isinstance(agent, Agent):
                continue
            w = wrapped.agents[agent_id]
            agent.observation_space.seed(11)
            agent.action_space.seed(12)
            points = []
            for _ in range(3):
                o, a = agent.observation_space.sample(), agent.action_space.sample()
                fo, fa = flatten(agent.observation_space, o), flatten(agent.action_space, a)
                back = unflatten(agent.action_space, fa)
                assert np.array_equal(flatten(agent.action_space, back), fa)
                points.append({'obs': jsonable(o), 'flat_obs': jsonable(np.asarray(fo)), 'action': jsonable(a),
                               'flat_action': jsonable(np.asarray(fa)), 'unflat_action': jsonable(back)})
            rec[agent_id] = {
                'obs_low': w.observation_space.low.tolist(), 'obs_high': w.observation_space.high.tolist(),
                'obs_dtype': str(np.dtype(w.observation_space.dtype)), 'obs_dim': flatdim(agent.observation_space),
                'act_low': w.action_space.low.tolist(), 'act_high': w.action_space.high.tolist(),
                'act_dtype': str(np.dtype(w.action_space.dtype)), 'act_dim': flatdim(agent.action_space),
                'null_observation': None if not np.any(np.asarray(jsonable(w.null_observation) or 0)) and not isinstance(w.null_observation, np.ndarray) else jsonable(np.asarray(w.null_observation)),
                'points': points}
            if len(rec) >= 6:                                    # the agents of a scenario repeat a few shapes
                break
        out[name] = rec
        print(name, len(rec), 'agents')
    with open(OUT, 'w') as f:
        json.dump(out, f)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
