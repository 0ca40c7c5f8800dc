"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py [scenario ...]

For each scenario in tests/scenarios.py the reference sim is built from the reference's own classes, driven
by the reference's own manager with its numpy draws replaced by the keyed Philox stream (oracle/refshim.py),
and the transcript -- actions, observations, float64 rewards, dones, __all__, and the full state after every
call -- is stored.  While recording, the C oracle (oracle/bgw_oracle.c) runs the same episode and every array
is compared; the script fails on the first mismatch, so a committed golden file also certifies the oracle.

The golden files are what travels to the GPU box: tests replay `actions` through the oracle (CPU tests) and
through the CUDA engine (gpu tests) and compare with the recorded reference outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from abmarl_b200 import _capi as K                      # noqa: E402
from abmarl_b200.spec import compile_sim, CompiledSpec  # noqa: E402
from oracle.oracle import OracleEnv                     # noqa: E402
from oracle.refshim import PhiloxReplay, extract_state  # noqa: E402
from abmarl_b200.layouts import layouts_for             # noqa: E402
from tests import scenarios                             # noqa: E402

SEED = 0xB200
OUT = os.path.dirname(os.path.abspath(__file__))


def action_dict(spec, sim, done_agents, act, only=None):
    """bytes [L,4] -> the reference's {agent_id: {'move': ..., 'attack': ...}} for learners not yet done
    (`only` = the learner whose turn it is under TurnBasedManager)."""
    out = {}
    for l, a in enumerate(spec.learner_agents):
        agent_id = spec.agent_ids[a]
        if agent_id in done_agents or (only is not None and l != only):
            continue
        agent = sim.agents[agent_id]
        d = {}
        if 'move' in agent.action_space.spaces:
            if spec.move_actor == K.MOVE_BOX and not spec.ravel_actions:
                d['move'] = np.array([int(act[l, 0]), int(act[l, 1])])
            else:
                d['move'] = int(np.uint8(act[l, 0]))
        if 'attack' in agent.action_space.spaces:
            att = act[l, 2:].view(np.uint8).astype(int)
            n = 2 * int(getattr(agent, 'attack_range', 0)) + 1
            if spec.attack_actor == K.ATTACK_ENCODING:        # {encoding: count}, ascending (actor.py:513-519)
                d['attack'] = {int(enc): int(att[enc - 1]) for enc in sorted(agent.action_space.spaces['attack'].spaces)}
            elif spec.attack_actor == K.ATTACK_RESTRICTED:     # actor.py:593-599
                d['attack'] = att[:agent.simultaneous_attacks].copy()
            elif spec.attack_actor == K.ATTACK_SELECTIVE:      # actor.py:669-679
                d['attack'] = att[:n * n].reshape(n, n).copy()
            else:
                d['attack'] = int(att[0])
        out[agent_id] = d
    return out


def ref_obs_rows(spec, ref_obs, stride, ammo_offset=-1, position_offset=-1):
    """reference obs dict -> (rows [L, stride] int8 zero padded, present [L] bool); the AmmoObserver's 'ammo' entry
    (observer.py:406-413) goes to the int32 slot at ammo_offset"""
    rows = np.zeros((spec.n_learners, stride), dtype=np.int8)
    present = np.zeros(spec.n_learners, dtype=bool)
    for l, a in enumerate(spec.learner_agents):
        agent_id = spec.agent_ids[a]
        if agent_id in ref_obs:
            entries = dict(ref_obs[agent_id])
            if 'ammo' in entries:
                assert ammo_offset >= 0
                rows[l, ammo_offset:ammo_offset + 4] = np.array([entries.pop('ammo')], dtype=np.int32).view(np.int8)
            if 'position' in entries:                 # AbsolutePositionObserver observer.py:366-373 -> two int16
                assert position_offset >= 0
                rows[l, position_offset:position_offset + 4] = np.asarray(entries.pop('position')).astype(np.int16).view(np.int8)
            (key, arr), = entries.items()
            assert arr.min() >= -128 and arr.max() <= 127
            flat = np.asarray(arr).astype(np.int8).ravel()
            rows[l, :flat.size] = flat
            present[l] = True
    return rows, present


def check(name, what, t, got, want):
    if not np.array_equal(got, want):
        bad = np.argwhere(np.asarray(got) != np.asarray(want))[:5]
        raise SystemExit(f"[{name}] oracle != reference: {what} at call {t}; first diffs at {bad.tolist()}\n"
                         f"oracle={np.asarray(got)[tuple(bad[0])]} reference={np.asarray(want)[tuple(bad[0])]}")


def record(name, builder, manager, n_steps):
    api = scenarios.reference_api()
    sim = builder(api)
    mgr = {'all_step': api.managers.AllStepManager, 'turn_based': api.managers.TurnBasedManager,
           'all_step_shuffled': lambda s_: api.managers.AllStepManager(s_, randomize_action_input=True),
           'dynamic_order': api.managers.DynamicOrderManager}[manager](sim)
    spec = compile_sim(sim, manager=manager, n_envs=1, seed=SEED, auto_reset=False)
    ora = OracleEnv(spec)
    L, stride, astride, ammo_off = spec.n_learners, ora.dims.obs_stride, ora.dims.action_stride, ora.dims.ammo_offset
    pos_off = ora.dims.position_offset
    learner_ids = spec.learner_ids

    rec = {k: [] for k in ('kind', 'actions', 'obs', 'obs_present', 'reward', 'done', 'all_done', 'cell', 'next',
                           'flags', 'health', 'ammo')}

    def snapshot(kind, act, obs_rows, present, reward, done, all_done):
        # BgwState marks the entities that never report (non-learners) DONE_REPORTED; AllStepManager / TurnBasedManager keep
        # them in done_agents from reset on (all_step_manager.py:41-44), DynamicOrderManager.reset starts with an empty set
        done_set = set(mgr.done_agents) | ({a for a in sim.agents if a not in learner_ids} if manager == 'dynamic_order' else set())
        st = extract_state(sim, done_set)
        rec['kind'].append(kind)
        rec['actions'].append(act)
        rec['obs'].append(obs_rows)
        rec['obs_present'].append(present)
        rec['reward'].append(reward)
        rec['done'].append(done)
        rec['all_done'].append(all_done)
        for k in ('cell', 'next', 'flags', 'health', 'ammo'):
            rec[k].append(st[k])
        return st

    def compare_state(t, st):
        o = ora.state
        in_grid = (st['flags'] & K.ST_IN_GRID) != 0
        check(name, 'flags', t, o['flags'][0], st['flags'])
        check(name, 'cell', t, o['cell'][0], st['cell'])
        check(name, 'next(in grid)', t, o['next'][0][in_grid], st['next'][in_grid])
        check(name, 'health', t, o['health'][0], st['health'])
        check(name, 'ammo', t, o['ammo'][0], st['ammo'])

    forced = scenarios.GOLDEN_RESET_EVERY.get(name, 0)
    with PhiloxReplay(sim, SEED) as rp:
        t = 0
        need_reset = True
        while t < n_steps:
            if need_reset:
                rp.episode += 1
                rp.step = 0
                if spec.layout_generator:
                    ora.set_layout(layouts_for(spec, [0], [rp.episode]))
                ref_obs = mgr.reset()
                ora.reset()
                rows, present = ref_obs_rows(spec, ref_obs, stride, ammo_off, pos_off)
                st = snapshot(0, np.zeros((L, astride), np.int8), rows, present, np.zeros(L), np.zeros(L, np.uint8), 0)
                compare_state(t, st)
                check(name, 'reset obs', t, ora.obs[0][present], rows[present])
                need_reset = False
                continue
            act = ora.sample_actions()[0]
            rp.step += 1
            only = int(ora.state['turn'][0]) if manager in ('turn_based', 'dynamic_order') else None
            ref_obs, ref_rew, ref_done, _ = mgr.step(action_dict(spec, sim, mgr.done_agents, act, only))
            ora.step(act[None])
            rows, present = ref_obs_rows(spec, ref_obs, stride, ammo_off, pos_off)
            reward = np.zeros(L)
            done = np.zeros(L, np.uint8)
            for l, agent_id in enumerate(learner_ids):
                if agent_id in ref_rew:
                    reward[l] = ref_rew[agent_id]
                    done[l] = K.OUT_VALID | (K.OUT_DONE if ref_done[agent_id] else 0)
            all_done = int(bool(ref_done['__all__']))
            st = snapshot(1, act, rows, present, reward, done, all_done)
            t += 1
            compare_state(t, st)
            check(name, 'done', t, ora.done[0], done)
            check(name, 'obs', t, ora.obs[0][present], rows[present])
            check(name, 'reward64', t, ora.reward64[0], reward)
            check(name, '__all__', t, int(ora.all_done[0] & K.ENV_ALL_DONE), all_done)
            need_reset = bool(all_done) or (forced and t % forced == 0)
        n_draws = len(rp.log)
        n_ammo_draws = sum(1 for site, _, _ in rp.log if site == K.SITE_AMMO)
        n_repeat_acc = sum(1 for site, _, k in rp.log if site == K.SITE_ACC and k >= 4096)
        n_multi_subset = sum(1 for site, _, k in rp.log if site == K.SITE_SUBSET and (k & 0xFF) > 0)

    out = {k: np.stack(v) for k, v in rec.items()}
    out['seed'] = np.uint64(SEED)
    out['agent_ids'] = np.array(spec.agent_ids)
    for s in CompiledSpec.SCALARS:
        out['spec_' + s] = np.int64(getattr(spec, s)) if s != 'seed' else np.uint64(spec.seed)
    for tname, _ in CompiledSpec.TABLES:
        out['spec_' + tname] = getattr(spec, tname)
    out['spec_overlap'], out['spec_attack_map'], out['spec_reward'] = spec.overlap, spec.attack_map, spec.reward
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    resets = int((out['kind'] == 0).sum())
    print(f"{name}: {n_steps} steps, {resets} episodes, {n_draws} replayed draws ({n_ammo_draws} ammo-filter, "
          f"{n_repeat_acc} repeated-pair accuracy, {n_multi_subset} 2nd+ subset), "
          f"{int(out['done'].astype(bool).sum())} agent-steps -> {os.path.getsize(path) / 1024:.0f} KiB   oracle == reference")


def record_los(R=16):
    """create_grid_and_mask (utils.py:5-117) of the unmodified reference for every blocker offset at range R."""
    scenarios.reference_api()
    from abmarl.sim.gridworld.utils import create_grid_and_mask
    from abmarl.sim.gridworld.grid import Grid
    from abmarl.sim.gridworld.agent import GridObservingAgent, GridWorldAgent
    from oracle.oracle import los_mask
    n = 2 * R + 1
    out = np.zeros((n, n, n * n), dtype=np.uint8)
    for rd in range(-R, R + 1):
        for cd in range(-R, R + 1):
            grid = Grid(n, n, overlapping={1: {2}})
            viewer = GridObservingAgent(id='o', encoding=1, view_range=R, initial_position=np.array([R, R]))
            blocker = GridWorldAgent(id='b', encoding=2, blocking=True, initial_position=np.array([R + rd, R + cd]))
            agents = {'o': viewer, 'b': blocker}
            grid.reset()
            for a in agents.values():
                a.active = True
                assert grid.place(a, a.initial_position)
            _, mask = create_grid_and_mask(viewer, grid, R, agents)
            out[rd + R, cd + R] = mask.astype(np.uint8).ravel()
            if not np.array_equal(los_mask(R, rd, cd).ravel(), out[rd + R, cd + R]):
                raise SystemExit(f"oracle LOS mask != reference at blocker offset {(rd, cd)}")
    path = os.path.join(OUT, f'los_r{R}.npz')
    np.savez_compressed(path, range=np.int64(R), packed=np.packbits(out, axis=-1))
    print(f"los_r{R}: {n * n} blocker offsets -> {os.path.getsize(path) / 1024:.0f} KiB   oracle == reference")


if __name__ == '__main__':
    if sys.argv[1:] == ['los']:
        record_los()
        sys.exit(0)
    names = sys.argv[1:] or list(scenarios.SCENARIOS)
    for n in names:
        b, m, steps = scenarios.SCENARIOS[n]
        record(n, b, m, steps)
