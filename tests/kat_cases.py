"""Known-answer cases restated from the reference's own unit tests (tests/sim/gridworld/test_{actor,observer,
done,wrapper,grid}.py of gillette7/Abmarl 0.2.7).  Each case builds the flat spec by hand (the reference tests
drive bare components, not a simulation class), runs it on a backend -- the CPU oracle or the CUDA engine --
and checks the values the reference test asserts.  The team-battle program is the harness: it calls the
attack actor, then the move actor, for every acting agent in dict order, and its rewards expose each actor's
return value (-0.1 on a failed move or an attack without victims, +1 per kill).
"""
import math

import numpy as np

from abmarl_b200 import _capi as K
from abmarl_b200.spec import CompiledSpec

OBS, MOV, ATT, HEA, ORI, LRN, BLK, AMM = (K.AG_OBSERVING, K.AG_MOVING, K.AG_ATTACKING, K.AG_HEALTH, K.AG_ORIENT,
                                          K.AG_LEARNER, K.AG_BLOCKING, K.AG_AMMO)


def make_spec(rows, cols, agents, overlapping=None, attack_mapping=None, program=K.PROG_TEAM_BATTLE,
              move_actor=K.MOVE_BOX, attack_actor=K.ATTACK_NONE, observer=K.OBS_POSITION_CENTERED, observe_self=True,
              done_mask=K.DONE_ACTIVE, manager=K.MANAGER_ALL_STEP, ravel=False, stacked=False, n_envs=1, seed=24,
              position_observer=False):
    """agents: list of dicts(enc, pos=(r, c) | None, klass, view=0, move=0, att_range=0, strength=0, accuracy=1,
    health=None | float, orient=0, simatt=1, ammo=0)."""
    sp = CompiledSpec()
    sp.rows, sp.cols, sp.n_agents, sp.n_envs, sp.seed = rows, cols, len(agents), n_envs, seed
    sp.program, sp.move_actor, sp.attack_actor, sp.observer = program, move_actor, attack_actor, observer
    sp.observe_self, sp.done_mask, sp.manager, sp.ravel_actions = int(observe_self), done_mask, manager, int(ravel)
    sp.stacked_attacks = int(stacked)
    sp.position_observer = int(position_observer)
    for name, dt in CompiledSpec.TABLES:
        setattr(sp, name, np.zeros(len(agents), dtype=dt))
    sp.init_row[:] = -1
    sp.init_col[:] = -1
    sp.init_health[:] = math.nan
    sp.target[:] = -1
    for i, a in enumerate(agents):
        if a.get('target') is not None:
            sp.target[i] = a['target']
    sp.agent_ids = [f'agent{i}' for i in range(len(agents))]
    for i, a in enumerate(agents):
        sp.encoding[i], sp.klass[i] = a['enc'], a['klass']
        if a.get('pos') is not None:
            sp.init_row[i], sp.init_col[i] = a['pos']
        sp.view_range[i], sp.move_range[i] = a.get('view', 0), a.get('move', 0)
        sp.attack_range[i], sp.attack_strength[i] = a.get('att_range', 0), a.get('strength', 0)
        sp.attack_accuracy[i], sp.simultaneous_attacks[i] = a.get('accuracy', 1), a.get('simatt', 1) if a['klass'] & ATT else 0
        sp.initial_ammo[i] = a.get('ammo', 0)
        if a.get('health') is not None:
            sp.init_health[i] = a['health']
        sp.init_orient[i] = a.get('orient', 0)

    def rows_of(mapping):
        out = np.zeros(K.BGW_MAX_ENCODING + 1, dtype=np.uint64)
        for e, others in (mapping or {}).items():
            for o in others:
                out[e] |= np.uint64(1) << np.uint64(o)
        return out
    sym = {k: set(v) for k, v in (overlapping or {}).items()}
    for k, v in (overlapping or {}).items():                       # Grid symmetrises the map (grid.py:64-68)
        for o in v:
            sym.setdefault(o, set()).add(k)
    sp.overlap, sp.attack_map = rows_of(sym), rows_of(attack_mapping)
    sp.reward[K.RW_ATTACK_FAIL], sp.reward[K.RW_KILL], sp.reward[K.RW_DIE] = -0.1, 1.0, -1.0
    sp.reward[K.RW_MOVE_FAIL], sp.reward[K.RW_ENTROPY] = -0.1, -0.01
    return sp


class Backend:
    """Uniform numpy view of the oracle (OracleEnv) or the CUDA engine (BatchedGridWorld)."""

    def __init__(self, spec, kind, static_mask=True):
        self.kind = kind
        if kind == 'oracle':
            from oracle.oracle import OracleEnv
            self.env = OracleEnv(spec)
        else:
            import os
            from abmarl_b200.engine import BatchedGridWorld
            if not static_mask:              # the case flips `active` of a wall from outside, which no sim can do:
                os.environ['BGW_NO_STATIC_MASK'] = '1'   # trace every blocker instead of using the static-wall table
            try:
                self.env = BatchedGridWorld(spec, device='cuda:0')
            finally:
                os.environ.pop('BGW_NO_STATIC_MASK', None)
        self.L, self.spec = self.env.L, spec

    def reset(self):
        self.env.reset()

    def step(self, actions):
        act = np.zeros((1, self.L, self.env.dims.action_stride), dtype=np.int8)
        for l, a in enumerate(actions):
            a = (a,) if np.isscalar(a) else a
            for j, v in enumerate(a):
                act[0, l, j] = v
        if self.kind == 'oracle':
            self.env.step(act)
        else:
            import torch
            self.env.step(torch.from_numpy(act).cuda())

    def _np(self, x):
        return x if isinstance(x, np.ndarray) else x.cpu().numpy()

    def state(self):
        return self.env.state if self.kind == 'oracle' else self.env.state_numpy()

    def positions(self):
        cell = self.state()['cell'][0]
        return [(int(c) // self.spec.cols, int(c) % self.spec.cols) for c in cell]

    def obs(self, learner, n=None):
        """Observation of one learner, cropped to its own (2R+1)^2 window when smaller than the row."""
        flat = self._np(self.env.obs)[0, learner]
        d = self.env.dims
        if self.spec.observer == K.OBS_ABSOLUTE:
            return flat[:d.obs_h * d.obs_w].reshape(d.obs_h, d.obs_w).astype(int)
        n = d.obs_h if n is None else n
        c = d.obs_c
        out = flat[:n * n * c].astype(int)
        return out.reshape(n, n) if c == 1 else out.reshape(n, n, c)

    def position(self, learner):
        """The AbsolutePositionObserver's entry of one learner's observation: (row, col)."""
        off = self.env.dims.position_offset
        raw = np.ascontiguousarray(self._np(self.env.obs)[0, learner, off:off + 4])
        return tuple(int(v) for v in raw.view(np.int16))

    def rewards(self):
        return self._np(self.env.reward)[0].astype(np.float64)

    def flags(self):
        return self.state()['flags'][0]

    def health(self):
        return self.state()['health'][0]

    def ammo(self):
        return self.state()['ammo'][0]

    def set_flags(self, agent, clear=0):
        st = {k: (None if v is None else np.array(v)) for k, v in self.state().items()}
        st['flags'][0, agent] &= ~np.uint8(clear)
        if self.kind == 'oracle':
            self.env.state['flags'][:] = st['flags']
        else:
            self.env.load_state(st)

    def observe_now(self):
        """Refresh every learner's observation for the current state without stepping: sim.get_obs(agent_id) for all of
        them (oracle: bgwo_observe; engine: bgw_observe)."""
        if self.kind == 'oracle':
            self.env.obs[0] = self.env.observe(0)
        else:
            self.env.observe()


# ---------------------------------------------------------------------------------------------------
# test_actor.py
# ---------------------------------------------------------------------------------------------------
def _movers(specs):
    return [dict(enc=e, pos=p, klass=LRN | OBS | MOV, move=m, view=1) for e, p, m in specs]


def case_move_actor(kind):                                   # test_actor.py:22-78
    be = Backend(make_spec(5, 6, _movers([(1, (3, 4), 1), (2, (2, 2), 2), (1, (0, 1), 1), (3, (3, 1), 3)])), kind)
    be.reset()
    be.step([(1, 1), (-1, 0), (0, 1), (-1, 1)])
    assert be.positions() == [(4, 5), (1, 2), (0, 2), (2, 2)]
    be.step([(1, 1), (0, 0), (-1, 1), (-1, 0)])
    assert be.positions() == [(4, 5), (1, 2), (0, 2), (2, 2)]
    np.testing.assert_allclose(be.rewards(), [-0.11, -0.01, -0.11, -0.11], atol=1e-6)   # off grid / stay / off grid / occupied


def case_move_actor_overlap(kind):                           # test_actor.py:81-131
    be = Backend(make_spec(5, 6, _movers([(1, (4, 4), 1), (2, (2, 2), 2), (1, (2, 4), 1), (3, (3, 2), 3)]),
                           overlapping={1: {1}, 2: {3}, 3: {2}}), kind)
    be.reset()
    be.step([(-1, 0), (0, 0), (1, 0), (-1, 0)])
    assert be.positions() == [(3, 4), (2, 2), (3, 4), (2, 2)]
    be.step([(-1, 0), (0, 2), (0, -1), (1, 1)])
    assert be.positions() == [(2, 4), (2, 2), (3, 3), (2, 2)]


def case_cross_move_actor(kind):                             # test_actor.py:134-195
    be = Backend(make_spec(5, 6, _movers([(1, (3, 5), 1), (2, (2, 2), 2), (1, (0, 1), 1), (3, (2, 3), 3)]),
                           move_actor=K.MOVE_CROSS), kind)
    be.reset()
    be.step([2, 4, 3, 1])
    assert be.positions() == [(4, 5), (1, 2), (0, 2), (2, 2)]
    be.step([3, 0, 4, 4])
    assert be.positions() == [(4, 5), (1, 2), (0, 2), (2, 2)]


def case_cross_move_actor_overlap(kind):                     # test_actor.py:196-243
    be = Backend(make_spec(5, 6, _movers([(1, (4, 4), 1), (2, (2, 2), 2), (1, (2, 4), 1), (3, (3, 2), 3)]),
                           overlapping={1: {1}, 2: {3}, 3: {2}}, move_actor=K.MOVE_CROSS), kind)
    be.reset()
    be.step([4, 3, 2, 4])
    assert be.positions() == [(3, 4), (2, 3), (3, 4), (2, 2)]
    be.step([4, 0, 1, 3])
    assert be.positions() == [(2, 4), (2, 3), (3, 3), (2, 3)]


def case_drift_move_actor(kind):                             # test_actor.py:246-451
    coords, orient = [(0, 2), (2, 0), (2, 4), (4, 4)], [2, 3, 1, 2]
    agents = [dict(enc=o + 1, pos=coords[o], klass=LRN | OBS | MOV | ORI, move=1, view=1, orient=orient[o]) for o in range(4)]
    agents.append(dict(enc=2, pos=(2, 2), klass=0))            # wall_agent
    be = Backend(make_spec(5, 5, agents, overlapping={2: {1}, 1: {1}}, move_actor=K.MOVE_DRIFT), kind)
    be.reset()
    script = [
        ([0, 0, 0, 0], [1, 1, 1, 0], [(1, 2), (2, 1), (2, 3), (4, 4)], [2, 3, 1, 2]),
        ([0, 0, 0, 3], [1, 0, 0, 0], [(2, 2), (2, 1), (2, 3), (4, 4)], [2, 3, 1, 2]),
        ([0, 1, 4, 2], [1, 1, 1, 0], [(3, 2), (2, 0), (1, 3), (4, 4)], [2, 1, 4, 2]),
        ([2, 2, 4, 4], [1, 1, 1, 1], [(4, 2), (3, 0), (0, 3), (3, 4)], [2, 2, 4, 4]),
        ([0, 1, 0, 3], [0, 1, 0, 1], [(4, 2), (4, 0), (0, 3), (2, 4)], [2, 2, 4, 4]),
    ]
    for actions, ok, pos, ori in script:
        be.step(actions)
        assert be.positions()[:4] == pos
        assert [(int(f) >> K.ST_ORIENT_SHIFT) & 7 for f in be.flags()[:4]] == ori
        np.testing.assert_allclose(be.rewards(), [-0.01 if r else -0.11 for r in ok], atol=1e-6)   # process_action's return


def case_binary_attack_actor(kind):                          # test_actor.py:454-500
    agents = [dict(enc=1, pos=(4, 4), klass=HEA),
              dict(enc=1, pos=(2, 2), klass=LRN | OBS | ATT, att_range=2, strength=1, accuracy=1, view=1),
              dict(enc=2, pos=(2, 3), klass=HEA), dict(enc=1, pos=(3, 2), klass=HEA)]
    be = Backend(make_spec(5, 6, agents, attack_mapping={1: {1}}, attack_actor=K.ATTACK_BINARY), kind)
    be.reset()
    be.step([(0, 0, 1)])
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)        # one victim, killed (health <= 1, strength 1)
    be.step([(0, 0, 1)])
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)
    fl, h = be.flags(), be.health()
    for dead in (0, 3):
        assert not fl[dead] & K.ST_ACTIVE and not fl[dead] & K.ST_IN_GRID and h[dead] <= 0     # inactive, removed from the grid
    be.step([(0, 0, 1)])
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.1 - 0.01], atol=1e-6)     # attack_status True, no attacked agents
    fl, h = be.flags(), be.health()
    assert fl[2] & K.ST_ACTIVE and fl[2] & K.ST_IN_GRID and h[2] > 0             # encoding 2 is not in the mapping
    st = be.state()
    np.testing.assert_allclose(st['reward_acc'][0], [-1, 0, 0, -1])              # the victims' -1 is never read (non-learners)


def case_stacked_attack_exact_health(kind):                  # test_actor.py:701-705 (the n == 1 part)
    agents = [dict(enc=1, pos=(2, 2), klass=LRN | OBS | ATT, att_range=2, strength=0.5, accuracy=1, view=1),
              dict(enc=1, pos=(4, 4), klass=HEA, health=1), dict(enc=2, pos=(2, 3), klass=HEA, health=1),
              dict(enc=1, pos=(3, 2), klass=HEA, health=1)]
    be = Backend(make_spec(5, 6, agents, attack_mapping={1: {2}}, attack_actor=K.ATTACK_BINARY, stacked=True), kind)
    be.reset()
    be.step([(0, 0, 1)])
    assert be.health()[2] == 0.5 and be.flags()[2] & K.ST_ACTIVE
    be.step([(0, 0, 1)])
    assert be.health()[2] == 0.0 and not be.flags()[2] & K.ST_IN_GRID


def _sel(*cells, n=5):
    """(move 0, 0) + an n x n SelectiveAttackActor action with one attack on each listed window cell"""
    m = np.zeros((n, n), dtype=int)
    for r, c in cells:
        m[r, c] += 1
    return (0, 0) + tuple(m.ravel())


def _selective_agents(**attacker):
    return [dict(enc=1, pos=(4, 4), klass=HEA),
            dict(enc=1, pos=(2, 2), klass=LRN | OBS | ATT | attacker.pop('klass', 0), att_range=2, strength=1, accuracy=1, view=1, **attacker),
            dict(enc=2, pos=(2, 3), klass=HEA), dict(enc=1, pos=(3, 2), klass=HEA)]


def case_selective_attack_actor(kind):                       # test_actor.py:724-882
    be = Backend(make_spec(5, 6, _selective_agents(), attack_mapping={1: {1}}, attack_actor=K.ATTACK_SELECTIVE), kind)
    everywhere = [(r, c) for r in range(5) for c in range(5)]
    alive = lambda: [bool(f & K.ST_ACTIVE) and bool(f & K.ST_IN_GRID) for f in be.flags()]
    be.reset()
    be.step([_sel((4, 4))])                                   # attacking agent0
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)
    assert alive() == [False, True, True, True] and be.health()[0] <= 0
    be.step([_sel((3, 2))])                                   # attacking agent3
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)
    assert alive() == [False, True, True, False] and be.health()[3] <= 0
    be.reset()
    be.step([_sel((3, 2), (4, 4))])                           # both at once
    np.testing.assert_allclose(be.rewards(), [2 - 0.1 - 0.01], atol=1e-6)
    assert alive() == [False, True, True, False]
    be.reset()
    be.step([_sel(*[rc for rc in everywhere if rc not in ((3, 2), (4, 4))])])   # everywhere but agent0 / agent3
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.1 - 0.01], atol=1e-6)   # attempted, nobody attacked
    assert alive() == [True, True, True, True]
    be.step([_sel(*everywhere)])                              # everywhere: agent2 (encoding 2) is not attackable
    np.testing.assert_allclose(be.rewards(), [2 - 0.1 - 0.01], atol=1e-6)
    assert alive() == [False, True, True, False]
    be.reset()
    be.step([_sel()])                                         # nowhere: no attack attempted
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.01], atol=1e-6)
    assert alive() == [True, True, True, True]


def case_selective_attack_actor_ammo(kind):                  # test_actor.py:885-987
    spec = make_spec(5, 6, _selective_agents(klass=AMM, ammo=3), attack_mapping={1: {1}}, attack_actor=K.ATTACK_SELECTIVE)
    be = Backend(spec, kind)
    be.reset()
    assert be.ammo()[1] == 3
    be.step([_sel((4, 4))])
    assert be.ammo()[1] == 2 and not be.flags()[0] & K.ST_ACTIVE
    be.step([_sel((3, 2))])
    assert be.ammo()[1] == 1 and not be.flags()[3] & K.ST_ACTIVE
    # "attacking both agent0 and agent3" with one round left, then everywhere with none (:957-986)
    be = Backend(make_spec(5, 6, _selective_agents(klass=AMM, ammo=1), attack_mapping={1: {1}}, attack_actor=K.ATTACK_SELECTIVE), kind)
    be.reset()
    be.step([_sel((3, 2), (4, 4))])
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)      # len(attacked_agents) == 1
    assert be.ammo()[1] == 0
    assert sorted(bool(be.flags()[a] & K.ST_ACTIVE) for a in (0, 3)) == [False, True]
    be.step([_sel(*[(r, c) for r in range(5) for c in range(5)])])
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.1 - 0.01], atol=1e-6)   # attack_status, not attacked_agents
    assert be.ammo()[1] == 0


def _enc_attackers(strength):
    return [dict(enc=1, pos=(0, 0), klass=HEA), dict(enc=2, pos=(0, 1), klass=HEA), dict(enc=2, pos=(1, 0), klass=HEA),
            dict(enc=3, pos=(1, 1), klass=LRN | OBS | ATT | AMM, att_range=1, strength=strength, accuracy=1, view=1, ammo=100),
            dict(enc=1, pos=(1, 1), klass=HEA)]


def case_encoding_based_attack_actor(kind):                  # test_actor.py:1175-1247
    kw = dict(overlapping={1: {3}, 3: {1}}, attack_mapping={3: {1, 2}}, attack_actor=K.ATTACK_ENCODING)
    be = Backend(make_spec(2, 2, _enc_attackers(0), **kw), kind)   # "should still be active because attacking agent is weak"
    be.reset()
    be.step([(0, 0, 0, 1)])                                   # {1: 0, 2: 1}
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.01], atol=1e-6)          # one agent attacked, none killed
    assert be.ammo()[3] == 99 and all(f & K.ST_ACTIVE for f in be.flags())
    be = Backend(make_spec(2, 2, _enc_attackers(1), **kw), kind)
    be.reset()
    dead = lambda: [a for a in range(5) if not be.flags()[a] & K.ST_ACTIVE]
    be.step([(0, 0, 1, 0)])                                   # {1: 1, 2: 0}: one of the two encoding-1 agents
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)
    assert len(dead()) == 1 and dead()[0] in (0, 4)
    be.step([(0, 0, 1, 1)])                                   # the other encoding 1 and one encoding 2
    np.testing.assert_allclose(be.rewards(), [2 - 0.1 - 0.01], atol=1e-6)
    assert len(dead()) == 3 and {0, 4} <= set(dead())
    be.step([(0, 0, 1, 1)])                                   # only an encoding 2 is left
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)
    assert dead() == [0, 1, 2, 4]
    be.step([(0, 0, 1, 1)])                                   # attack_status True, len(attacked_agents) == 0
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.1 - 0.01], atol=1e-6)
    assert be.ammo()[3] == 96


def case_restricted_selective_attack_actor(kind):            # test_actor.py:1505-1573 (+ stacked :1576-1645)
    agents = [dict(enc=1, pos=(0, 0), klass=HEA), dict(enc=2, pos=(0, 1), klass=HEA), dict(enc=2, pos=(1, 0), klass=HEA),
              dict(enc=3, pos=(1, 1), klass=LRN | OBS | ATT | AMM, att_range=1, strength=0, accuracy=1, view=1, simatt=2, ammo=100),
              dict(enc=1, pos=(0, 0), klass=HEA)]
    kw = dict(overlapping={1: {1}}, attack_mapping={3: {1, 2}}, attack_actor=K.ATTACK_RESTRICTED)
    be = Backend(make_spec(2, 2, agents, **kw), kind)
    be.reset()
    be.step([(0, 0, 0, 0)])                                   # [0, 0]: not attack_status
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.01], atol=1e-6)
    assert be.ammo()[3] == 100
    be.step([(0, 0, 1, 1)])                                   # twice the cell (0, 0): two different encoding-1 agents
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.01], atol=1e-6)
    assert be.ammo()[3] == 98
    be.step([(0, 0, 2, 2)])                                   # twice the cell (1, 0): its one agent only once
    assert be.ammo()[3] == 97
    be.step([(0, 0, 1, 4)])                                   # cells (0, 0) and (0, 1)
    assert be.ammo()[3] == 95
    be.step([(0, 0, 5, 9)])                                   # own cell and an off-grid cell: attempted, nobody attacked
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.1 - 0.01], atol=1e-6)
    assert be.ammo()[3] == 95
    be = Backend(make_spec(2, 2, agents, stacked=True, **kw), kind)   # stacked_attacks: the same agent may be named twice
    be.reset()
    be.step([(0, 0, 2, 2)])
    assert be.ammo()[3] == 98


def _sim3_agents(strength):
    """the 2x2 board of test_actor.py:988-1012 / 1088-1113: the attacker (ammo 100 so that len(attacked_agents) shows)
    shares its cell with three attackable agents, a fourth sits on (0, 0)"""
    return [dict(enc=3, pos=(1, 1), klass=LRN | OBS | ATT | AMM, att_range=1, strength=strength, accuracy=1, view=1, simatt=3, ammo=100),
            dict(enc=1, pos=(0, 0), klass=HEA, health=1), dict(enc=2, pos=(1, 1), klass=HEA, health=1),
            dict(enc=2, pos=(1, 1), klass=HEA, health=1), dict(enc=1, pos=(1, 1), klass=HEA, health=1)]


_ALL3 = {1: {1, 2, 3}, 2: {1, 2, 3}, 3: {1, 2, 3}}


def case_selective_attack_actor_simultaneous_attacks(kind):  # test_actor.py:988-1085
    kw = dict(overlapping=_ALL3, attack_mapping={3: {1, 2}}, attack_actor=K.ATTACK_SELECTIVE)
    be = Backend(make_spec(2, 2, _sim3_agents(0), **kw), kind)
    be.reset()
    used = []
    for k in (0, 1, 2, 3):                                    # k attacks on the cells (0, 0) and (1, 1)
        before = int(be.ammo()[0])
        be.step([(0, 0) + tuple(np.array([[k, 0, 0], [0, k, 0], [0, 0, 0]]).ravel())])
        used.append(before - int(be.ammo()[0]))
        np.testing.assert_allclose(be.rewards(), [-0.1 - 0.01], atol=1e-6)          # attempted or not, nobody dies, no failure
    assert used == [0, 2, 3, 4]                               # 1 + min(k, 3 agents on the own cell)
    be = Backend(make_spec(2, 2, _sim3_agents(1), **kw), kind)
    be.reset()
    be.step([(0, 0) + tuple(np.array([[3, 0, 0], [0, 3, 0], [0, 0, 0]]).ravel())])
    np.testing.assert_allclose(be.rewards(), [4 - 0.1 - 0.01], atol=1e-6)
    assert not any(be.flags()[a] & K.ST_ACTIVE for a in (1, 2, 3, 4)) and be.flags()[0] & K.ST_ACTIVE


def case_selective_attack_actor_stacked_attack(kind):        # test_actor.py:1088-1172
    kw = dict(overlapping=_ALL3, attack_mapping={3: {1, 2}}, attack_actor=K.ATTACK_SELECTIVE)
    be = Backend(make_spec(2, 2, _sim3_agents(1), **kw), kind)
    be.reset()
    be.step([(0, 0) + tuple(np.array([[0, 0, 0], [0, 2, 0], [0, 0, 0]]).ravel())])   # two different agents of the own cell
    np.testing.assert_allclose(be.rewards(), [2 - 0.1 - 0.01], atol=1e-6)
    assert sum(1 for a in (2, 3, 4) if not be.flags()[a] & K.ST_ACTIVE) == 2 and be.flags()[1] & K.ST_ACTIVE
    be = Backend(make_spec(2, 2, _sim3_agents(0.5), stacked=True, **kw), kind)
    be.reset()
    be.step([(0, 0) + tuple(np.array([[1, 0, 0], [0, 3, 0], [0, 0, 0]]).ravel())])   # stacked: 1 + 3 draws with replacement
    assert be.ammo()[0] == 96
    h = np.array(be.health())
    assert h[1] == 0.5 and sorted(2 * (1 - h[a]) for a in (2, 3, 4)) in ([0, 0, 2], [0, 1, 2], [1, 1, 1])   # three hits of 0.5 in all
    be.step([(0, 0) + tuple(np.array([[3, 3, 3], [3, 3, 3], [3, 3, 3]]).ravel())])
    assert be.ammo()[0] == 96 - 3 - 3 * (sum(1 for a in (2, 3, 4) if h[a] > 0) > 0)   # 3 on (0, 0); 3 on the own cell while anyone is left


def case_encoding_based_attack_actor_simultaneous_attacks(kind):   # test_actor.py:1324-1436
    def agents(strength):
        return [dict(enc=1, pos=(0, 0), klass=HEA), dict(enc=2, pos=(0, 1), klass=HEA), dict(enc=2, pos=(1, 0), klass=HEA),
                dict(enc=3, pos=(1, 1), klass=LRN | OBS | ATT | AMM, att_range=1, strength=strength, accuracy=1, view=1, simatt=2, ammo=100),
                dict(enc=1, pos=(1, 1), klass=HEA)]
    kw = dict(overlapping={1: {3}, 3: {1}}, attack_mapping={3: {1, 2}}, attack_actor=K.ATTACK_ENCODING)
    be = Backend(make_spec(2, 2, agents(0), **kw), kind)
    be.reset()
    used = []
    for e1, e2 in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 1), (1, 2), (2, 2)):
        before = int(be.ammo()[3])
        be.step([(0, 0, e1, e2)])
        used.append(before - int(be.ammo()[3]))
    assert used == [0, 1, 1, 2, 3, 3, 4]
    be = Backend(make_spec(2, 2, agents(1), **kw), kind)
    be.reset()
    be.step([(0, 0, 2, 0)])                                   # both encoding-1 agents
    np.testing.assert_allclose(be.rewards(), [2 - 0.1 - 0.01], atol=1e-6)
    assert [bool(f & K.ST_ACTIVE) for f in be.flags()] == [False, True, True, True, False]
    be.step([(0, 0, 2, 2)])                                   # only the two encoding-2 agents are left
    np.testing.assert_allclose(be.rewards(), [2 - 0.1 - 0.01], atol=1e-6)
    be.step([(0, 0, 1, 1)])                                   # attack_status True, nobody left
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.1 - 0.01], atol=1e-6)


def case_restricted_selective_attack_actor_ammo(kind):       # test_actor.py:1648-1709: ammo bounds the attack
    agents = [dict(enc=1, pos=(0, 0), klass=HEA), dict(enc=2, pos=(0, 1), klass=HEA), dict(enc=2, pos=(1, 0), klass=HEA),
              dict(enc=3, pos=(1, 1), klass=LRN | OBS | ATT | AMM, att_range=1, strength=1, accuracy=1, view=1, simatt=2, ammo=3),
              dict(enc=1, pos=(0, 0), klass=HEA)]
    be = Backend(make_spec(2, 2, agents, overlapping={1: {1}}, attack_mapping={3: {1, 2}}, attack_actor=K.ATTACK_RESTRICTED), kind)
    be.reset()
    be.step([(0, 0, 1, 1)])                                   # both agents of the cell (0, 0)
    assert be.ammo()[3] == 1 and not be.flags()[0] & K.ST_ACTIVE and not be.flags()[4] & K.ST_ACTIVE
    be.step([(0, 0, 2, 4)])                                   # two more named, one round left: the ammo filter keeps one
    np.testing.assert_allclose(be.rewards(), [1 - 0.1 - 0.01], atol=1e-6)
    assert be.ammo()[3] == 0 and sorted(bool(be.flags()[a] & K.ST_ACTIVE) for a in (1, 2)) == [False, True]
    be.step([(0, 0, 2, 4)])                                   # out of ammo: attempted, nobody attacked
    np.testing.assert_allclose(be.rewards(), [-0.1 - 0.1 - 0.01], atol=1e-6)
    assert be.ammo()[3] == 0


# ---------------------------------------------------------------------------------------------------
# test_observer.py
# ---------------------------------------------------------------------------------------------------
def _observer_agents(blocking, with_agent6):
    b = BLK if blocking else 0
    agents = [dict(enc=1, pos=(2, 2), klass=LRN | OBS, view=2), dict(enc=2, pos=(0, 0), klass=LRN | OBS, view=1),
              dict(enc=3, pos=(4, 4), klass=LRN | OBS, view=4), dict(enc=5, pos=(3, 3), klass=b),
              dict(enc=4, pos=(1, 1), klass=b), dict(enc=6, pos=(2, 1), klass=b)]
    if with_agent6:
        agents.append(dict(enc=6, pos=(2, 2), klass=0))
    return agents


def case_absolute_encoding_observer(kind):                   # test_observer.py:42-103
    be = Backend(make_spec(5, 5, _observer_agents(False, True), overlapping={1: {6}, 6: {1}}, observer=K.OBS_ABSOLUTE), kind)
    be.reset()
    np.testing.assert_array_equal(be.obs(0), [[2, 0, 0, 0, 0], [0, 4, 0, 0, 0], [0, 6, -1, 0, 0], [0, 0, 0, 5, 0], [0, 0, 0, 0, 3]])
    np.testing.assert_array_equal(be.obs(1), [[-1, 0, -2, -2, -2], [0, 4, -2, -2, -2]] + [[-2] * 5] * 3)
    got = be.obs(2)
    assert got[2, 2] in (1, 6)             # np.random.choice over [1, 6]: the reference pins it with MT19937 seed 24
    got[2, 2] = 1
    np.testing.assert_array_equal(got, [[2, 0, 0, 0, 0], [0, 4, 0, 0, 0], [0, 6, 1, 0, 0], [0, 0, 0, 5, 0], [0, 0, 0, 0, -1]])


def case_absolute_encoding_observer_blocking(kind):          # test_observer.py:106-191
    be = Backend(make_spec(5, 5, _observer_agents(True, True), overlapping={1: {6}, 6: {1}}, observer=K.OBS_ABSOLUTE), kind,
                 static_mask=False)
    be.reset()
    np.testing.assert_array_equal(be.obs(0), [[-2, -2, 0, 0, 0], [-2, 4, 0, 0, 0], [-2, 6, -1, 0, 0], [-2, 0, 0, 5, -2], [0, 0, 0, -2, -2]])
    np.testing.assert_array_equal(be.obs(1), [[-1, 0, -2, -2, -2], [0, 4, -2, -2, -2]] + [[-2] * 5] * 3)
    np.testing.assert_array_equal(be.obs(2), [[-2, -2, -2, 0, 0], [-2, -2, -2, 0, 0], [-2, -2, -2, -2, 0], [0, 0, -2, 5, 0], [0, 0, 0, 0, -1]])
    be.set_flags(3, clear=K.ST_ACTIVE)                         # agents['agent3'].active = False: it no longer blocks
    be.observe_now()
    np.testing.assert_array_equal(be.obs(0), [[-2, -2, 0, 0, 0], [-2, 4, 0, 0, 0], [-2, 6, -1, 0, 0], [-2, 0, 0, 5, 0], [0, 0, 0, 0, 3]])
    got = be.obs(2)
    assert got[2, 2] in (1, 6)
    got[2, 2] = 1
    np.testing.assert_array_equal(got, [[-2, -2, 0, 0, 0], [-2, 4, 0, 0, 0], [-2, 6, 1, 0, 0], [0, 0, 0, 5, 0], [0, 0, 0, 0, -1]])


_SINGLE_FAR = [[2, 0, 0, 0, 0], [0, 4, 0, 0, 0], [0, 6, 1, 0, 0], [0, 0, 0, 5, 0], [0, 0, 0, 0, 3]]


def _pad9(top):
    return [row + [-1] * 4 for row in top] + [[-1] * 9] * 4


def case_single_grid_observer(kind):                         # test_observer.py:194-277
    be = Backend(make_spec(5, 5, _observer_agents(False, False)), kind)
    be.reset()
    np.testing.assert_array_equal(be.obs(0, 5), _SINGLE_FAR)
    np.testing.assert_array_equal(be.obs(1, 3), [[-1, -1, -1], [-1, 2, 0], [-1, 0, 4]])
    np.testing.assert_array_equal(be.obs(2, 9), _pad9(_SINGLE_FAR))


def case_single_grid_observer_blocking(kind):                # test_observer.py:280-339
    be = Backend(make_spec(5, 5, _observer_agents(True, False)), kind)
    be.reset()
    np.testing.assert_array_equal(be.obs(0, 5), [[-2, -2, 0, 0, 0], [-2, 4, 0, 0, 0], [-2, 6, 1, 0, 0], [-2, 0, 0, 5, -2], [0, 0, 0, -2, -2]])
    np.testing.assert_array_equal(be.obs(1, 3), [[-1, -1, -1], [-1, 2, 0], [-1, 0, 4]])
    np.testing.assert_array_equal(be.obs(2, 9), _pad9([[-2, -2, -2, 0, 0], [-2, -2, -2, 0, 0], [-2, -2, -2, -2, 0],
                                                        [0, 0, -2, 5, 0], [0, 0, 0, 0, 3]]))


def case_multi_grid_observer(kind):                          # test_observer.py:342-470 (agent0's six channels)
    agents = [dict(enc=1, pos=(2, 2), klass=LRN | OBS, view=2), dict(enc=2, pos=(0, 0), klass=LRN | OBS, view=1),
              dict(enc=3, pos=(4, 4), klass=LRN | OBS, view=4), dict(enc=2, pos=(4, 4), klass=LRN | OBS | MOV, view=1, move=1),
              dict(enc=3, pos=(0, 0), klass=LRN | OBS | MOV, view=4, move=1), dict(enc=5, pos=(3, 3), klass=0),
              dict(enc=5, pos=(3, 3), klass=MOV, move=1), dict(enc=4, pos=(1, 1), klass=0), dict(enc=6, pos=(2, 1), klass=0)]
    be = Backend(make_spec(5, 5, agents, overlapping={2: {3}, 3: {2}, 5: {5}}, observer=K.OBS_STACKED), kind)
    be.reset()
    got = be.obs(0, 5)
    want = np.zeros((5, 5, 6), dtype=int)
    want[2, 2, 0] = 1
    want[0, 0, 1] = want[4, 4, 1] = 1
    want[0, 0, 2] = want[4, 4, 2] = 1
    want[1, 1, 3] = 1
    want[3, 3, 4] = 2
    want[2, 1, 5] = 1
    np.testing.assert_array_equal(got, want)


# ---------------------------------------------------------------------------------------------------
# test_wrapper.py (RavelActionWrapper) and test_done.py
# ---------------------------------------------------------------------------------------------------
def case_absolute_position_observer(kind):                   # test_observer.py:905-988 (+ the combined test :991-1070)
    """Agents observe their absolute position; with a grid observer next to it both entries are filled."""
    pos = [(0, 0), (5, 0), (0, 6), (5, 6), (0, 0), (5, 6)]
    agents = [dict(enc=e + 1, pos=p, klass=LRN | OBS, view=2) for e, p in enumerate(pos)]
    b = Backend(make_spec(6, 7, agents, overlapping={1: {5}, 4: {6}, 5: {1}, 6: {4}}, position_observer=True), kind)
    b.reset()
    for l, p in enumerate(pos):
        assert b.position(l) == p, (l, b.position(l), p)
    # the grid part of the row is still the position-centred window: agent0 at (0, 0) sees agent4 (encoding 5) on its own
    # cell or itself (encoding 1) -- one draw -- and the border
    o = b.obs(0)
    assert o.shape == (5, 5) and o[2, 2] in (1, 5) and (o[:2] == -1).all() and (o[:, :2] == -1).all()


def case_ravel_action_wrapper(kind):                         # test_wrapper.py:111-144: 7 -> [1, 0], 3 -> [-2, 1], 34 -> [1, 3]
    agents = _movers([(1, (2, 2), 1), (2, (4, 4), 2), (3, (4, 1), 3)])
    be = Backend(make_spec(8, 8, agents, ravel=True), kind)
    be.reset()
    be.step([7, 3, 34])
    assert be.positions() == [(3, 2), (2, 5), (5, 4)]


def case_active_done(kind):                                  # test_done.py: ActiveDone -- done iff inactive
    agents = [dict(enc=1, pos=(0, 0), klass=LRN | OBS | ATT | HEA, att_range=1, strength=1, accuracy=1, view=1, health=1),
              dict(enc=2, pos=(0, 1), klass=LRN | OBS | HEA, view=1, health=0.5),
              dict(enc=2, pos=(3, 3), klass=LRN | OBS | HEA, view=1, health=1)]
    be = Backend(make_spec(4, 4, agents, attack_mapping={1: {2}}, attack_actor=K.ATTACK_BINARY), kind)
    be.reset()
    be.step([(0, 0, 1), (0, 0, 0), (0, 0, 0)])
    done = be._np(be.env.done)[0]
    assert list(done & K.OUT_DONE) == [0, 1, 0] and all(done & K.OUT_VALID)
    assert not be._np(be.env.all_done)[0] & K.ENV_ALL_DONE       # two entities are still active


def case_target_agent_done(kind):                            # test_done.py:45-118 (TargetAgentDone) through the manager
    pos = [(0, 0), (0, 1), (1, 0), (1, 1)]
    agents = [dict(enc=1 + (i >= 2), pos=pos[i], klass=LRN | OBS | MOV, move=1, view=1, target=(i + 1) % 4) for i in range(4)]
    be = Backend(make_spec(2, 2, agents, overlapping={1: {1, 2}, 2: {1, 2}}, done_mask=K.DONE_TARGET_AGENT,
                           attack_actor=K.ATTACK_BINARY), kind)
    be.reset()
    script = [  # (mover, move) -> learners reported done this step, __all__
        ((0, (0, 1)), [0], False),       # agent0 steps onto agent1's cell
        ((1, (1, -1)), [1], False),      # agent1 onto agent2's cell (agent0 is done and no longer reported)
        ((3, (-1, 0)), [3], False),      # agent3 onto agent0's cell
        ((2, (-1, 1)), [2], True),       # agent2 onto agent3's cell: every learner has been reported done
    ]
    reported = set()
    for (mover, mv), newly_done, all_done in script:
        be.step([mv if l == mover else (0, 0) for l in range(4)])
        done = be._np(be.env.done)[0]
        for l in range(4):
            assert bool(done[l] & K.OUT_VALID) == (l not in reported)
            assert bool(done[l] & K.OUT_DONE) == (l in newly_done)
        reported |= set(newly_done)
        assert bool(be._np(be.env.all_done)[0] & K.ENV_ALL_DONE) == all_done


def case_target_destroyed_done(kind):                        # test_done.py:120-189 (TargetDestroyedDone) through attacks
    agents = [dict(enc=1, pos=(0, 0), klass=LRN | OBS | ATT, att_range=1, strength=1, accuracy=1, view=1, target=2),
              dict(enc=1, pos=(0, 3), klass=LRN | OBS | ATT, att_range=1, strength=1, accuracy=1, view=1, target=3),
              dict(enc=3, pos=(1, 0), klass=HEA, health=1), dict(enc=3, pos=(1, 3), klass=HEA, health=0.5)]
    be = Backend(make_spec(2, 4, agents, attack_mapping={1: {3}}, attack_actor=K.ATTACK_BINARY,
                           done_mask=K.DONE_TARGET_DESTROYED), kind)
    be.reset()
    be.step([(0, 0, 0), (0, 0, 1)])                            # agent1 destroys its target
    done = be._np(be.env.done)[0]
    assert [bool(d & K.OUT_DONE) for d in done] == [False, True]
    assert not be._np(be.env.all_done)[0] & K.ENV_ALL_DONE     # agent0's target is still active
    be.step([(0, 0, 1), (0, 0, 0)])
    done = be._np(be.env.done)[0]
    assert bool(done[0] & K.OUT_DONE) and not done[1] & K.OUT_VALID
    assert be._np(be.env.all_done)[0] & K.ENV_ALL_DONE         # every target destroyed (done.py:133-137)


CASES = [v for k, v in sorted(globals().items()) if k.startswith('case_')]
