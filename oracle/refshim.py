"""Run the UNMODIFIED reference as a live oracle with its random draws replayed from the keyed Philox
stream.  TEST INFRASTRUCTURE ONLY; used in the build container (where /root/reference exists) by
tests/golden/make_golden.py and by tests that skip when the reference is absent.

The reference draws from numpy's global legacy generator in call order (SURVEY.md section 8(a) "RNG draw
sites").  `PhiloxReplay` patches numpy.random.{uniform,choice,randint} for the duration of a `with`
block: each call inspects the CALLER's frame to find the draw site and the agents involved, and returns the
value the engine would draw for the key (seed, env, episode, step, site, slot, k) -- see
include/bgw_philox.h for the mapping.  Nothing in /root/reference is modified.
"""
import os
import random
import sys

import numpy as np

from abmarl_b200 import philox
from abmarl_b200 import _capi as K

REFERENCE_ROOT = os.environ.get('BGW_REFERENCE_ROOT', '/root/reference')
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_shims')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'abmarl'))


def load_reference():
    """Make `import abmarl` resolve to the reference (with the gym stand-in on the path)."""
    assert reference_available(), f"reference not found at {REFERENCE_ROOT}"
    for p in (SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import abmarl  # noqa: F401
    return abmarl


class PhiloxReplay:
    """Context manager replaying keyed Philox draws into the reference's numpy.random call sites."""

    def __init__(self, sim, seed, env=0):
        self.sim, self.seed, self.env = sim, int(seed), int(env)
        self.index = {agent_id: i for i, agent_id in enumerate(sim.agents)}
        self.episode, self.step = -1, 0
        self.maze_k, self.maze_episode = 0, None                  # position in the episode's maze draw sequence
        self.acc_seen = {}                  # (episode, step, attacker, candidate) -> evaluations so far (ACC draw key)
        self.log = []                       # (site, slot, k) of every replayed draw, for debugging

    # -- keyed draw -------------------------------------------------------------------------------
    def _x(self, site, slot, k=0):
        self.log.append((site, slot, k))
        return philox.draw(self.seed, self.env, self.episode & 0xFFFFFFFF, self.step, site, slot, k)

    # -- patched numpy.random functions ----------------------------------------------------------------
    def _uniform(self, low=0.0, high=1.0, size=None):
        f = sys._getframe(1)
        name, loc = f.f_code.co_name, f.f_locals
        assert size is None
        if name == '_basic_criteria':                        # actor.py:388
            att, cand = self.index[loc['attacking_agent'].id], self.index[loc['candidate'].id]
            key = (self.episode, self.step, att, cand)
            occ = self.acc_seen.get(key, 0)                  # only RestrictedSelective evaluates a pair twice
            if len(self.acc_seen) > 4096:
                self.acc_seen.clear()
            self.acc_seen[key] = occ + 1
            x = self._x(K.SITE_ACC, att, cand + 4096 * occ)
        elif name == 'reset' and 'agent' in loc:             # HealthState.reset state.py:641
            x = self._x(K.SITE_HEALTH, self.index[loc['agent'].id])
        else:
            raise RuntimeError(f"unexpected np.random.uniform call site: {name}")
        return low + (high - low) * philox.u01(x)

    def _choice(self, a, size=None, replace=True, p=None):
        f = sys._getframe(1)
        name, loc = f.f_code.co_name, f.f_locals
        assert p is None
        seq = list(a)
        n = len(seq)
        if name == '_place_variable_position_agent':         # state.py:159
            if n == 0:
                raise ValueError("'a' cannot be empty unless no samples are taken")
            x = self._x(K.SITE_PLACE, self.index[loc['var_agent_to_place'].id])
            return np.array([seq[philox.index(x, n)]])
        if name == '_subset_attackables':                    # actor.py:412
            outer = sys._getframe(2).f_locals                # _determine_attack(self, agent, attack)
            attacker = outer['agent']
            slot = self.index[attacker.id]
            actor = type(outer['self']).__name__
            if actor == 'EncodingBasedAttackActor':          # one call per encoding :575-581
                group = int(outer['encoding'])
            elif actor == 'SelectiveAttackActor':            # one call per attacked cell :711-726
                group = int(outer['r']) * (2 * attacker.attack_range + 1) + int(outer['c'])
            else:
                group = 0
            size = int(size)
            out = np.empty(size, dtype=object)
            if replace:
                for t in range(size):
                    out[t] = seq[philox.index(self._x(K.SITE_SUBSET, slot, (group << 8) | t), n)]
            else:                                            # partial Fisher-Yates, as the oracle/engine
                if size > n:
                    raise ValueError("Cannot take a larger sample than population when 'replace=False'")
                for t in range(size):
                    j = t + philox.index(self._x(K.SITE_SUBSET, slot, (group << 8) | t), n - t)
                    seq[t], seq[j] = seq[j], seq[t]
                    out[t] = seq[t]
            return out
        if name == '_determine_attack':                      # RestrictedSelectiveAttackActor actor.py:655-656
            assert size is None
            slot = self.index[loc['agent'].id]
            group = len(loc['attacked_agents'])
            return seq[philox.index(self._x(K.SITE_SUBSET, slot, group << 8), n)]
        if name == 'process_action':                         # ammo filter actor.py:346-350
            assert not replace
            slot = self.index[loc['attacking_agent'].id]
            size = int(size)
            out = np.empty(size, dtype=object)
            for t in range(size):
                j = t + philox.index(self._x(K.SITE_AMMO, slot, t), n - t)
                seq[t], seq[j] = seq[j], seq[t]
                out[t] = seq[t]
            return out
        if name == 'get_obs':                                # observer.py:131,234,246
            agent = loc['agent']
            if n == 1:
                return seq[0]
            R = agent.view_range
            r, c = int(agent.position[0]) - R + loc['r'], int(agent.position[1]) - R + loc['c']
            cell = r * self.sim.grid.cols + c
            return seq[philox.index(self._x(K.SITE_OBS, self.index[agent.id], cell), n)]
        raise RuntimeError(f"unexpected np.random.choice call site: {name}")

    def _randint(self, low, high=None, size=None, dtype=int):
        f = sys._getframe(1)
        name, loc = f.f_code.co_name, f.f_locals
        if name == 'reset' and 'agent' in loc:               # OrientationState.reset state.py:675
            x = self._x(K.SITE_ORIENT, self.index[loc['agent'].id])
            return low + philox.index(x, high - low)
        if name == 'step' and 'action_dict' in loc:          # PacmanSimSimple's random baddie move, pacman.py:240
            x = self._x(K.SITE_SCRIPT, self.index['baddie_159'])
            return low + philox.index(x, high - low)
        if name in ('_build_available_positions', 'generate_maze'):   # state.py:534, utils.py:193,198
            if self.maze_episode != self.episode:
                self.maze_episode, self.maze_k = self.episode, 0
            saved_step, self.step = self.step, 0              # maze draws are keyed by (episode, k) only

            def one(lo, hi):
                x = self._x(K.SITE_MAZE, 0, self.maze_k)
                self.maze_k += 1
                return int(lo) + philox.index(x, int(hi) - int(lo))
            try:
                if np.ndim(high) == 0:
                    return one(low, high)
                lows = np.broadcast_to(low, np.shape(high))
                return np.array([one(lo, hi) for lo, hi in zip(lows, high)])
            finally:
                self.step = saved_step
        raise RuntimeError(f"unexpected np.random.randint call site: {name}")

    def _shuffle(self, x):
        """random.shuffle (Python's generator) -> the keyed order of include/bgw_philox.h: the items, which name agents,
        sorted by (first Philox word of the agent's key, agent index).  In place, like random.shuffle."""
        f = sys._getframe(1)
        name, loc = f.f_code.co_name, f.f_locals
        if name == 'reset' and 'agents' in loc:              # PositionState.reset state.py:97-101: [(id, agent), ...]
            site, step = K.SITE_PLACE_ORDER, 0
        elif name == 'step' and 'action_list' in loc:        # AllStepManager.step all_step_manager.py:62-65: [(id, action), ...]
            site, step = K.SITE_ORDER, self.step
        else:
            raise RuntimeError(f"unexpected random.shuffle call site: {name}")
        saved, self.step = self.step, step
        try:
            x.sort(key=lambda item: (self._x(site, self.index[item[0]]), self.index[item[0]]))
        finally:
            self.step = saved

    def __enter__(self):
        self._saved = (np.random.uniform, np.random.choice, np.random.randint, random.shuffle)
        np.random.uniform, np.random.choice, np.random.randint = self._uniform, self._choice, self._randint
        random.shuffle = self._shuffle
        return self

    def __exit__(self, *exc):
        np.random.uniform, np.random.choice, np.random.randint, random.shuffle = self._saved
        return False


# ---------------------------------------------------------------------------------------------------
# state extraction: reference sim  ->  BgwState layout
# ---------------------------------------------------------------------------------------------------
def extract_state(sim, done_agents=()):
    """cell / next / flags / health arrays ([A]) of a live reference sim, in the layout of include/bgw.h."""
    ids = list(sim.agents)
    index = {a: i for i, a in enumerate(ids)}
    A = len(ids)
    cols = sim.grid.cols
    cell = np.full(A, K.BGW_NONE, dtype=np.uint16)
    nxt = np.full(A, K.BGW_NONE, dtype=np.uint16)
    flags = np.zeros(A, dtype=np.uint8)
    health = np.zeros(A, dtype=np.float64)
    ammo = np.zeros(A, dtype=np.int32)
    for i, agent in enumerate(sim.agents.values()):
        if hasattr(agent, 'initial_ammo'):
            ammo[i] = int(getattr(agent, '_ammo', 0))
        pos = getattr(agent, 'position', None)
        if pos is not None:
            cell[i] = int(pos[0]) * cols + int(pos[1])
        if agent.active:
            flags[i] |= K.ST_ACTIVE
        if hasattr(agent, 'initial_health'):
            health[i] = float(getattr(agent, '_health', 0.0))
        if hasattr(agent, 'initial_orientation') and hasattr(agent, '_orientation'):
            flags[i] |= int(agent.orientation) << K.ST_ORIENT_SHIFT
        if agent.id in done_agents:
            flags[i] |= K.ST_DONE_REPORTED
    for r in range(sim.grid.rows):
        for c in range(cols):
            occupants = sim.grid[r, c]
            if not occupants:
                continue
            order = [index[a] for a in occupants]          # dict insertion order = arrival order
            for j, a in enumerate(order):
                flags[a] |= K.ST_IN_GRID
                assert cell[a] == r * cols + c, "grid / position mismatch in the reference sim"
                nxt[a] = order[j + 1] if j + 1 < len(order) else K.BGW_NONE
    return dict(cell=cell, next=nxt, flags=flags, health=health, ammo=ammo)
