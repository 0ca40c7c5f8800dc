"""FlattenWrapper / FlattenActionWrapper -- mirror of abmarl/sim/wrappers/flatten_wrapper.py over the batched engine.

`flatdim / flatten / unflatten / flatten_space` restate the reference's functions (flatten_wrapper.py:12-153) on
abmarl_b200.spaces (Box, Discrete, MultiDiscrete, Dict: the spaces the GridWorld components build; the golden vectors of
tests/golden/flatten_golden.json come from the reference's own functions).  `FlattenWrapper(sim)` is the definition
(same constructor, the agents' spaces replaced by flat Boxes, null observation / action flattened, :175-204);
`FlattenView` applies it to the engine's tensors:

    observations  the flattened point of {'<grid observer key>': (n, n[, c]) ints[, 'ammo': int]} is the first n*n*c bytes of
                  the learner's observation row followed by its ammo -- a slice of the int8 rows the step kernel wrote
                  (a view, no copy) when there is no ammo channel
    actions       a flat integer Box point [move..., attack...] in the order of the agent's action Dict is scattered into
                  the engine's action row (include/bgw.h, bgw_step), i.e. wrap_action = unflatten (:200-201) + the
                  managers' encode_actions in one indexing operation
"""
import copy
from collections import OrderedDict

import numpy as np
import torch

from abmarl_b200 import _capi as K
from abmarl_b200.spaces import Box, Discrete, MultiDiscrete, Dict
from abmarl_b200.sim import Agent


def flatdim(space):                                                    # flatten_wrapper.py:12-34
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return 1
    if isinstance(space, MultiDiscrete):
        return len(space)
    if isinstance(space, Dict):
        return int(sum(flatdim(s) for s in space.spaces.values()))
    raise TypeError(f"flatdim: unsupported space {space!r}")


def flatten(space, point):                                             # :37-63
    if isinstance(space, Box):
        return np.asarray(point, dtype=space.dtype).flatten()
    if isinstance(space, Discrete):
        return np.array([point], dtype=int)
    if isinstance(space, MultiDiscrete):
        return point
    if isinstance(space, Dict):
        return np.concatenate([flatten(s, point[key]) for key, s in space.spaces.items()])
    raise TypeError(f"flatten: unsupported space {space!r}")


def unflatten(space, point):                                           # :66-105
    if isinstance(space, Box):
        return np.asarray(point, dtype=space.dtype).reshape(space.shape)
    if isinstance(space, Discrete):
        return point[0]
    if isinstance(space, MultiDiscrete):
        return point
    if isinstance(space, Dict):
        dims = [flatdim(s) for s in space.spaces.values()]
        parts = np.split(point, np.cumsum(dims)[:-1])
        return OrderedDict((key, unflatten(s, part)) for part, (key, s) in zip(parts, space.spaces.items()))
    raise TypeError(f"unflatten: unsupported space {space!r}")


def flatten_space(space):                                              # :108-153
    if isinstance(space, Box):
        return Box(space.low.flatten(), space.high.flatten(), dtype=space.dtype)
    if isinstance(space, Discrete):
        return Box(low=0, high=space.n - 1, shape=(1,), dtype=int)
    if isinstance(space, MultiDiscrete):
        return Box(low=np.zeros_like(space.nvec), high=space.nvec - 1, dtype=int)
    if isinstance(space, Dict):
        parts = [flatten_space(s) for s in space.spaces.values()]
        dtype = int if all(p.dtype == int for p in parts) else float
        return Box(low=np.concatenate([p.low for p in parts]), high=np.concatenate([p.high for p in parts]), dtype=dtype)
    raise TypeError(f"flatten_space: unsupported space {space!r}")


class _SARWrapper:
    """The part of sar_wrapper.py the flatten wrappers use: a copy of the agents dict whose spaces can be replaced."""

    def __init__(self, sim):
        self.sim = sim
        self.agents = {agent_id: copy.copy(agent) for agent_id, agent in sim.agents.items()}

    @property
    def unwrapped(self):
        return self.sim.unwrapped if hasattr(self.sim, 'unwrapped') else self.sim


def _is_null(x):
    """`if agent.null_observation:` of the reference for the values the components produce (dicts, arrays, scalars)."""
    if x is None:
        return False
    if isinstance(x, dict):
        return len(x) > 0
    return bool(np.any(np.asarray(x)))


class FlattenWrapper(_SARWrapper):
    def __init__(self, sim):
        super().__init__(sim)
        for agent_id, wrapped in sim.agents.items():                   # :177-192
            if not isinstance(wrapped, Agent):
                continue
            agent = self.agents[agent_id]
            agent.action_space = flatten_space(wrapped.action_space)
            agent.observation_space = flatten_space(wrapped.observation_space)
            if _is_null(getattr(wrapped, 'null_observation', None)):
                agent.null_observation = flatten(wrapped.observation_space, wrapped.null_observation)
            if _is_null(getattr(wrapped, 'null_action', None)):
                agent.null_action = flatten(wrapped.action_space, wrapped.null_action)

    def wrap_observation(self, from_agent, observation):               # :194-204
        return flatten(from_agent.observation_space, observation)

    def unwrap_observation(self, from_agent, observation):
        return unflatten(from_agent.observation_space, observation)

    def wrap_action(self, from_agent, action):
        return unflatten(from_agent.action_space, action)

    def unwrap_action(self, from_agent, action):
        return flatten(from_agent.action_space, action)


class FlattenActionWrapper(_SARWrapper):
    def __init__(self, sim):                                           # :211-221
        super().__init__(sim)
        for agent_id, wrapped in sim.agents.items():
            if not isinstance(wrapped, Agent):
                continue
            agent = self.agents[agent_id]
            agent.action_space = flatten_space(wrapped.action_space)
            if _is_null(getattr(wrapped, 'null_action', None)):
                agent.null_action = flatten(wrapped.action_space, wrapped.null_action)

    def wrap_action(self, from_agent, action):
        return unflatten(from_agent.action_space, action)

    def unwrap_action(self, from_agent, action):
        return flatten(from_agent.action_space, action)


class FlattenView:
    """FlattenWrapper over the tensors of a manager (abmarl_b200.managers): flat observations out, flat actions in."""

    GRID_KEYS = ('position_centered_encoding', 'absolute_encoding', 'stacked_position_centered_encoding')

    def __init__(self, manager):
        self.manager, self.engine, self.spec = manager, manager.engine, manager.spec
        sim = manager.sim.unwrapped if hasattr(manager.sim, 'unwrapped') else manager.sim
        eng, sp = self.engine, self.spec
        agents = [sim.agents[aid] for aid in manager.learner_ids]
        h, w, c, stride = sp.obs_shape()
        # ---- observations: per learner, the row bytes (and the ammo channel) in the order of its observation Dict
        obs_cols = []
        for l, agent in enumerate(agents):
            a = sp.learner_agents[l]
            cols = []
            for key, space in agent.observation_space.spaces.items():
                if key in self.GRID_KEYS:
                    n = flatdim(space)
                    cols += list(range(n))                             # the row holds the (n, n[, c]) window row-major
                elif key == 'ammo':
                    cols.append(-1)                                    # taken from the ammo view
                else:
                    raise ValueError(f"FlattenView: observation channel '{key}' of agent {agent.id} is not produced by the engine")
            obs_cols.append(cols)
            assert len(cols) == flatdim(agent.observation_space)
            assert all(x < stride for x in cols), (a, stride)
        self.obs_dim = max(len(c_) for c_ in obs_cols) if obs_cols else 0
        self.obs_uniform = all(len(c_) == self.obs_dim for c_ in obs_cols)
        self.obs_has_ammo = any(-1 in c_ for c_ in obs_cols)
        pad = np.zeros((len(agents), self.obs_dim), dtype=np.int64)
        valid = np.zeros((len(agents), self.obs_dim), dtype=bool)
        for l, cols in enumerate(obs_cols):
            pad[l, :len(cols)] = cols
            valid[l, :len(cols)] = True
        self._obs_cols = torch.from_numpy(pad).to(eng.device)
        self._obs_valid = torch.from_numpy(valid).to(eng.device)
        self._obs_prefix = self.obs_uniform and not self.obs_has_ammo and all(c_ == list(range(self.obs_dim)) for c_ in obs_cols)
        # ---- actions: per learner, the engine's action-row byte of every flat component
        act_cols = []
        for l, agent in enumerate(agents):
            cols = []
            for key, space in agent.action_space.spaces.items():
                n = flatdim(space)
                if key == 'move':
                    assert n == (2 if sp.move_actor == K.MOVE_BOX and not sp.ravel_actions else 1)
                    cols += list(range(n))
                elif key == 'attack':
                    if sp.attack_actor == K.ATTACK_ENCODING:           # Dict {encoding: Discrete}: byte 2 + encoding - 1
                        cols += [2 + int(enc) - 1 for enc in space.spaces.keys()]
                    else:                                              # binary / restricted / selective: bytes 2..
                        cols += [2 + j for j in range(n)]
                else:
                    raise ValueError(f"FlattenView: action channel '{key}' of agent {agent.id} is not consumed by the engine")
            assert len(cols) == flatdim(agent.action_space) and all(x < eng.action_stride for x in cols)
            act_cols.append(cols)
        self.act_dim = max(len(c_) for c_ in act_cols) if act_cols else 0
        apad = np.zeros((len(agents), self.act_dim), dtype=np.int64)
        avalid = np.zeros((len(agents), self.act_dim), dtype=bool)
        for l, cols in enumerate(act_cols):
            apad[l, :len(cols)] = cols
            avalid[l, :len(cols)] = True
        self._act_cols = torch.from_numpy(apad).to(eng.device)
        self._act_valid = torch.from_numpy(avalid).to(eng.device)

    def observations(self):
        """[E, L, obs_dim] flattened observations of the last reset / step (rows of learners that received nothing are
        whatever the engine left there, as in the dense outputs).  int8 view of the engine's rows when every learner's
        flat observation is a prefix of its row; otherwise int64 with the ammo channel gathered in (0 beyond a learner's
        own length)."""
        obs = self.engine.obs
        if self._obs_prefix:
            return obs[..., :self.obs_dim]
        E = obs.shape[0]
        cols = self._obs_cols[None].expand(E, -1, -1)
        out = torch.gather(obs.long(), 2, cols.clamp(min=0))
        if self.obs_has_ammo:
            ammo = self.engine.ammo_view().long()[..., None].expand_as(out)
            out = torch.where(cols < 0, ammo, out)
        return torch.where(self._obs_valid[None], out, torch.zeros_like(out))

    def encode_actions(self, flat):
        """[E, L, act_dim] integer points of the flattened action spaces -> int8 [E, L, action_stride] for manager.step
        (components beyond a learner's own length are ignored)."""
        eng = self.engine
        flat = torch.as_tensor(flat, device=eng.device).long()
        E = flat.shape[0]
        rows = torch.zeros((E, eng.L, eng.action_stride), dtype=torch.int64, device=eng.device)
        cols = self._act_cols[None].expand(E, -1, -1)
        vals = torch.where(self._act_valid[None], flat, torch.zeros_like(flat))
        # padded components point at byte 0 with value 0: scatter_add leaves the real byte-0 component intact
        rows.scatter_add_(2, cols, vals)
        return rows.to(torch.int8)
