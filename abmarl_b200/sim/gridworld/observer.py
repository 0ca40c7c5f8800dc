"""Observer components (declarations) -- mirror of abmarl/sim/gridworld/observer.py.

Constructing an observer assigns observation / null-observation entries like the reference
(observer.py:69-79,161-174,262-278).  `get_obs` is the device's observation gather
(csrc/bgw_fast.cuh: fast_obs_rows; csrc/bgw_dev.cuh: obs_chunk / observe_learners).  Reference dtype is int64; the engine emits int8 that compares
equal after widening (max encoding <= 63).
"""
from abc import ABC, abstractmethod

import numpy as np

from abmarl_b200.spaces import Box
from abmarl_b200.sim.gridworld.base import GridWorldBaseComponent
from abmarl_b200.sim.agent_based_simulation import ObservingAgent
from abmarl_b200.sim.gridworld.agent import GridObservingAgent, AmmoObservingAgent


class ObserverBaseComponent(GridWorldBaseComponent, ABC):
    supported_agent_type = GridObservingAgent

    @property
    @abstractmethod
    def key(self):
        """Entry of the observation dict this observer fills."""

    def _shape(self, agent):
        raise NotImplementedError

    def _assign(self, high):
        for agent in self.agents.values():
            if isinstance(agent, self.supported_agent_type):
                shape = self._shape(agent)
                agent.observation_space[self.key] = Box(-2, high, shape, int)
                agent.null_observation[self.key] = -2 * np.ones(shape, dtype=int)

    @property
    def max_encoding(self):
        return max(agent.encoding for agent in self.agents.values())


class AbsoluteEncodingObserver(ObserverBaseComponent):
    """observer.py:55-150: whole grid; -1 = myself, -2 = masked / out of view."""
    key = 'absolute_encoding'

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._assign(self.max_encoding)

    def _shape(self, agent):
        return (self.rows, self.cols)


class PositionCenteredEncodingObserver(ObserverBaseComponent):
    """observer.py:153-250: (2R+1)^2 window centred on the agent; -1 out of bounds, -2 masked."""
    key = 'position_centered_encoding'

    def __init__(self, observe_self=True, **kwargs):
        super().__init__(**kwargs)
        assert type(observe_self) is bool, "Observe self must be a boolean."
        self.observe_self = observe_self
        self._assign(self.max_encoding)

    def _shape(self, agent):
        return (agent.view_range * 2 + 1, agent.view_range * 2 + 1)


class StackedPositionCenteredEncodingObserver(ObserverBaseComponent):
    """observer.py:253-334: one count channel per encoding."""
    key = 'stacked_position_centered_encoding'

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.number_of_encodings = self.max_encoding
        self._assign(len(self.agents))

    def _shape(self, agent):
        return (agent.view_range * 2 + 1, agent.view_range * 2 + 1, self.number_of_encodings)


class AbsolutePositionObserver(ObserverBaseComponent):
    """observer.py:337-373: agents observe their absolute position (two int16 at BgwDims.position_offset of the obs row)."""
    key = 'position'
    supported_agent_type = ObservingAgent

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        for agent in self.agents.values():
            if isinstance(agent, self.supported_agent_type):
                agent.observation_space[self.key] = Box(np.array([0, 0], dtype=int),
                                                        np.array([self.rows - 1, self.cols - 1], dtype=int), dtype=int)
                agent.null_observation[self.key] = np.zeros((2,), dtype=int)


class AmmoObserver(ObserverBaseComponent):
    """observer.py:376-413: agents observe their own ammo (an int32 slot at BgwDims.ammo_offset of the obs row)."""
    key = 'ammo'
    supported_agent_type = AmmoObservingAgent

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        for agent in self.agents.values():
            if isinstance(agent, self.supported_agent_type):
                agent.observation_space[self.key] = Box(0, agent.initial_ammo, (1,), int)
                agent.null_observation[self.key] = 0


# pre-0.2.6 names (docs/src/release.rst:96-99; still used by tests/sim/gridworld/test_observer.py)
SingleGridObserver = PositionCenteredEncodingObserver
MultiGridObserver = StackedPositionCenteredEncodingObserver
