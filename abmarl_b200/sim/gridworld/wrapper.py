"""Component wrappers -- mirror of abmarl/sim/gridworld/wrapper.py:92-210.

RavelActionWrapper turns MoveActor's Box(-m, m, (2,)) into Discrete((2m+1)^2); on the device the decode is
a = (dr+m)(2m+1) + (dc+m)  (np.unravel_index(a, high+1-low) + low, ravel_discrete_wrapper.py:90-92)
applied in the action-load stage of the step kernel (BgwSpec.ravel_actions).
"""
import numpy as np

from abmarl_b200.spaces import Box, Discrete
from abmarl_b200.sim.gridworld.actor import ActorBaseComponent


def ravel(point, space):
    """Box(int) point -> index (ravel_discrete_wrapper.py:13-70, Box branch)."""
    span = (space.high - space.low + 1).ravel()
    return int(np.ravel_multi_index((np.asarray(point) - space.low).ravel(), span))


def unravel(index, space):
    """index -> Box(int) point (ravel_discrete_wrapper.py:73-106, Box branch)."""
    span = (space.high - space.low + 1).ravel()
    return (np.array(np.unravel_index(index, span)).reshape(space.shape) + space.low).astype(int)


class ActorWrapper(ActorBaseComponent):
    """wrapper.py:92-159"""

    def __init__(self, component):
        assert isinstance(component, ActorBaseComponent), "Wrapped component must be an ActorBaseComponent."
        self._wrapped = component
        self.from_space = {}
        for agent in component.agents.values():
            if isinstance(agent, component.supported_agent_type):
                space = agent.action_space[component.key]
                assert self.check_space(space), "Cannot wrap this space."
                self.from_space[agent.id] = space
                agent.action_space[component.key] = self.wrap_space(space)
                agent.null_action[component.key] = self.wrap_point(space, agent.null_action[component.key])

    @property
    def wrapped_component(self):
        return self._wrapped

    @property
    def unwrapped(self):
        return getattr(self._wrapped, 'unwrapped', self._wrapped)

    @property
    def agents(self):
        return self._wrapped.agents

    @property
    def grid(self):
        return self._wrapped.grid

    @property
    def key(self):
        return self._wrapped.key

    @property
    def supported_agent_type(self):
        return self._wrapped.supported_agent_type


class RavelActionWrapper(ActorWrapper):
    """wrapper.py:180-210"""

    def check_space(self, space):
        return isinstance(space, Discrete) or (isinstance(space, Box) and space.dtype.kind in 'iu')

    def wrap_space(self, space):
        if isinstance(space, Discrete):
            return space
        return Discrete(int(np.prod(space.high - space.low + 1)))

    def unwrap_point(self, space, point):
        return unravel(point, space)

    def wrap_point(self, space, point):
        return ravel(point, space)
