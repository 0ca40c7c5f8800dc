"""State components (declarations) -- mirror of abmarl/sim/gridworld/state.py.

`reset()` of each component runs on the device (bgw_reset; csrc/bgw_dev.cuh: sim_reset):
PositionState state.py:88-166, HealthState :629-641, OrientationState :666-675.
MazePlacementState (:385-619) and TargetBarriersFreePlacementState (:169-383) build their layouts on the device
(bgw_generate_layouts, csrc/bgw_maze.cuh) or, above its limits, host-side (abmarl_b200.layouts).
"""
from abc import ABC

from abmarl_b200.sim.gridworld.base import GridWorldBaseComponent
from abmarl_b200.sim.gridworld.agent import GridWorldAgent


class StateBaseComponent(GridWorldBaseComponent, ABC):
    pass


class PositionState(StateBaseComponent):
    def __init__(self, no_overlap_at_reset=False, randomize_placement_order=False, **kwargs):
        super().__init__(**kwargs)
        assert type(no_overlap_at_reset) is bool, "No overlap at reset must be a boolean."
        assert type(randomize_placement_order) is bool, "Randomize placement order must be True or False."
        self.no_overlap_at_reset = no_overlap_at_reset
        self.randomize_placement_order = randomize_placement_order


class TargetBarriersFreePlacementState(PositionState):
    """state.py:169-383: the target first, barrier encodings clustered near it, free encodings scattered away from it."""

    def __init__(self, target_agent=None, barrier_encodings=None, free_encodings=None,
                 cluster_barriers=False, scatter_free_agents=False, **kwargs):
        super().__init__(**kwargs)
        if type(target_agent) is str:
            assert target_agent in self.agents, "The target agent must be an agent in the simulation."
            target_agent = self.agents[target_agent]
        assert isinstance(target_agent, GridWorldAgent) and target_agent.id in self.agents, \
            "The target agent must be an agent in the simulation."
        self.target_agent = target_agent
        for name, value in (('barrier', barrier_encodings), ('free', free_encodings)):
            if value is not None:
                assert type(value) is set and all(type(e) is int for e in value), \
                    f"{name} encodings must be a set of integers."
        self.barrier_encodings = barrier_encodings or set()
        self.free_encodings = free_encodings or set()
        assert type(cluster_barriers) is bool, "Cluster barriers must be a boolean."
        assert type(scatter_free_agents) is bool, "Scatter free agents must be a boolean."
        self.cluster_barriers, self.scatter_free_agents = cluster_barriers, scatter_free_agents


class MazePlacementState(TargetBarriersFreePlacementState):
    """state.py:385-619: the same placement rules over a maze grown from the target (walls = barrier cells, passages =
    free cells).  (In the reference the two classes are siblings with identical constructors.)"""


class HealthState(StateBaseComponent):
    pass


class OrientationState(StateBaseComponent):
    pass


class AmmoState(StateBaseComponent):
    pass
