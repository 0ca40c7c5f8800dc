"""Done components (declarations) -- mirror of abmarl/sim/gridworld/done.py.

Evaluated on the device in the reward/done reduction (csrc/bgw_dev.cuh: prog_done / compute_all_done).
"""
from abc import ABC

from abmarl_b200.sim.gridworld.base import GridWorldBaseComponent
from abmarl_b200.sim.gridworld.agent import GridWorldAgent


class DoneBaseComponent(GridWorldBaseComponent, ABC):
    pass


class ActiveDone(DoneBaseComponent):
    """done.py:39-56: an entity is done when inactive; all done when none is active."""


class _TargetMapped(DoneBaseComponent):
    def __init__(self, target_mapping=None, **kwargs):
        super().__init__(**kwargs)
        assert type(target_mapping) is dict, "Target mapping must be a dictionary."
        for agent_id, target_id in target_mapping.items():
            for x in (agent_id, target_id):
                assert x in self.agents and isinstance(self.agents[x], GridWorldAgent), \
                    f"{x} must be a GridWorldAgent in the simulation."
        self.target_mapping = target_mapping


class TargetAgentDone(_TargetMapped):
    """done.py:59-99: done when the agent stands on its target."""


class TargetDestroyedDone(_TargetMapped):
    """done.py:102-137: done when the agent's target is inactive."""


class OneTeamRemainingDone(ActiveDone):
    """done.py:140-153: all done when the active entities share one encoding."""
