"""Actor components (declarations) -- mirror of abmarl/sim/gridworld/actor.py.

Constructing an actor assigns the action / null-action entries on the agents it supports exactly as the
reference does (actor.py:58-66,121-126,451-453).  `process_action` itself is the device's ordered actor
resolution (csrc/bgw_fast.cuh: ordered rounds; csrc/bgw_dev.cuh: exec_attack / exec_attack_ext / team_battle_step), not a host
method.
"""
from abc import ABC, abstractmethod

import numpy as np

from abmarl_b200.spaces import Box, Discrete, MultiDiscrete, Dict
from abmarl_b200.sim.gridworld.base import GridWorldBaseComponent
from abmarl_b200.sim.gridworld.agent import MovingAgent, AttackingAgent, OrientationAgent


class ActorBaseComponent(GridWorldBaseComponent, ABC):
    @property
    @abstractmethod
    def key(self):
        """Entry of the action dict this actor consumes."""

    @property
    @abstractmethod
    def supported_agent_type(self):
        """Agent type this actor works with."""


class MoveActor(ActorBaseComponent):
    """actor.py:55-114: Box(-move_range, move_range, (2,), int); null action (0, 0)."""
    key = "move"
    supported_agent_type = MovingAgent

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        for agent in self.agents.values():
            if isinstance(agent, self.supported_agent_type):
                agent.action_space[self.key] = Box(-agent.move_range, agent.move_range, (2,), int)
                agent.null_action[self.key] = np.zeros((2,), dtype=int)


class CrossMoveActor(ActorBaseComponent):
    """actor.py:117-194: Discrete(5) = stay, left, down, right, up."""
    key = "move"
    supported_agent_type = MovingAgent
    GRID_ACTION = {0: (0, 0), 1: (0, -1), 2: (1, 0), 3: (0, 1), 4: (-1, 0)}

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        for agent in self.agents.values():
            if isinstance(agent, self.supported_agent_type):
                agent.action_space[self.key] = Discrete(5)
                agent.null_action[self.key] = 0

    def grid_action(self, cross_action):
        assert cross_action in self.GRID_ACTION, "Cross action must be 0, 1, 2, 3, or 4."
        return np.array(self.GRID_ACTION[cross_action])


class DriftMoveActor(CrossMoveActor):
    """actor.py:197-234: failed / absent turns fall back to drifting along the orientation."""
    drift_agent_types = (OrientationAgent, MovingAgent)


class AttackActorBaseComponent(ActorBaseComponent, ABC):
    """actor.py:237-438"""
    key = 'attack'
    supported_agent_type = AttackingAgent

    def __init__(self, attack_mapping=None, stacked_attacks=False, **kwargs):
        super().__init__(**kwargs)
        assert type(attack_mapping) is dict, "Attack mapping must be dictionary."
        for k, v in attack_mapping.items():
            assert type(k) is int, "All keys in attack mapping must be an integer."
            assert type(v) is set, "All values in attack mapping must be a set."
            assert all(type(i) is int for i in v), "All elements in the attack mapping values must be integers."
        assert type(stacked_attacks) is bool, "Stacked attacks must be a boolean."
        self.attack_mapping, self.stacked_attacks = attack_mapping, stacked_attacks
        for agent in self.agents.values():
            if isinstance(agent, self.supported_agent_type):
                self._assign_space(agent)

    @abstractmethod
    def _assign_space(self, agent):
        pass


class BinaryAttackActor(AttackActorBaseComponent):
    """actor.py:441-501: Discrete(simultaneous_attacks + 1); null action 0."""

    def _assign_space(self, agent):
        agent.action_space[self.key] = Discrete(agent.simultaneous_attacks + 1)
        agent.null_action[self.key] = 0


class EncodingBasedAttackActor(AttackActorBaseComponent):
    """actor.py:504-582: Dict{encoding: Discrete(simultaneous_attacks + 1)} over the encodings the agent can attack;
    the count is an upper bound per encoding."""

    def _assign_space(self, agent):
        attackable_encodings = self.attack_mapping[agent.encoding]
        agent.action_space[self.key] = Dict({i: Discrete(agent.simultaneous_attacks + 1) for i in sorted(attackable_encodings)})
        agent.null_action[self.key] = {i: 0 for i in sorted(attackable_encodings)}


class RestrictedSelectiveAttackActor(AttackActorBaseComponent):
    """actor.py:585-658: MultiDiscrete([cells + 1] * simultaneous_attacks): each entry names one cell of the attack
    window (0 = attack not used, else 1 + ravelled cell with row = (v-1) % n, column = (v-1) // n)."""

    def _assign_space(self, agent):
        grid_cells = (2 * agent.attack_range + 1) ** 2
        agent.action_space[self.key] = MultiDiscrete([grid_cells + 1] * agent.simultaneous_attacks)
        agent.null_action[self.key] = np.zeros((agent.simultaneous_attacks,), dtype=int)


class SelectiveAttackActor(AttackActorBaseComponent):
    """actor.py:661-728: Box(0, simultaneous_attacks, (n, n), int): attacks per cell of the attack window."""

    def _assign_space(self, agent):
        n = 2 * agent.attack_range + 1
        agent.action_space[self.key] = Box(0, agent.simultaneous_attacks, (n, n), int)
        agent.null_action[self.key] = np.zeros((n, n), dtype=int)


AttackActor = BinaryAttackActor   # pre-0.2.6 name (docs/src/release.rst:96-99)
