"""Example simulation definitions.

Each class names the built-in device *program* that reproduces the reference sim's hand-written
step()/get_reward()/get_done()/get_all_done(), and carries that sim's reward constants:

  TeamBattleSim           abmarl/examples/sim/team_battle_example.py:23-59
  MazeNavigationSim       abmarl/examples/sim/maze_navigation.py:14-42
  MultiMazeNavigationSim  abmarl/examples/sim/multi_maze_navigation.py:17-74
  PacmanSim               abmarl/examples/sim/pacman.py:29-151
  ReachTheTargetSim       abmarl/examples/sim/reach_the_target.py:90-176
"""
from abmarl_b200.sim.gridworld.smart import SmartGridWorldSimulation
from abmarl_b200.sim.agent_based_simulation import DynamicOrderSimulation
from abmarl_b200.sim.gridworld.base import GridWorldSimulation
from abmarl_b200.sim.gridworld.agent import (
    GridObservingAgent, MovingAgent, AttackingAgent, HealthAgent, OrientationAgent, GridWorldAgent,
)
from abmarl_b200.sim.gridworld.actor import MoveActor, BinaryAttackActor, DriftMoveActor, SelectiveAttackActor
from abmarl_b200.sim.gridworld.state import PositionState, HealthState
from abmarl_b200.sim.gridworld.done import ActiveDone, DoneBaseComponent
from abmarl_b200.sim.gridworld.state import MazePlacementState
from abmarl_b200.sim.gridworld.observer import PositionCenteredEncodingObserver


class BattleAgent(GridObservingAgent, MovingAgent, AttackingAgent, HealthAgent):
    """team_battle_example.py:11-20"""

    def __init__(self, **kwargs):
        super().__init__(move_range=1, attack_range=1, attack_strength=1, attack_accuracy=1,
                         view_range=3, **kwargs)


class TeamBattleSim(SmartGridWorldSimulation):
    reward_constants = dict(attack_fail=-0.1, kill=1.0, die=-1.0, move_fail=-0.1, entropy=-0.01)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.move_actor = MoveActor(**kwargs)
        self.attack_actor = BinaryAttackActor(**kwargs)
        self.finalize()

    def program(self):
        return 'team_battle'


class MazeNavigationAgent(GridObservingAgent, MovingAgent):
    """maze_navigation.py:9-11"""

    def __init__(self, **kwargs):
        super().__init__(move_range=1, **kwargs)


class MazeNavigationSim(SmartGridWorldSimulation):
    reward_constants = dict(move_fail=-0.1, target=1.0, entropy=-0.01)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.navigator = self.agents['navigator']
        self.target = self.agents['target']
        self.move_actor = MoveActor(**kwargs)
        self.finalize()

    def program(self):
        return 'maze'


class MultiMazeNavigationAgent(GridObservingAgent, MovingAgent):
    """multi_maze_navigation.py:11-13"""

    def __init__(self, **kwargs):
        super().__init__(move_range=1, **kwargs)


class MultiMazeNavigationSim(GridWorldSimulation):
    reward_constants = dict(move_fail=-0.1, target=1.0, entropy=-0.01)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.position_state = MazePlacementState(**kwargs)
        self.move_actor = MoveActor(**kwargs)
        self.grid_observer = PositionCenteredEncodingObserver(**kwargs)
        self.finalize()

    def program(self):
        return 'multi_maze'


class DynamicOrderMultiMazeSim(MultiMazeNavigationSim, DynamicOrderSimulation):
    """A DynamicOrderSimulation over the grid world (the reference ships the manager, managers/dynamic_order_manager.py,
    but no grid-world sim that uses it): the navigators of MultiMazeNavigationSim take turns, the simulation skips those
    that have reached the target.  Rule (what the step kernel implements, include/bgw.h BGW_MANAGER_DYNAMIC_ORDER, and what
    the reference-side twin in tests/scenarios.py spells out in Python): after a step, next_agent = the agent that just
    acted if that step took it to the target (it expects its last observation, reward and done), followed by the next
    navigator in dict order that is not at the target; reset names the first navigator."""


class PacmanAgent(MovingAgent, OrientationAgent, GridObservingAgent, HealthAgent):
    """pacman.py:11-13"""

    def __init__(self, **kwargs):
        kwargs.setdefault('view_range', 100)
        super().__init__(move_range=1, initial_health=1, **kwargs)


class WallAgent(GridWorldAgent):
    pass


class FoodAgent(HealthAgent):
    def __init__(self, **kwargs):
        super().__init__(render_size=50, initial_health=1, **kwargs)


class BaddieAgent(MovingAgent, OrientationAgent, GridObservingAgent):
    def __init__(self, **kwargs):
        kwargs.setdefault('view_range', 100)
        super().__init__(move_range=1, **kwargs)


class PacmanSim(SmartGridWorldSimulation):
    """pacman.py:29-78: reward_scheme events 'bad_move', 'entropy', 'eat_food', 'kill', 'die'."""
    default_reward_scheme = {'bad_move': -0.1, 'entropy': 0.01, 'eat_food': 0.1, 'kill': 1, 'die': -1}

    def __init__(self, reward_scheme=None, **kwargs):
        super().__init__(**kwargs)
        self.pacman = self.agents['pacman']
        self.move_actor = DriftMoveActor(**kwargs)
        if reward_scheme is not None:
            assert type(reward_scheme) is dict, "Reward scheme must be a dictionary."
            for event, reward in reward_scheme.items():
                assert event in self.default_reward_scheme, \
                    "Supported events: 'bad_move', 'entropy', 'eat_food', 'kill', and 'die'."
                assert type(reward) in [int, float], f"Reward for {event} must be numerical."
        self.reward_scheme = reward_scheme if reward_scheme is not None else dict(self.default_reward_scheme)
        self.finalize()

    def program(self):
        return 'pacman'


class PacmanSimSimple(PacmanSim):
    """pacman.py:172-326: pacman against baddies that follow a fixed script (the sim `examples/rllib_pacman.py` builds).
    The script names the baddies of the example grid by id; reward_scheme events: 'bad_move', 'entropy', 'eat_food', 'die'."""
    default_reward_scheme = {'bad_move': -0.1, 'entropy': 0.01, 'eat_food': 0.1, 'die': -1}
    SCRIPTED_BADDIES = ('baddie_20', 'baddie_36', 'baddie_156', 'baddie_157', 'baddie_159', 'baddie_161', 'baddie_162',
                        'baddie_206', 'baddie_222', 'baddie_328')                 # pacman.py:235-246, in this order

    def __init__(self, reward_scheme=None, **kwargs):
        if reward_scheme is not None:
            assert type(reward_scheme) is dict, "Reward scheme must be a dictionary."
            for event, reward in reward_scheme.items():
                assert event in self.default_reward_scheme, "Supported events: 'bad_move', 'entropy', 'eat_food', and 'die'."
                assert type(reward) in [int, float], f"Reward for {event} must be numerical."
        super().__init__(reward_scheme=None, **kwargs)
        self.reward_scheme = reward_scheme if reward_scheme is not None else dict(self.default_reward_scheme)

    def program(self):
        return 'pacman_simple'

    class _ExampleGrid:
        def __get__(self, obj, owner):
            import os
            import numpy as np
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'layouts', 'pacman.txt')
            with open(path) as f:
                return np.array([line.split() for line in f if line.strip()])
    example_grid = _ExampleGrid()                        # pacman.py:328-377 (the layout of examples/pacman.txt)


class TargetDone(ActiveDone):
    """reach_the_target.py:14-40: an agent is done when it stands on the target's cell."""

    def __init__(self, target=None, **kwargs):
        super().__init__(**kwargs)
        assert target in self.agents.values(), "Target must be an agent."
        self.target = target


class OnlyAgentLeftDone(DoneBaseComponent):
    """reach_the_target.py:43-57: agent and simulation are done when at most one learning agent is active."""


class BarrierAgent(GridWorldAgent):
    """reach_the_target.py:60-67"""

    def __init__(self, **kwargs):
        super().__init__(encoding=1, blocking=True, render_shape='s', **kwargs)


class TargetAgent(AttackingAgent, GridObservingAgent):
    """reach_the_target.py:70-77"""

    def __init__(self, **kwargs):
        super().__init__(id='target', encoding=2, render_color='g', **kwargs)


class RunningAgent(MovingAgent, GridObservingAgent, HealthAgent):
    """reach_the_target.py:80-87"""

    def __init__(self, **kwargs):
        super().__init__(encoding=3, render_color='b', **kwargs)


class ReachTheTargetSim(GridWorldSimulation):
    """reach_the_target.py:90-176: runners try to reach the target, which shoots at them (SelectiveAttackActor)."""
    reward_constants = dict(attack_fail=-0.1, kill=1.0, die=-1.0, move_fail=-0.1, target=1.0, entropy=-0.01)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.target = self.agents['target']
        self.position_state = PositionState(**kwargs)
        self.health_state = HealthState(**kwargs)
        self.move_actor = MoveActor(**kwargs)
        self.attack_actor = SelectiveAttackActor(**kwargs)
        self.grid_observer = PositionCenteredEncodingObserver(**kwargs)
        self.active_done = ActiveDone(**kwargs)
        self.target_done = TargetDone(target=self.target, **kwargs)
        self.only_agent_done = OnlyAgentLeftDone(**kwargs)
        self.finalize()

    def program(self):
        return 'reach_the_target'
