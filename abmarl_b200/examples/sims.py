"""Example simulation definitions.

Each class names the built-in device *program* that reproduces the reference sim's hand-written
step()/get_reward()/get_done()/get_all_done(), and carries that sim's reward constants:

  TeamBattleSim           abmarl/examples/sim/team_battle_example.py:23-59
  MazeNavigationSim       abmarl/examples/sim/maze_navigation.py:14-42
  MultiMazeNavigationSim  abmarl/examples/sim/multi_maze_navigation.py:17-74
  PacmanSim               abmarl/examples/sim/pacman.py:29-151
"""
from abmarl_b200.sim.gridworld.smart import SmartGridWorldSimulation
from abmarl_b200.sim.gridworld.base import GridWorldSimulation
from abmarl_b200.sim.gridworld.agent import (
    GridObservingAgent, MovingAgent, AttackingAgent, HealthAgent, OrientationAgent, GridWorldAgent,
)
from abmarl_b200.sim.gridworld.actor import MoveActor, BinaryAttackActor, DriftMoveActor
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
