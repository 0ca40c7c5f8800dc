"""Traffic corridor example (definitions) -- mirror of abmarl/examples/sim/traffic_corridor.py.

Agents cross a one-cell-wide corridor from their start cells to their targets; the device program reproduces
TrafficCorridorSimulation.step (traffic_corridor.py:46-53): every acting agent moves, -0.1 for a failed move, +1 when
it stands on its target (TargetAgentDone).
"""
from abmarl_b200.sim.gridworld.smart import SmartGridWorldSimulation
from abmarl_b200.sim.gridworld.agent import GridWorldAgent, MovingAgent, GridObservingAgent
from abmarl_b200.sim.gridworld.actor import MoveActor


class WallAgent(GridWorldAgent):
    pass


class TargetAgent(GridWorldAgent):
    pass


class TrafficAgent(MovingAgent, GridObservingAgent):
    """traffic_corridor.py:13-21"""

    def __init__(self, **kwargs):
        super().__init__(view_range=3, move_range=1, **kwargs)


class TrafficCorridorSimulation(SmartGridWorldSimulation):
    reward_constants = dict(move_fail=-0.1, target=1.0)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.move_actor = MoveActor(**kwargs)
        self.finalize()

    def program(self):
        return 'traffic_corridor'
