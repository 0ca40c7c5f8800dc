"""Example simulations (definitions): the four sims named by the BASELINE configs and ReachTheTargetSim."""
from .sims import (  # noqa: F401
    BattleAgent, TeamBattleSim,
    MazeNavigationAgent, MazeNavigationSim,
    MultiMazeNavigationAgent, MultiMazeNavigationSim, DynamicOrderMultiMazeSim,
    PacmanAgent, WallAgent, FoodAgent, BaddieAgent, PacmanSim, PacmanSimSimple,
    BarrierAgent, TargetAgent, RunningAgent, ReachTheTargetSim, TargetDone, OnlyAgentLeftDone,
)
