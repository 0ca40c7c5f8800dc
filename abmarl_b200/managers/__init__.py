"""Simulation managers over the batched engine -- mirrors abmarl/managers/{all_step,turn_based}_manager.py."""
from .managers import SimulationManager, AllStepManager, TurnBasedManager  # noqa: F401
