"""Simulation managers over the batched engine -- mirrors abmarl/managers/{all_step,turn_based,dynamic_order}_manager.py."""
from .managers import SimulationManager, AllStepManager, TurnBasedManager, DynamicOrderManager  # noqa: F401
