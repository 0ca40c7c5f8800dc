"""Agent hierarchy and the AgentBasedSimulation interface (abmarl/sim/agent_based_simulation.py)."""
from .agent_based_simulation import (  # noqa: F401
    PrincipleAgent, ActingAgent, ObservingAgent, Agent, AgentBasedSimulation, DynamicOrderSimulation,
)
