"""GridWorldSimulation builders and the component base -- mirrors abmarl/sim/gridworld/base.py.

`build_sim`, `build_sim_from_grid`, `build_sim_from_array`, `build_sim_from_file` take the reference's
arguments (base.py:38-198): every keyword is broadcast to every component constructor.  The result is a
simulation *definition*; hand it to a manager (abmarl_b200.managers) to get the batched device env.
"""
from abc import ABC

import numpy as np

from abmarl_b200.sim import AgentBasedSimulation
from abmarl_b200.sim.gridworld.agent import GridWorldAgent
from abmarl_b200.sim.gridworld.grid import Grid

_EMPTY = (0, '.', '_')   # reserved for empty space (base.py:118-119,168-169)


class GridWorldSimulation(AgentBasedSimulation, ABC):
    def __init__(self, grid=None, **kwargs):
        super().__init__(**kwargs)
        assert isinstance(grid, Grid), "Grid must be a Grid object."
        self.grid = grid

    # -- builders ------------------------------------------------------------------------------
    @classmethod
    def build_sim(cls, rows, cols, **kwargs):
        """base.py:38-59"""
        assert type(rows) is int and rows > 0, "Rows must be a positive integer."
        assert type(cols) is int and cols > 0, "Cols must be a positive integer."
        return cls._build_sim(rows, cols, **kwargs)

    @classmethod
    def build_sim_from_grid(cls, grid, extra_agents=None, **kwargs):
        """base.py:61-96.  `grid` here is a mapping {(r, c): [agents...]} or an object array of dicts."""
        agents = cls._extra(extra_agents)
        cells = grid.items() if isinstance(grid, dict) else (
            ((r, c), grid[r, c]) for r in range(grid.shape[0]) for c in range(grid.shape[1]))
        rows = cols = None
        if not isinstance(grid, dict):
            rows, cols = grid.shape
        for (r, c), occupants in cells:
            if not occupants:
                continue
            occupants = occupants.values() if isinstance(occupants, dict) else occupants
            for agent in occupants:
                np.testing.assert_array_equal(agent.initial_position, np.array([r, c]))
                agents[agent.id] = agent
        assert rows is not None, "Pass an object array so the grid shape is known."
        return cls._build_sim(rows, cols, agents=agents, **kwargs)

    @classmethod
    def build_sim_from_array(cls, array, object_registry, extra_agents=None, **kwargs):
        """base.py:98-142"""
        assert type(array) is np.ndarray, "The array must be a numpy array."
        cls._check_registry(object_registry)
        agents = cls._extra(extra_agents)
        rows, cols = array.shape[0], array.shape[1]
        cls._populate(agents, object_registry,
                      ((r, c, array[r, c]) for r in range(rows) for c in range(cols)))
        return cls._build_sim(rows, cols, agents=agents, **kwargs)

    @classmethod
    def build_sim_from_file(cls, file_name, object_registry, extra_agents=None, **kwargs):
        """base.py:144-192: space-separated characters, one row per line."""
        assert type(file_name) is str, "The file_name must be the name of the file."
        cls._check_registry(object_registry)
        agents = cls._extra(extra_agents)
        with open(file_name, 'r') as fp:
            lines = [line.split(' ') for line in fp.read().splitlines()]
        rows, cols = len(lines), len(lines[0])
        for chars in lines:
            assert len(chars) == cols, f"Mismatched number of columns per row in {file_name}"
        cls._populate(agents, object_registry,
                      ((r, c, ch) for r, chars in enumerate(lines) for c, ch in enumerate(chars)))
        return cls._build_sim(rows, cols, agents=agents, **kwargs)

    @staticmethod
    def _check_registry(object_registry):
        assert type(object_registry) is dict, "The object_registry must be a dictionary."
        assert all(i not in object_registry for i in _EMPTY), "0, '.', and '_' are reserved for empty space."

    @staticmethod
    def _extra(extra_agents):
        if extra_agents is None:
            return {}
        assert type(extra_agents) is dict, "Extra agents must be a dictionary."
        return extra_agents

    @staticmethod
    def _populate(agents, object_registry, cells):
        n = 0                              # running counter over registered characters only (base.py:178-191)
        for r, c, char in cells:
            if char in object_registry:
                agent = object_registry[char](n)
                agent.initial_position = np.array([r, c])
                agents[agent.id] = agent
                n += 1

    @classmethod
    def _build_sim(cls, rows, cols, **kwargs):
        return cls(grid=Grid(rows, cols, **kwargs), **kwargs)   # base.py:195-198

    def render(self, **kwargs):
        raise NotImplementedError("rendering is out of scope for the batched engine (matplotlib; base.py:200-245)")


class GridWorldBaseComponent(ABC):
    """Every component sees the agents dict and the grid (base.py:248-304)."""

    def __init__(self, agents=None, grid=None, **kwargs):
        assert type(agents) is dict, "Agents must be a dict."
        for agent_id, agent in agents.items():
            assert isinstance(agent, GridWorldAgent), "Values of agents dict must be instance of GridWorldAgent."
            assert agent_id == agent.id, "Keys of agents dict must be the same as the Agent's id."
        assert isinstance(grid, Grid), "The grid must be a Grid object."
        self.agents, self.grid = agents, grid

    @property
    def rows(self):
        return self.grid.rows

    @property
    def cols(self):
        return self.grid.cols
