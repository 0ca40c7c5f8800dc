"""SmartGridWorldSimulation -- mirror of abmarl/sim/gridworld/smart.py:10-120.

`states`, `observers`, `dones` are sets of component classes or registered names; every keyword is
broadcast to each constructor.  Actors and the step program are supplied by the subclass.
"""
from abc import ABC

from abmarl_b200.sim.gridworld.base import GridWorldSimulation
from abmarl_b200.sim.gridworld.registry import registry


class SmartGridWorldSimulation(GridWorldSimulation, ABC):
    def __init__(self, states=None, observers=None, dones=None, **kwargs):
        super().__init__(**kwargs)
        for attr, kind, chosen in (('_states', 'state', states), ('_observers', 'observer', observers),
                                   ('_dones', 'done', dones)):
            if not chosen:
                continue
            assert type(chosen) is set, f"{kind}s must be a set of {kind} components"
            built = []
            for item in sorted(chosen, key=lambda x: x if type(x) is str else x.__name__):
                if type(item) is str:
                    assert item in registry[kind], f"{item} is not registered as a {kind}."
                    item = registry[kind][item]
                elif item.__name__ not in registry[kind] and item not in registry[kind].values():
                    raise ValueError(f"{item} must be a {kind} component or the name of a registered one.")
                built.append(item(**kwargs))
            setattr(self, attr, built)
