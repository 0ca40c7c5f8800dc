"""Name -> component registry, as abmarl/sim/gridworld/registry.py:52-77."""
from abmarl_b200.sim.gridworld.state import StateBaseComponent
from abmarl_b200.sim.gridworld.observer import ObserverBaseComponent
from abmarl_b200.sim.gridworld.done import DoneBaseComponent
from abmarl_b200.sim.gridworld import state as _state, observer as _observer, done as _done


def _collect(module, base):
    return {name: obj for name, obj in vars(module).items()
            if isinstance(obj, type) and issubclass(obj, base) and obj is not base and not name.startswith('_')}


registry = {
    'state': _collect(_state, StateBaseComponent),
    'observer': _collect(_observer, ObserverBaseComponent),
    'done': _collect(_done, DoneBaseComponent),
}


def register(component):
    """registry.py:58-77"""
    for kind, base in (('state', StateBaseComponent), ('observer', ObserverBaseComponent), ('done', DoneBaseComponent)):
        if issubclass(component, base):
            registry[kind][component.__name__] = component
            return
    raise TypeError("Component must be a state, observer, or done component.")
