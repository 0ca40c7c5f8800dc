"""GridWorld agent types -- host-side declarations mirroring abmarl/sim/gridworld/agent.py.

Constructor arguments, attribute names and validation rules follow the reference (cited per class); the
objects carry configuration only.  Per-env, per-step values (position, health, orientation, active) live in
the engine's HBM state arrays, not on these objects.
"""
import numpy as np

from abmarl_b200.sim import PrincipleAgent, ActingAgent, ObservingAgent


class _Checked:
    """Validated attribute: `check(value)` must hold (None allowed when `optional`)."""

    def __init__(self, check, message, optional=False):
        self.check, self.message, self.optional = check, message, optional

    def __set_name__(self, owner, name):
        self.slot = '_' + name

    def __get__(self, obj, owner=None):
        return self if obj is None else getattr(obj, self.slot)

    def __set__(self, obj, value):
        if not (self.optional and value is None):
            assert self.check(value), f"{getattr(obj, '_id', '?')}: {self.message}"
        setattr(obj, self.slot, value)


def _is_int(v):
    return type(v) is int


def _is_num(v):
    return type(v) in (int, float)


class GridWorldAgent(PrincipleAgent):
    """agent.py:6-112.  encoding must be an int other than -2 (masked), -1 (out of bounds), 0 (empty)."""
    encoding = _Checked(lambda v: _is_int(v) and v not in (-2, -1, 0),
                        "encoding must be an integer other than -2, -1, 0.")
    initial_position = _Checked(
        lambda v: type(v) is np.ndarray and v.shape == (2,) and v.dtype in [int, float],
        "Initial position must be a 2-element numerical numpy array.", optional=True)
    blocking = _Checked(lambda v: type(v) is bool, "Blocking must be either True or False.")
    render_size = _Checked(lambda v: _is_int(v) and v > 0, "Render size must be a positive integer.")

    def __init__(self, initial_position=None, blocking=False, encoding=None, render_shape='o',
                 render_color='gray', render_size=200, **kwargs):
        super().__init__(**kwargs)
        self.encoding = encoding
        self.initial_position = initial_position
        self.blocking = blocking
        self.render_shape, self.render_color, self.render_size = render_shape, render_color, render_size

    @property
    def configured(self):
        return super().configured and self.encoding is not None and self.blocking is not None


class GridObservingAgent(ObservingAgent, GridWorldAgent):
    """agent.py:115-132"""
    view_range = _Checked(lambda v: _is_int(v) and v >= 0, "View range must be a nonnegative integer.")

    def __init__(self, view_range=None, **kwargs):
        super().__init__(**kwargs)
        self.view_range = view_range


class MovingAgent(ActingAgent, GridWorldAgent):
    """agent.py:135-157"""
    move_range = _Checked(lambda v: _is_int(v) and v >= 0, "Move range must be a nonnegative integer.")

    def __init__(self, move_range=None, **kwargs):
        super().__init__(**kwargs)
        self.move_range = move_range


class HealthAgent(GridWorldAgent):
    """agent.py:160-196.  Health is float64 in [0,1]; the entity is active while health > 0."""
    initial_health = _Checked(lambda v: _is_num(v) and 0 < v <= 1,
                              "Initial health must be a number in (0, 1].", optional=True)

    def __init__(self, initial_health=None, **kwargs):
        super().__init__(**kwargs)
        self.initial_health = initial_health


class AttackingAgent(ActingAgent, GridWorldAgent):
    """agent.py:199-288"""
    attack_range = _Checked(lambda v: _is_int(v) and v >= 0, "Attack range must be a nonnegative integer.")
    attack_strength = _Checked(lambda v: _is_num(v) and 0 <= v <= 1, "Attack strength must be in [0, 1].")
    attack_accuracy = _Checked(lambda v: _is_num(v) and 0 <= v <= 1, "Attack accuracy must be in [0, 1].")
    simultaneous_attacks = _Checked(lambda v: _is_int(v) and v >= 0,
                                    "Simultaneous attacks must be a nonnegative integer.")

    def __init__(self, attack_range=None, attack_strength=None, attack_accuracy=None,
                 simultaneous_attacks=1, attack_count=None, **kwargs):
        super().__init__(**kwargs)
        self.attack_range = attack_range
        self.attack_strength = attack_strength
        self.attack_accuracy = attack_accuracy
        # `attack_count` is the pre-0.2.6 spelling (docs/src/release.rst:92-93)
        self.simultaneous_attacks = simultaneous_attacks if attack_count is None else attack_count

    @property
    def attack_count(self):
        return self.simultaneous_attacks


class AmmoAgent(GridWorldAgent):
    """agent.py:291-322.  `ammo` lives in BgwState.ammo on the device: AmmoState.reset gives initial_ammo, every
    agent an attack names costs one round (actor.py:343-351)."""
    initial_ammo = _Checked(_is_int, "Initial ammo must be an integer.")

    def __init__(self, initial_ammo=None, **kwargs):
        super().__init__(**kwargs)
        self.initial_ammo = initial_ammo


class _AmmoObservingMeta(type(GridWorldAgent)):
    """agent.py:324-332: anything that is both an AmmoAgent and an ObservingAgent."""

    def __instancecheck__(cls, instance):
        return isinstance(instance, ObservingAgent) and isinstance(instance, AmmoAgent)


class AmmoObservingAgent(AmmoAgent, ObservingAgent, metaclass=_AmmoObservingMeta):
    """agent.py:335-339: boilerplate required by the AmmoObserver."""


class OrientationAgent(GridWorldAgent):
    """agent.py:342-373.  1: Left, 2: Down, 3: Right, 4: Up."""
    initial_orientation = _Checked(lambda v: v in range(1, 5), "Orientation must be 1, 2, 3, or 4.",
                                   optional=True)

    def __init__(self, initial_orientation=None, **kwargs):
        super().__init__(**kwargs)
        self.initial_orientation = initial_orientation
