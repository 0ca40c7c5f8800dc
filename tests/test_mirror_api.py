"""Host-side mirror of the reference's component API: same names (plus the pre-0.2.6 aliases), same spaces and
null actions, same validation, and the builders compile to the flat spec (no GPU needed)."""
import numpy as np
import pytest

from abmarl_b200 import _capi as K
from abmarl_b200.spaces import Box, Discrete
from abmarl_b200.spec import compile_sim
from abmarl_b200.sim.gridworld import actor, agent, observer, state, done, wrapper
from abmarl_b200.sim.gridworld.grid import Grid
from abmarl_b200 import examples as ex
from tests import scenarios


def test_old_and_new_component_names():
    assert observer.SingleGridObserver is observer.PositionCenteredEncodingObserver
    assert observer.MultiGridObserver is observer.StackedPositionCenteredEncodingObserver
    assert actor.AttackActor is actor.BinaryAttackActor
    for name in ('ActiveDone', 'TargetAgentDone', 'OneTeamRemainingDone'):
        assert hasattr(done, name)
    for name in ('PositionState', 'HealthState', 'OrientationState', 'MazePlacementState'):
        assert hasattr(state, name)


def test_grid_overlapping_is_symmetrised():                    # test_grid.py:30-32
    g = Grid(3, 3, overlapping={1: {2}, 3: {1}})
    assert g.overlapping == {1: {2, 3}, 2: {1}, 3: {1}}
    with pytest.raises(AssertionError):
        Grid(0, 3)
    with pytest.raises(AssertionError):
        Grid(3, 3, overlapping={1: [2]})


def test_agent_validation():                                   # test_agent.py
    with pytest.raises(AssertionError):
        agent.GridWorldAgent(id='a', encoding=0)
    with pytest.raises(AssertionError):
        agent.GridWorldAgent(id='a', encoding=-2)
    a = agent.HealthAgent(id='a', encoding=1, initial_health=0.5)
    assert a.initial_health == 0.5
    with pytest.raises(AssertionError):
        agent.MovingAgent(id='m', encoding=1, move_range=-1)


def test_actor_spaces_and_ravel_wrapper():                     # test_actor.py:36-48, test_wrapper.py:66-144
    agents = {f'agent{i}': agent.MovingAgent(id=f'agent{i}', encoding=1, move_range=i + 1,
                                             initial_position=np.array([i, i])) for i in range(3)}
    grid = Grid(8, 8)
    mv = actor.MoveActor(agents=agents, grid=grid)
    assert agents['agent0'].action_space['move'] == Box(-1, 1, (2,), int)
    assert agents['agent2'].action_space['move'] == Box(-3, 3, (2,), int)
    rv = wrapper.RavelActionWrapper(mv)
    assert [agents[f'agent{i}'].action_space['move'] for i in range(3)] == [Discrete(9), Discrete(25), Discrete(49)]
    assert [agents[f'agent{i}'].null_action['move'] for i in range(3)] == [4, 12, 24]
    assert list(rv.unwrap_point(rv.from_space['agent0'], 7)) == [1, 0]
    assert list(rv.unwrap_point(rv.from_space['agent1'], 3)) == [-2, 1]
    assert list(rv.unwrap_point(rv.from_space['agent2'], 34)) == [1, 3]
    assert rv.wrap_point(rv.from_space['agent2'], np.array([1, 3])) == 34


def test_attack_actor_validation():                            # test_actor.py:503-528
    agents = {'a': agent.AttackingAgent(id='a', encoding=1, attack_range=1, attack_strength=1, attack_accuracy=1,
                                        initial_position=np.array([0, 0]))}
    grid = Grid(3, 3)
    for bad in ([1, 2, 3], {'1': {3}}, {1: 3}, {1: {'2'}}):
        with pytest.raises(AssertionError):
            actor.BinaryAttackActor(agents=agents, grid=grid, attack_mapping=bad)
    actor.BinaryAttackActor(agents=agents, grid=grid, attack_mapping={1: {1}})
    assert agents['a'].action_space['attack'] == Discrete(2) and agents['a'].null_action['attack'] == 0


def test_other_attack_actor_spaces():          # test_actor.py:743-750, 1196-1203, 1528-1535; test_observer (ammo) :376-413
    from abmarl_b200.spaces import Dict, MultiDiscrete
    mk = lambda **kw: {'a': agent.AttackingAgent(id='a', encoding=3, attack_range=2, attack_strength=1, attack_accuracy=1,
                                                 initial_position=np.array([0, 0]), **kw)}
    grid = Grid(5, 6)
    ag = mk()
    actor.SelectiveAttackActor(agents=ag, grid=grid, attack_mapping={3: {1}})
    assert ag['a'].action_space['attack'] == Box(0, 1, (5, 5), int)
    np.testing.assert_array_equal(ag['a'].null_action['attack'], np.zeros((5, 5), dtype=int))
    ag = mk()
    actor.EncodingBasedAttackActor(agents=ag, grid=grid, attack_mapping={3: {1, 2}})
    assert ag['a'].action_space['attack'] == Dict({1: Discrete(2), 2: Discrete(2)})
    assert ag['a'].null_action['attack'] == {1: 0, 2: 0}
    ag = mk(simultaneous_attacks=2)
    ag['a'].attack_range = 1
    actor.RestrictedSelectiveAttackActor(agents=ag, grid=grid, attack_mapping={3: {1, 2}})
    assert ag['a'].action_space['attack'] == MultiDiscrete([10, 10])
    np.testing.assert_array_equal(ag['a'].null_action['attack'], np.zeros((2,), dtype=int))

    class AmmoObserving(agent.AmmoAgent, agent.GridObservingAgent):
        pass
    from abmarl_b200.sim.gridworld import observer
    watchers = {'w': AmmoObserving(id='w', encoding=1, view_range=1, initial_ammo=7), 'x': agent.AmmoAgent(id='x', encoding=1, initial_ammo=2)}
    assert isinstance(watchers['w'], agent.AmmoObservingAgent) and not isinstance(watchers['x'], agent.AmmoObservingAgent)
    observer.AmmoObserver(agents=watchers, grid=grid)
    assert watchers['w'].observation_space['ammo'] == Box(0, 7, (1,), int) and watchers['w'].null_observation['ammo'] == 0


def test_attack_variants_compile_and_encode(mirror):
    """compile_sim + the managers' action encoding for the wider attack actions (layout: include/bgw.h, bgw_step)."""
    spec = compile_sim(scenarios.build_tb_ammo_selective(mirror))
    assert spec.attack_actor == K.ATTACK_SELECTIVE and spec.stacked_attacks == 1 and spec.ammo_observer == 1
    assert all(spec.klass[a] & K.AG_AMMO for a in range(spec.n_agents)) and list(spec.initial_ammo[:4]) == [3, 4, 5, 3]
    assert compile_sim(scenarios.build_tb_encoding(mirror)).attack_actor == K.ATTACK_ENCODING
    assert compile_sim(scenarios.build_tb_restricted(mirror)).attack_actor == K.ATTACK_RESTRICTED


def test_build_sim_from_file_counts_registered_characters(mirror):   # base.py:178-191
    sim = scenarios.build_maze_c1(mirror)
    assert list(sim.agents)[0].startswith('wall') and 'navigator' in sim.agents and 'target' in sim.agents
    spec = compile_sim(sim)
    assert (spec.rows, spec.cols, spec.n_agents, spec.n_learners) == (8, 18, 67, 1)
    assert spec.program == K.PROG_MAZE and spec.role[spec.agent_ids.index('navigator')] == K.ROLE_NAVIGATOR
    blocking = [(spec.klass[i] & K.AG_BLOCKING) != 0 for i in range(spec.n_agents)]
    assert sum(blocking) == 65


def test_team_battle_spec_tables(mirror):
    spec = compile_sim(scenarios.build_tb_c2(mirror), n_envs=16, seed=5, horizon=200, auto_reset=True)
    assert (spec.rows, spec.cols, spec.n_agents, spec.n_learners, spec.max_encoding) == (8, 8, 24, 24, 4)
    assert spec.obs_shape() == (7, 7, 1, 64)
    assert spec.done_mask == K.DONE_ONE_TEAM and spec.attack_actor == K.ATTACK_BINARY and spec.move_actor == K.MOVE_BOX
    assert int(spec.overlap[1]) == 1 << 1 and int(spec.attack_map[1]) == (1 << 2) | (1 << 3) | (1 << 4)
    np.testing.assert_allclose(spec.reward[:5], [-0.1, 1.0, -1.0, -0.1, -0.01])
    assert np.isnan(spec.init_health).all()                     # rllib_team_battle.py leaves initial_health unset


def test_team_battle_rejects_simultaneous_attacks(mirror):     # ValueError in the reference (Binary only)
    agents = {'a0': ex.BattleAgent(id='a0', encoding=1), 'a1': ex.BattleAgent(id='a1', encoding=2)}
    agents['a0'].simultaneous_attacks = 2
    sim = ex.TeamBattleSim.build_sim(4, 4, agents=agents, overlapping={1: {1}}, attack_mapping={1: {2}, 2: {1}},
                                     states={'PositionState', 'HealthState'},
                                     observers={'PositionCenteredEncodingObserver'}, dones={'ActiveDone'})
    with pytest.raises(AssertionError):
        compile_sim(sim)


def test_dynamic_order_manager_wrong_sim(mirror):
    """tests/test_dynamic_order_manager.py:42-44: the manager only takes a DynamicOrderSimulation (the assertion comes
    before anything touches a device)."""
    import pytest
    from abmarl_b200.managers import DynamicOrderManager
    from tests import scenarios
    with pytest.raises(AssertionError):
        DynamicOrderManager(scenarios.build_mm_tiny(mirror))


def test_dynamic_order_simulation_next_agent_property(mirror):
    """tests/sim/test_agent_based_simulation.py: next_agent accepts an id or a container of ids of the sim's agents."""
    import pytest
    from tests import scenarios
    sim = scenarios.build_mm_dynamic(mirror)
    sim.next_agent = 'navigator1'
    assert sim.next_agent == ['navigator1']
    sim.next_agent = ['navigator0', 'navigator2']
    assert sim.next_agent == ['navigator0', 'navigator2']
    with pytest.raises(AssertionError):
        sim.next_agent = ['nobody']
    with pytest.raises(AssertionError):
        sim.next_agent = 3


def test_randomize_action_input_must_be_a_boolean(mirror):
    """tests/test_all_step_multi_corridor.py:240-241: AllStepManager(sim, randomize_action_input=0) asserts (before any
    device work)."""
    import pytest
    from abmarl_b200.managers import AllStepManager, TurnBasedManager
    from tests import scenarios
    with pytest.raises(AssertionError):
        AllStepManager(scenarios.build_tb_c2(mirror), randomize_action_input=0)
    with pytest.raises(AssertionError):
        TurnBasedManager(scenarios.build_mm_tiny(mirror), randomize_action_input=True)


def test_randomize_placement_order_must_be_a_boolean(mirror):
    """tests/sim/gridworld/test_state.py:970-985."""
    import pytest
    from abmarl_b200.sim.gridworld.state import PositionState
    from abmarl_b200.sim.gridworld.grid import Grid
    from abmarl_b200.sim.gridworld.agent import GridWorldAgent
    agents = {'a': GridWorldAgent(id='a', encoding=1)}
    state = PositionState(grid=Grid(2, 2), agents=agents, randomize_placement_order=False)
    assert not state.randomize_placement_order
    with pytest.raises(AssertionError):
        PositionState(grid=Grid(2, 2), agents=agents, randomize_placement_order=1)
