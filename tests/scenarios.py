"""Simulation definitions used by the parity tests and the golden generator.

Every `build_*` function takes an `api` namespace and builds the sim from it.  `api` is either this
package (`mirror_api()`) or the unmodified reference (`reference_api()`, build container only), so the
SAME definition code produces the reference object the goldens were recorded from and the mirror object
the tests compile -- the compiled specs must be identical (checked in tests/test_golden_oracle.py).
"""
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LAYOUTS = os.path.join(HERE, 'golden', 'layouts')


def mirror_api():
    from abmarl_b200 import examples as ex, managers
    from abmarl_b200.sim.gridworld import agent, actor, observer, state, done, wrapper
    from abmarl_b200.examples import traffic_corridor
    return types.SimpleNamespace(
        name='mirror', ex=ex, agent=agent, actor=actor, observer=observer, state=state, done=done,
        wrapper=wrapper, managers=managers, pacman=ex, traffic=traffic_corridor)


def reference_api():
    from oracle.refshim import load_reference
    load_reference()
    import abmarl.examples as ex
    import abmarl.managers as managers
    from abmarl.examples.sim import pacman
    from abmarl.sim.gridworld import agent, actor, observer, state, done, wrapper
    from abmarl.examples.sim import traffic_corridor
    return types.SimpleNamespace(
        name='reference', ex=ex, agent=agent, actor=actor, observer=observer, state=state, done=done,
        wrapper=wrapper, managers=managers, pacman=pacman, traffic=traffic_corridor)


# ---------------------------------------------------------------------------------------------------
# team battle family (team_battle_example.py; examples/rllib_team_battle.py:9-43)
# ---------------------------------------------------------------------------------------------------
def _team_maps(n_teams):
    overlap = {k: {k} for k in range(1, n_teams + 1)}
    attack = {k: {j for j in range(1, n_teams + 1) if j != k} for k in range(1, n_teams + 1)}
    return overlap, attack


def build_tb_c2(api):
    """BASELINE config 2: 4 teams on 8x8, 24 agents spawning on four corner cells, random health."""
    positions = [np.array([1, 1]), np.array([1, 6]), np.array([6, 1]), np.array([6, 6])]
    agents = {
        f'agent{i}': api.ex.BattleAgent(id=f'agent{i}', encoding=i % 4 + 1, initial_position=positions[i % 4])
        for i in range(24)
    }
    overlap, attack = _team_maps(4)
    return api.ex.TeamBattleSim.build_sim(
        8, 8, agents=agents, overlapping=overlap, attack_mapping=attack,
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def build_tb_c5(api, rows=64, cols=64, n_agents=256, view_range=5, initial_health=None):
    """BASELINE config 5: synthetic team battle, random placement and health."""
    agents = {}
    for i in range(n_agents):
        ag = api.ex.BattleAgent(id=f'agent{i}', encoding=i % 4 + 1, initial_health=initial_health)
        ag.view_range = view_range
        agents[ag.id] = ag
    overlap, attack = _team_maps(4)
    return api.ex.TeamBattleSim.build_sim(
        rows, cols, agents=agents, overlapping=overlap, attack_mapping=attack,
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def build_tb_c5_small(api):
    return build_tb_c5(api, rows=24, cols=24, n_agents=96, view_range=5)


def build_tb_dense(api):
    """Crowded 6x7 board, 3 teams, teams 1 and 2 may share cells (mixed-encoding cells => observer draws),
    imperfect accuracy, strength < 1 (several hits to kill), heterogeneous ranges."""
    agents = {}
    for i in range(20):
        ag = api.ex.BattleAgent(id=f'a{i}', encoding=i % 3 + 1, initial_health=1.0 if i % 2 else None)
        ag.view_range = 2 + i % 3
        ag.attack_range = 1 + i % 2
        ag.move_range = 1 + (i % 5 == 0)
        ag.attack_strength = 0.6 if i % 4 else 1
        ag.attack_accuracy = 0.75 if i % 3 else 1
        agents[ag.id] = ag
    return api.ex.TeamBattleSim.build_sim(
        6, 7, agents=agents, overlapping={1: {1, 2}, 2: {2}, 3: {3}},
        attack_mapping={1: {2, 3}, 2: {1, 3}, 3: {1, 2}},
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def build_tb_blocking(api):
    """Team battle among view- and attack-blocking pillars (encoding 5, never attackable), plus two
    blocking fighters: exercises create_grid_and_mask in both the attack and the observation path."""
    agents = {}
    pillars = [(2, 2), (2, 7), (4, 4), (4, 5), (5, 4), (7, 2), (7, 7), (0, 5), (9, 4), (5, 9)]
    for n, (r, c) in enumerate(pillars):
        agents[f'pillar{n}'] = api.agent.GridWorldAgent(
            id=f'pillar{n}', encoding=5, blocking=True, initial_position=np.array([r, c]))
    for i in range(18):
        ag = api.ex.BattleAgent(id=f'f{i}', encoding=i % 4 + 1, initial_health=1.0, blocking=(i in (3, 8)))
        ag.view_range = 4
        ag.attack_range = 2
        agents[ag.id] = ag
    overlap, attack = _team_maps(4)
    return api.ex.TeamBattleSim.build_sim(
        10, 10, agents=agents, overlapping=overlap, attack_mapping=attack,
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def build_tb_stacked(api):
    """Stacked (per-encoding count) observer + observe-everything overlap."""
    agents = {}
    for i in range(14):
        ag = api.ex.BattleAgent(id=f'a{i}', encoding=i % 3 + 1, initial_health=1.0)
        ag.view_range = 2
        agents[ag.id] = ag
    return api.ex.TeamBattleSim.build_sim(
        5, 6, agents=agents, overlapping={1: {1, 2, 3}, 2: {2, 3}, 3: {3}},
        attack_mapping={1: {2}, 2: {3}, 3: {1}},
        states={'PositionState', 'HealthState'}, observers={'StackedPositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def build_tb_shuffled(api):
    """Keyed shuffles (Python's random.shuffle in the reference): PositionState(randomize_placement_order=True)
    state.py:97-101 -- with fixed-position agents that share a cell (arrival order = shuffled order), a crowded grid so
    that the order of the random placements matters -- run under AllStepManager(randomize_action_input=True)
    all_step_manager.py:62-65 (scenario manager 'all_step_shuffled')."""
    agents = {}
    for i in range(22):
        fixed = {0: (1, 1), 3: (1, 1), 6: (1, 1), 1: (3, 3), 4: (3, 3), 9: (0, 4)}.get(i)
        ag = api.ex.BattleAgent(id=f'a{i}', encoding=i % 3 + 1, initial_health=None if i % 2 else 1.0,
                                initial_position=None if fixed is None else np.array(fixed))
        ag.view_range = 2
        ag.attack_strength = 0.6
        agents[ag.id] = ag
    return api.ex.TeamBattleSim.build_sim(
        5, 6, agents=agents, overlapping={1: {1}, 2: {2, 3}, 3: {3}},
        attack_mapping={1: {2, 3}, 2: {1, 3}, 3: {1, 2}}, randomize_placement_order=True,
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def build_tb_c5_placement_shuffled(api):
    """The reduced headline shape (the specialised kernel) with shuffled placement order only."""
    agents = {}
    for i in range(96):
        ag = api.ex.BattleAgent(id=f'agent{i}', encoding=i % 4 + 1, initial_health=None)
        ag.view_range = 5
        agents[ag.id] = ag
    overlap, attack = _team_maps(4)
    return api.ex.TeamBattleSim.build_sim(
        24, 24, agents=agents, overlapping=overlap, attack_mapping=attack, randomize_placement_order=True,
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def build_tb_position(api):
    """AbsolutePositionObserver (observer.py:337-373) next to the grid observer: every agent also observes its own
    (row, col); an agent that died keeps observing the cell it died in."""
    agents = {}
    for i in range(16):
        ag = api.ex.BattleAgent(id=f'a{i}', encoding=i % 2 + 1, initial_health=None if i % 3 else 0.7)
        ag.view_range = 2
        agents[ag.id] = ag
    return api.ex.TeamBattleSim.build_sim(
        6, 7, agents=agents, overlapping={1: {1}, 2: {2}}, attack_mapping={1: {2}, 2: {1}},
        states={'PositionState', 'HealthState'},
        observers={'PositionCenteredEncodingObserver', 'AbsolutePositionObserver'}, dones={'OneTeamRemainingDone'})


def build_tb_noself(api):
    """observe_self=False (observer.py:238-246) with overlapping teams."""
    agents = {}
    for i in range(12):
        ag = api.ex.BattleAgent(id=f'a{i}', encoding=i % 2 + 1, initial_health=1.0)
        ag.view_range = 2
        agents[ag.id] = ag
    return api.ex.TeamBattleSim.build_sim(
        5, 5, agents=agents, overlapping={1: {1, 2}, 2: {2}}, attack_mapping={1: {2}, 2: {1}},
        observe_self=False,
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


# ---------------------------------------------------------------------------------------------------
# team battle with the other attack actors / ammo (actor.py:504-728, agent.py:291-339)
# ---------------------------------------------------------------------------------------------------
def _tb_attack_variant(api, actor_name, stacked, n_agents=18, rows=6, cols=7, sim_attacks=2, ammo=None,
                       ammo_observer=False, blocking=()):
    """TeamBattleSim (its step() only needs process_action's (status, attacked_agents) contract) with the attack
    actor swapped for `actor_name`; three teams, teams 1 and 2 may share cells, mixed ranges / strength /
    accuracy so that several hits, misses and repeated evaluations of the same pair all occur."""
    base = api.ex.BattleAgent
    if ammo is not None:
        class AmmoBattleAgent(api.ex.BattleAgent, api.agent.AmmoAgent):
            pass
        base = AmmoBattleAgent
    agents = {}
    for i in range(n_agents):
        kw = dict(id=f'a{i}', encoding=i % 3 + 1, initial_health=1.0 if i % 2 else None, blocking=i in blocking)
        if ammo is not None:
            kw['initial_ammo'] = ammo + i % 3
        ag = base(**kw)
        ag.view_range = 2
        ag.attack_range = 1 + (i % 4 == 0)
        ag.simultaneous_attacks = sim_attacks + (i % 5 == 0 and actor_name != 'BinaryAttackActor')
        ag.attack_strength = 0.5 if i % 3 else 1
        ag.attack_accuracy = 0.7 if i % 2 else 1
        agents[ag.id] = ag
    attack = {1: {2, 3}, 2: {1, 3}, 3: {1, 2}}
    states = {'PositionState', 'HealthState'} | ({'AmmoState'} if ammo is not None else set())
    observers = {'PositionCenteredEncodingObserver'} | ({'AmmoObserver'} if ammo_observer else set())
    sim = api.ex.TeamBattleSim.build_sim(
        rows, cols, agents=agents, overlapping={1: {1, 2}, 2: {2}, 3: {3}}, attack_mapping=attack,
        states=states, observers=observers, dones={'OneTeamRemainingDone'})
    if actor_name != 'BinaryAttackActor' or stacked:
        sim.attack_actor = getattr(api.actor, actor_name)(
            agents=sim.agents, grid=sim.grid, attack_mapping=attack, stacked_attacks=stacked)
    return sim


def build_tb_encoding(api):
    return _tb_attack_variant(api, 'EncodingBasedAttackActor', False)


def build_tb_encoding_stacked(api):
    return _tb_attack_variant(api, 'EncodingBasedAttackActor', True)


def build_tb_restricted(api):
    return _tb_attack_variant(api, 'RestrictedSelectiveAttackActor', False, sim_attacks=3, rows=4, cols=5)


def build_tb_restricted_stacked(api):
    return _tb_attack_variant(api, 'RestrictedSelectiveAttackActor', True, sim_attacks=3, rows=4, cols=5, blocking=(2, 8))


def build_tb_selective(api):
    return _tb_attack_variant(api, 'SelectiveAttackActor', False, blocking=(0, 4, 7))


def build_tb_selective_stacked(api):
    return _tb_attack_variant(api, 'SelectiveAttackActor', True)


def build_tb_ammo(api):
    """BinaryAttackActor + AmmoAgent / AmmoState / AmmoObserver: attacks stop when the ammo is spent."""
    return _tb_attack_variant(api, 'BinaryAttackActor', False, sim_attacks=1, ammo=2, ammo_observer=True, n_agents=15)


def build_tb_ammo_selective(api):
    """SelectiveAttackActor naming more agents than the attacker has ammo left: the ammo filter's choice."""
    return _tb_attack_variant(api, 'SelectiveAttackActor', True, ammo=3, ammo_observer=True)


# ---------------------------------------------------------------------------------------------------
# reach the target (reach_the_target.py:90-176; examples/rllib_reach_the_target.py:7-56)
# ---------------------------------------------------------------------------------------------------
def build_reach_target(api, grid_size=7, n_barriers=10, n_runners=4, corners=True, move_range=2, sim_attacks=1):
    """examples/rllib_reach_the_target.py: ten view-blocking barriers at random cells, four runners starting in the
    corners, the target in the middle shooting with the SelectiveAttackActor."""
    ex = api.ex
    corner = [np.array([0, 0]), np.array([grid_size - 1, 0]), np.array([0, grid_size - 1]), np.array([grid_size - 1, grid_size - 1])]
    agents = {f'barrier{i}': ex.BarrierAgent(id=f'barrier{i}') for i in range(n_barriers)}
    for i in range(n_runners):
        agents[f'runner{i}'] = ex.RunningAgent(
            id=f'runner{i}', move_range=move_range, view_range=grid_size // 2, initial_health=1,
            initial_position=corner[i] if corners and i < 4 else None)
    agents['target'] = ex.TargetAgent(
        view_range=grid_size, attack_range=1, attack_strength=1, attack_accuracy=1, simultaneous_attacks=sim_attacks,
        initial_position=np.array([grid_size // 2, grid_size // 2]))
    return ex.ReachTheTargetSim.build_sim(
        grid_size, grid_size, agents=agents, overlapping={2: {3}, 3: {1, 2, 3}}, attack_mapping={2: {3}})


def build_reach_target_crowd(api):
    """A busier variant: more runners placed at random (some may start on the target's cell), weaker and less
    accurate shots from a target that may shoot twice per cell, move range 1."""
    sim = build_reach_target(api, grid_size=6, n_barriers=6, n_runners=9, corners=False, move_range=1, sim_attacks=2)
    sim.target.attack_strength = 0.6
    sim.target.attack_accuracy = 0.8
    sim.target.attack_range = 2
    sim.attack_actor = api.actor.SelectiveAttackActor(agents=sim.agents, grid=sim.grid, attack_mapping={2: {3}})
    return sim


# ---------------------------------------------------------------------------------------------------
# traffic corridor (traffic_corridor.py:24-53; examples/rllib_traffic_corridor_2_teams.py:7-66)
# ---------------------------------------------------------------------------------------------------
def build_traffic(api):
    """examples/rllib_traffic_corridor_2_teams.py: two teams cross a one-cell-wide corridor in opposite directions."""
    tc = api.traffic
    grid = np.array([['G', 'W', 'W', 'W', 'R'],
                     ['r', '_', '_', '_', 'g'],
                     ['G', 'W', 'W', 'W', 'R']])
    registry = {
        'R': lambda n: tc.TrafficAgent(id=f'red{n}', encoding=1),
        'G': lambda n: tc.TrafficAgent(id=f'green{n}', encoding=2),
        'r': lambda n: tc.TargetAgent(id='red_target', encoding=1),
        'g': lambda n: tc.TargetAgent(id='green_target', encoding=2),
        'W': lambda n: tc.WallAgent(id=f'wall{n}', encoding=3),
    }
    return tc.TrafficCorridorSimulation.build_sim_from_array(
        grid, registry, overlapping={1: {1}, 2: {2}}, states={'PositionState'}, dones={'TargetAgentDone'},
        observers={'PositionCenteredEncodingObserver'},
        target_mapping={'red4': 'red_target', 'red11': 'red_target', 'green0': 'green_target', 'green7': 'green_target'})


# ---------------------------------------------------------------------------------------------------
# maze (maze_navigation.py; examples/rllib_maze_navigation.py:7-38)
# ---------------------------------------------------------------------------------------------------
def build_maze_c1(api):
    """BASELINE config 1: maze.txt, one navigator (view 2), 65 view-blocking walls, one target."""
    registry = {
        'N': lambda n: api.ex.MazeNavigationAgent(id='navigator', encoding=1, view_range=2),
        'T': lambda n: api.agent.GridWorldAgent(id='target', encoding=3),
        'W': lambda n: api.agent.GridWorldAgent(id=f'wall{n}', encoding=2, blocking=True),
    }
    return api.ex.MazeNavigationSim.build_sim_from_file(
        os.path.join(LAYOUTS, 'maze.txt'), registry, overlapping={1: {3}, 3: {1}},
        states={'PositionState'}, observers={'PositionCenteredEncodingObserver'})


# ---------------------------------------------------------------------------------------------------
# pacman (pacman.py:29-151; examples/rllib_pacman.py with blocking walls)
# ---------------------------------------------------------------------------------------------------
def build_pacman_c3(api, view_range=20):
    """BASELINE config 3: pacman.txt, multi-agent PacmanSim, view-blocking walls, absolute observer.
    view_range 20 covers the 21x21 board exactly like the class default of 100."""
    px = api.pacman

    def ranged(agent):
        agent.view_range = view_range
        return agent

    registry = {
        'P': lambda n: ranged(px.PacmanAgent(id='pacman', encoding=1)),
        'W': lambda n: px.WallAgent(id=f'wall_{n}', encoding=2, blocking=True),
        'F': lambda n: px.FoodAgent(id=f'food_{n}', encoding=3),
        'B': lambda n: ranged(px.BaddieAgent(id=f'baddie_{n}', encoding=4)),
    }
    return px.PacmanSim.build_sim_from_file(
        os.path.join(LAYOUTS, 'pacman.txt'), registry,
        states={'PositionState', 'OrientationState', 'HealthState'}, observers={'AbsoluteEncodingObserver'},
        overlapping={1: {3, 4}, 4: {3, 4}},
        reward_scheme={'bad_move': 0, 'entropy': -0.01, 'eat_food': 0.05, 'kill': 1, 'die': -1})


def build_pacman_simple(api, blocking=False):
    """examples/rllib_pacman.py as shipped: PacmanSimSimple (scripted baddies) on the example grid, walls that do NOT block
    (`blocking=True` is commented out there), default view range (the whole board), the script's reward scheme."""
    px = api.pacman
    registry = {
        'P': lambda n: px.PacmanAgent(id='pacman', encoding=1),
        'W': lambda n: px.WallAgent(id=f'wall_{n}', encoding=2, blocking=blocking),
        'F': lambda n: px.FoodAgent(id=f'food_{n}', encoding=3),
        'B': lambda n: px.BaddieAgent(id=f'baddie_{n}', encoding=4),
    }
    return px.PacmanSimSimple.build_sim_from_file(
        os.path.join(LAYOUTS, 'pacman.txt'), registry,
        states={'PositionState', 'OrientationState', 'HealthState'}, observers={'AbsoluteEncodingObserver'},
        overlapping={1: {3, 4}, 4: {3, 4}},
        reward_scheme={'bad_move': 0, 'entropy': -0.01, 'eat_food': 0.05, 'die': -1})


# ---------------------------------------------------------------------------------------------------
# multi maze (multi_maze_navigation.py; examples/rllib_multi_maze_navigation.py:7-41)
# ---------------------------------------------------------------------------------------------------
def build_mm_c4(api, ravel=True):
    """BASELINE config 4: 10x10 maze rebuilt every episode around the target (MazePlacementState), 20 clustered
    barriers, 5 scattered navigators with view 5; RavelActionWrapper on the move actor."""
    agents = {'target': api.agent.GridWorldAgent(id='target', encoding=1)}
    agents.update({f'barrier{i}': api.agent.GridWorldAgent(id=f'barrier{i}', encoding=2) for i in range(20)})
    agents.update({f'navigator{i}': api.ex.MultiMazeNavigationAgent(id=f'navigator{i}', encoding=3, view_range=5)
                   for i in range(5)})
    sim = api.ex.MultiMazeNavigationSim.build_sim(
        10, 10, agents=agents, overlapping={1: {3}, 3: {3}}, target_agent=agents['target'],
        barrier_encodings={2}, free_encodings={1, 3}, cluster_barriers=True, scatter_free_agents=True,
        no_overlap_at_reset=True)
    if ravel:
        sim.move_actor = api.wrapper.RavelActionWrapper(sim.move_actor)
    return sim


def build_mm_random(api):
    """Multi maze with randomly placed barriers and navigators (no clustering / scattering), move range 2, and a
    fixed target: exercises the PLACE draws of MazePlacementState and the ravel decode for a 5x5 action box."""
    agents = {'target': api.agent.GridWorldAgent(id='target', encoding=1, initial_position=np.array([4, 3]))}
    agents.update({f'barrier{i}': api.agent.GridWorldAgent(id=f'barrier{i}', encoding=2) for i in range(12)})
    for i in range(4):
        nav = api.ex.MultiMazeNavigationAgent(id=f'navigator{i}', encoding=3, view_range=3)
        nav.move_range = 2
        agents[nav.id] = nav
    sim = api.ex.MultiMazeNavigationSim.build_sim(
        8, 9, agents=agents, overlapping={1: {3}, 3: {3}}, target_agent=agents['target'],
        barrier_encodings={2}, free_encodings={1, 3})
    sim.move_actor = api.wrapper.RavelActionWrapper(api.actor.MoveActor(agents=sim.agents, grid=sim.grid))
    return sim


def build_mm_tiny(api):
    """A 4x5 multi maze: navigators reach the target within a few turns, so a short transcript covers several
    episodes -- done agents leaving the turn cycle, the 'just finished + next agent' double report and the
    cycle that is not rewound by reset (turn_based_manager.py:17-20,53-92)."""
    agents = {'target': api.agent.GridWorldAgent(id='target', encoding=1)}
    agents.update({f'barrier{i}': api.agent.GridWorldAgent(id=f'barrier{i}', encoding=2) for i in range(3)})
    agents.update({f'navigator{i}': api.ex.MultiMazeNavigationAgent(id=f'navigator{i}', encoding=3, view_range=2)
                   for i in range(3)})
    sim = api.ex.MultiMazeNavigationSim.build_sim(
        4, 5, agents=agents, overlapping={1: {3}, 3: {3}}, target_agent=agents['target'],
        barrier_encodings={2}, free_encodings={1, 3}, cluster_barriers=True, scatter_free_agents=False)
    sim.move_actor = api.wrapper.RavelActionWrapper(sim.move_actor)
    return sim


def dynamic_order_multi_maze_class(api):
    """DynamicOrderMultiMazeSim: the mirror's declaration, or -- for the reference -- its twin written out in Python on
    the reference's own classes (MultiMazeNavigationSim + DynamicOrderSimulation), which is what the goldens record."""
    if api.name == 'mirror':
        return api.ex.DynamicOrderMultiMazeSim
    from abmarl.sim import DynamicOrderSimulation

    class DynamicOrderMultiMazeSim(api.ex.MultiMazeNavigationSim, DynamicOrderSimulation):
        def reset(self, **kwargs):
            super().reset(**kwargs)
            self._navigators = [a.id for a in self.agents.values() if isinstance(a, api.ex.MultiMazeNavigationAgent)]
            self.next_agent = self._navigators[0]

        def step(self, action_dict, **kwargs):
            super().step(action_dict, **kwargs)
            last = list(action_dict)[-1]
            nxt = [agent_id for agent_id in action_dict if self.get_done(agent_id)]     # just finished: last report
            i = self._navigators.index(last)
            for k in range(1, len(self._navigators) + 1):
                cand = self._navigators[(i + k) % len(self._navigators)]
                if not self.get_done(cand):
                    nxt.append(cand)
                    break
            self.next_agent = nxt or [last]

        def get_reward(self, agent_id, **kwargs):
            # once the sim is done, DynamicOrderManager.step asks for EVERY agent that is not done (dynamic_order_manager.py:
            # 43-51), walls and target included (it does not park the non-learning entities in done_agents at reset as
            # AllStepManager does, all_step_manager.py:41-44); MultiMazeNavigationSim tracks rewards for Agents only
            if agent_id not in self.reward:
                return 0
            return super().get_reward(agent_id, **kwargs)
    return DynamicOrderMultiMazeSim


def build_mm_dynamic(api):
    """DynamicOrderManager (managers/dynamic_order_manager.py:7-87) over a 5x6 multi maze with four navigators."""
    agents = {'target': api.agent.GridWorldAgent(id='target', encoding=1)}
    agents.update({f'barrier{i}': api.agent.GridWorldAgent(id=f'barrier{i}', encoding=2) for i in range(4)})
    agents.update({f'navigator{i}': api.ex.MultiMazeNavigationAgent(id=f'navigator{i}', encoding=3, view_range=2)
                   for i in range(4)})
    sim = dynamic_order_multi_maze_class(api).build_sim(
        5, 6, agents=agents, overlapping={1: {3}, 3: {3}}, target_agent=agents['target'],
        barrier_encodings={2}, free_encodings={1, 3}, cluster_barriers=True, scatter_free_agents=False)
    sim.move_actor = api.wrapper.RavelActionWrapper(sim.move_actor)
    return sim


def build_mm_tbf(api, cluster=True, scatter=False, fixed_target=False):
    """Multi maze navigation with its placement state swapped for TargetBarriersFreePlacementState (state.py:169-383):
    the target at a random cell, barriers clustered around it, navigators placed at random."""
    agents = {'target': api.agent.GridWorldAgent(id='target', encoding=1,
                                                 initial_position=np.array([2, 5]) if fixed_target else None)}
    agents.update({f'barrier{i}': api.agent.GridWorldAgent(id=f'barrier{i}', encoding=2) for i in range(9)})
    agents.update({f'navigator{i}': api.ex.MultiMazeNavigationAgent(id=f'navigator{i}', encoding=3, view_range=3)
                   for i in range(4)})
    kw = dict(target_agent=agents['target'], barrier_encodings={2}, free_encodings={1, 3}, cluster_barriers=cluster,
              scatter_free_agents=scatter)
    sim = api.ex.MultiMazeNavigationSim.build_sim(7, 8, agents=agents, overlapping={1: {3}, 3: {3}}, **kw)
    kw['target_agent'] = sim.agents['target']
    sim.position_state = api.state.TargetBarriersFreePlacementState(agents=sim.agents, grid=sim.grid, **kw)
    return sim


def build_mm_tbf_scatter(api):
    return build_mm_tbf(api, cluster=False, scatter=True, fixed_target=True)


# the recording harness resets these scenarios every k steps even if the episode is not over (what a horizon does; the
# reference's managers can be reset at any time): an episode of the 64x64 battle outlasts any transcript worth committing
GOLDEN_RESET_EVERY = {'tb_c5': 60}

SCENARIOS = {
    # name: (builder, manager, steps recorded in the golden file)
    'tb_c2': (build_tb_c2, 'all_step', 400),
    'tb_c5': (build_tb_c5, 'all_step', 150),             # the true headline shape: 64x64, 256 agents, view 5
    'tb_c5_small': (build_tb_c5_small, 'all_step', 25),
    'tb_dense': (build_tb_dense, 'all_step', 40),
    'tb_blocking': (build_tb_blocking, 'all_step', 30),
    'tb_stacked': (build_tb_stacked, 'all_step', 25),
    'tb_noself': (build_tb_noself, 'all_step', 25),
    'tb_shuffled': (build_tb_shuffled, 'all_step_shuffled', 160),
    'tb_position': (build_tb_position, 'all_step', 60),
    'tb_c5_shuffled': (build_tb_c5_placement_shuffled, 'all_step_shuffled', 30),
    'tb_encoding': (build_tb_encoding, 'all_step', 40),
    'tb_encoding_stacked': (build_tb_encoding_stacked, 'all_step', 40),
    'tb_restricted': (build_tb_restricted, 'all_step', 40),
    'tb_restricted_stacked': (build_tb_restricted_stacked, 'all_step', 40),
    'tb_selective': (build_tb_selective, 'all_step', 40),
    'tb_selective_stacked': (build_tb_selective_stacked, 'all_step', 40),
    'tb_ammo': (build_tb_ammo, 'all_step', 40),
    'tb_ammo_selective': (build_tb_ammo_selective, 'all_step', 40),
    'reach_target': (build_reach_target, 'all_step', 60),
    'reach_target_crowd': (build_reach_target_crowd, 'all_step', 60),
    'traffic': (build_traffic, 'all_step', 120),
    'maze_c1': (build_maze_c1, 'all_step', 60),
    'pacman_c3': (build_pacman_c3, 'all_step', 64),
    'pacman_simple': (build_pacman_simple, 'all_step', 150),
    'mm_c4': (build_mm_c4, 'turn_based', 120),
    'mm_random': (build_mm_random, 'turn_based', 90),
    'mm_allstep': (build_mm_c4, 'all_step', 60),
    'mm_tbf': (build_mm_tbf, 'all_step', 80),
    'mm_tbf_scatter': (build_mm_tbf_scatter, 'turn_based', 120),
    'mm_tiny': (build_mm_tiny, 'turn_based', 400),
    'mm_tiny_allstep': (build_mm_tiny, 'all_step', 200),
    'mm_dynamic': (build_mm_dynamic, 'dynamic_order', 500),
}
