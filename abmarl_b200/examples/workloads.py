"""The synthetic headline workload (BASELINE.json configs[4], SURVEY.md 8(d) "C5"): a team battle the reference's own
example classes define (examples/sim/team_battle_example.py:11-59, built as examples/rllib_team_battle.py:9-43 builds its
smaller one), scaled to a 64x64 grid with 256 agents in 4 teams."""
from .sims import BattleAgent, TeamBattleSim


def team_maps(n_teams):
    """overlapping: a team shares cells with itself only; attack_mapping: every team attacks all the others."""
    overlap = {k: {k} for k in range(1, n_teams + 1)}
    attack = {k: {j for j in range(1, n_teams + 1) if j != k} for k in range(1, n_teams + 1)}
    return overlap, attack


def synthetic_team_battle(rows=64, cols=64, n_agents=256, view_range=5, initial_health=None, n_teams=4):
    """encoding = i % 4 + 1, view 5, move / attack range 1, strength and accuracy 1, random placement (PositionState),
    initial health U(0, 1) unless given, PositionCenteredEncodingObserver, OneTeamRemainingDone."""
    agents = {}
    for i in range(n_agents):
        ag = BattleAgent(id=f'agent{i}', encoding=i % n_teams + 1, initial_health=initial_health)
        ag.view_range = view_range
        agents[ag.id] = ag
    overlap, attack = team_maps(n_teams)
    return TeamBattleSim.build_sim(
        rows, cols, agents=agents, overlapping=overlap, attack_mapping=attack,
        states={'PositionState', 'HealthState'}, observers={'PositionCenteredEncodingObserver'},
        dones={'OneTeamRemainingDone'})


def headline_spec(n_envs, env_offset=0, seed=0xB200, horizon=200):
    """The compiled spec bench.py measures: AllStepManager semantics, horizon 200, auto-reset, keyed seed 0xB200."""
    from abmarl_b200.spec import compile_sim
    return compile_sim(synthetic_team_battle(), manager='all_step', n_envs=n_envs, env_offset=env_offset, seed=seed,
                       horizon=horizon, auto_reset=True)
