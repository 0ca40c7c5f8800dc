"""MazePlacementState on the device path (abmarl_b200/csrc/bgw_maze.cuh): the native generator -- CPython's set iteration
order restated, Prim's maze, clustered / scattered / random placement -- against abmarl_b200/layouts.py, the Python
restatement that the golden transcripts pin to the unmodified reference (state.py:385-619, utils.py:120-212)."""
import ctypes as C

import numpy as np
import pytest

from abmarl_b200 import _capi as K
from abmarl_b200.layouts import maze_layout
from abmarl_b200.spec import compile_sim
from tests import scenarios


def _host_layout(lib, cs, A, env, episode):
    out = np.zeros(A, dtype=np.uint16)
    rc = lib.bgw_maze_layout_host(C.byref(cs), env, episode, out.ctypes.data_as(C.c_void_p))
    assert rc == 0, lib.bgw_last_error()
    return out


@pytest.mark.parametrize('name', ['mm_c4', 'mm_random', 'mm_tiny', 'mm_tbf', 'mm_tbf_scatter'])
def test_native_maze_layouts_equal_the_python_restatement(mirror, name):
    from abmarl_b200.csrc.build import build
    build()
    lib = K.load()
    builder, manager, _ = scenarios.SCENARIOS[name]
    spec = compile_sim(builder(mirror), manager=manager, n_envs=4, env_offset=11, seed=0xC0FFEE)
    cs = spec.c_struct()
    assert cs.layout_kind == (K.LAYOUT_TARGET_BARRIERS_FREE if name.startswith('mm_tbf') else K.LAYOUT_MAZE)
    for env in (0, 1, 7, 4095, 70000):
        for episode in (0, 1, 2, 17, 2**32 - 1):
            want = maze_layout(spec, env, episode)
            got = _host_layout(lib, cs, spec.n_agents, env, episode)
            np.testing.assert_array_equal(got, want, err_msg=f'{name} env {env} episode {episode}')


def test_native_maze_layouts_many_shapes(mirror):
    """Other grid shapes (up to the generator's 18x18 limit), random target, no clustering / scattering."""
    from abmarl_b200.csrc.build import build
    build()
    lib = K.load()
    api = mirror
    rng = np.random.default_rng(3)
    for rows, cols, n_bar, n_nav, cluster, scatter in ((3, 3, 2, 1, True, True), (5, 9, 10, 3, False, True), (12, 7, 25, 4, True, False),
                                                      (18, 18, 60, 6, False, False), (9, 4, 6, 2, True, True)):
        agents = {'target': api.agent.GridWorldAgent(id='target', encoding=1)}
        agents.update({f'barrier{i}': api.agent.GridWorldAgent(id=f'barrier{i}', encoding=2) for i in range(n_bar)})
        agents.update({f'navigator{i}': api.ex.MultiMazeNavigationAgent(id=f'navigator{i}', encoding=3, view_range=2) for i in range(n_nav)})
        sim = api.ex.MultiMazeNavigationSim.build_sim(
            rows, cols, agents=agents, overlapping={1: {3}, 3: {3}}, target_agent=agents['target'], barrier_encodings={2},
            free_encodings={1, 3}, cluster_barriers=cluster, scatter_free_agents=scatter)
        spec = compile_sim(sim, manager='all_step', n_envs=1, seed=int(rng.integers(0, 2**62)))
        cs = spec.c_struct()
        for _ in range(6):
            env, episode = int(rng.integers(0, 2**20)), int(rng.integers(0, 2**16))
            try:
                want = maze_layout(spec, env, episode)
            except RuntimeError:                                # more barriers than wall cells in this maze (state.py:598-603)
                out = np.zeros(spec.n_agents, dtype=np.uint16)
                assert lib.bgw_maze_layout_host(C.byref(cs), env, episode, out.ctypes.data_as(C.c_void_p)) != 0
                continue
            np.testing.assert_array_equal(_host_layout(lib, cs, spec.n_agents, env, episode), want,
                                          err_msg=f'{rows}x{cols} env {env} episode {episode}')
