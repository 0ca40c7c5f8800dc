"""Host-side start layouts for placement states that build a random structure per episode.

MazePlacementState (abmarl/sim/gridworld/state.py:385-619) partitions the grid with a maze grown from the target
(Prim's algorithm, utils.py:120-212), then hands barrier-encoded entities the wall cells and free-encoded entities
the passage cells.  The structure is small, irregular and drawn once per episode, so it is generated on the
host from the keyed Philox stream (site MAZE, k = the draw's position in the episode's maze sequence) and fed to
the engine through BgwState.layout ([E][A] start cells, consumed by the next reset of each env).  The replay shim
(oracle/refshim.py) returns the same draws to the unmodified reference, which is how the layouts are pinned.
"""
import numpy as np

from abmarl_b200 import _capi as K
from abmarl_b200 import philox


class _Stream:
    """The episode's maze draws, in call order."""

    def __init__(self, seed, env, episode):
        self.seed, self.env, self.episode, self.k = seed, env, episode & 0xFFFFFFFF, 0

    def randint(self, low, high):
        x = philox.draw(self.seed, self.env, self.episode, 0, K.SITE_MAZE, 0, self.k)
        self.k += 1
        return low + philox.index(x, high - low)


def generate_maze(rows, cols, start, stream):
    """utils.py:120-212: 0 = passage, 1 = wall; `start` is a cell of the un-bordered grid."""
    rows, cols = rows + 2, cols + 2                       # a border ring is added and removed again (:183-186)
    grid = [[2] * cols for _ in range(rows)]

    def unvisited_neighbours(cell):                       # :133-153: mark unvisited neighbours as walls
        out = []
        for nb in ((cell[0] - 1, cell[1]), (cell[0] + 1, cell[1]), (cell[0], cell[1] - 1), (cell[0], cell[1] + 1)):
            if nb[0] in (0, rows - 1) or nb[1] in (0, cols - 1):
                continue
            if grid[nb[0]][nb[1]] == 2:
                out.append(nb)
                grid[nb[0]][nb[1]] = 1
        return out

    def free_neighbours(cell):                            # :155-176
        return sum(grid[r][c] == 0 for r, c in ((cell[0] - 1, cell[1]), (cell[0] + 1, cell[1]),
                                                (cell[0], cell[1] - 1), (cell[0], cell[1] + 1)))

    start = (int(start[0]) + 1, int(start[1]) + 1)
    grid[start[0]][start[1]] = 0
    walls = unvisited_neighbours(start)
    while walls:                                          # :196-208
        cur = walls[stream.randint(0, len(walls))]
        if ((grid[cur[0] - 1][cur[1]] == 2) ^ (grid[cur[0] + 1][cur[1]] == 2)) or \
                ((grid[cur[0]][cur[1] - 1] == 2) ^ (grid[cur[0]][cur[1] + 1] == 2)):
            if free_neighbours(cur) < 2:
                grid[cur[0]][cur[1]] = 0
                # the reference de-duplicates through a set: iteration order of a set of int tuples is a pure
                # function of the hashes and the insertion sequence, both identical here
                walls = list(set(walls + unvisited_neighbours(cur)))
        walls.remove(cur)
    maze = np.array(grid)
    maze[maze == 2] = 1
    return maze[1:-1, 1:-1]


def maze_layout(spec, env, episode):
    """MazePlacementState.reset (state.py:487-619) / TargetBarriersFreePlacementState.reset (state.py:281-383: the same
    flow without the maze) for global env `env`, episode `episode` -> uint16[A] cells."""
    p = spec.layout_generator[1]
    rows, cols, A = spec.rows, spec.cols, spec.n_agents
    stream = _Stream(spec.seed, env, episode)
    target = p['target']
    if spec.init_row[target] >= 0:                        # :531-534
        start = (int(spec.init_row[target]), int(spec.init_col[target]))
    else:
        start = (stream.randint(0, rows), stream.randint(0, cols))
    if spec.layout_generator[0] == 'maze':
        maze = generate_maze(rows, cols, start, stream)
    else:                                                 # TargetBarriersFreePlacementState state.py:311-339: every cell is both
        maze = None

    def dist(n):
        return float(np.linalg.norm(np.array([np.unravel_index(n, (rows, cols))]) - np.array(start)))
    barrier = [int(n) for n in np.flatnonzero(maze.ravel() == 1)] if maze is not None else list(range(rows * cols))
    free = [int(n) for n in np.flatnonzero(maze.ravel() == 0)] if maze is not None else list(range(rows * cols))
    if p['cluster_barriers']:
        barrier.sort(key=dist, reverse=True)              # closest last (:546-553)
    if p['scatter_free_agents']:
        free.sort(key=dist)                               # furthest last (:562-569)
    avail = {e: list(barrier) for e in p['barrier_encodings']}
    avail.update({e: list(free) for e in p['free_encodings']})
    overlap = {e: {f for f in range(1, K.BGW_MAX_ENCODING + 1) if (int(spec.overlap[e]) >> f) & 1}
               for e in range(1, K.BGW_MAX_ENCODING + 1)}

    layout = np.full(A, K.BGW_NONE, dtype=np.uint16)

    def place(a, cell):
        layout[a] = cell
        for e, lst in avail.items():                      # _update_available_positions state.py:126-141
            if spec.no_overlap_at_reset or e not in overlap[int(spec.encoding[a])]:
                if cell in lst:
                    lst.remove(cell)

    place(target, start[0] * cols + start[1])             # :503-505
    for a in range(A):                                    # :507-512
        if a != target and spec.init_row[a] >= 0:
            place(a, int(spec.init_row[a]) * cols + int(spec.init_col[a]))
    for a in range(A):                                    # :514-519, 586-619
        if a == target or spec.init_row[a] >= 0:
            continue
        e = int(spec.encoding[a])
        lst = avail[e]
        if not lst:
            raise RuntimeError(f"Could not find a cell for {spec.agent_ids[a]}")
        if (e in p['barrier_encodings'] and p['cluster_barriers']) or (e in p['free_encodings'] and p['scatter_free_agents']):
            cell = lst[-1]
        else:
            x = philox.draw(spec.seed, env, episode & 0xFFFFFFFF, 0, K.SITE_PLACE, a, 0)
            cell = lst[philox.index(x, len(lst))]
        place(a, cell)
    return layout


def layouts_for(spec, envs, episodes):
    """[len(envs), A] layouts; `envs` are LOCAL env indices (the global index adds spec.env_offset)."""
    kind = spec.layout_generator[0]
    assert kind in ('maze', 'target_barriers_free'), kind
    return np.stack([maze_layout(spec, spec.env_offset + int(e), int(ep)) for e, ep in zip(envs, episodes)])


class LayoutFeeder:
    """Keeps the [E, A] layout array that the NEXT reset of each env will consume.

    prime(episode)   before a reset: rows of the selected envs <- layout of episode[e] + 1
    after_step(...)  after a step: envs that reported __all__ (and will be reset by the next call when the engine
                     auto-resets) get the layout of their next episode
    """

    def __init__(self, spec):
        self.spec = spec
        self.rows = np.full((spec.n_envs, spec.n_agents), K.BGW_NONE, dtype=np.uint16)

    def prime(self, episode, env_mask=None):
        envs = [e for e in range(self.spec.n_envs) if env_mask is None or env_mask[e]]
        if envs:
            nxt = [(int(episode[e]) + 1) & 0xFFFFFFFF for e in envs]
            self.rows[envs] = layouts_for(self.spec, envs, nxt)
        return self.rows

    def after_step(self, env_flags, episode):
        envs = [e for e in range(self.spec.n_envs) if env_flags[e] & K.ENV_ALL_DONE]
        if envs:
            self.rows[envs] = layouts_for(self.spec, envs, [(int(episode[e]) + 1) & 0xFFFFFFFF for e in envs])
        return bool(envs)
