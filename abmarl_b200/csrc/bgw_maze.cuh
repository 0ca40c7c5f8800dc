/*
 * bgw_maze.cuh -- MazePlacementState on the device: the per-episode start layout of
 * abmarl/sim/gridworld/state.py:385-619 (a maze grown from the target with Prim's algorithm, utils.py:120-212; barrier
 * encodings get the wall cells, free encodings the passage cells; optional clustering / scattering by distance from
 * the target).  Host and device compile the same code: bgw_maze_layout_host() is the host-callable form the tests
 * compare with abmarl_b200/layouts.py (which is pinned to the unmodified reference through the golden transcripts).
 *
 * The one delicate point: the reference rebuilds its frontier with `list(set(walls + new_walls))` (utils.py:203), so the
 * order of the frontier -- and with it which wall the next keyed draw picks -- is the iteration order of a CPython set
 * of (row, col) tuples.  PySet below restates CPython's setobject.c (open addressing, LINEAR_PROBES = 9, perturbation
 * shift 5, growth to the first power of two above 4 * used once fill * 5 >= mask * 3, re-insertion in table order) and
 * tupleobject.c's xxHash-style tuple hash for two small non-negative ints (Python >= 3.8).
 */
#pragma once
#include <stdint.h>

#include "../../include/bgw.h"
#include "../../include/bgw_philox.h"

#define BGW_MAZE_MAX_PADDED 400     /* (rows + 2) * (cols + 2) */
#define BGW_MAZE_MAX_CELLS 324      /* rows * cols */
#define BGW_MAZE_MAX_LISTS 8        /* encodings named by barrier_encodings | free_encodings */
#define BGW_MAZE_TABLE 2048         /* largest set table: first power of two above 4 * BGW_MAZE_MAX_PADDED */

struct MazeParams {                 /* what the generator needs of the spec (host or device pointers alike) */
    int kind;                       /* BGW_LAYOUT_MAZE or BGW_LAYOUT_TARGET_BARRIERS_FREE (no maze: every cell is barrier and free) */
    int rows, cols, A, max_enc, no_overlap, target, cluster_barriers, scatter_free;
    unsigned long long seed, barrier_encodings, free_encodings;
    const int8_t *enc;
    const int16_t *init_row, *init_col;
    const unsigned long long *overlap;
};

/* hash((a, b)) of CPython >= 3.8 for two ints in [0, 2^61): Objects/tupleobject.c tuplehash */
BGW_HD unsigned long long py_tuple2_hash(unsigned long long a, unsigned long long b)
{
    const unsigned long long P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P5 = 2870177450012600261ull;
    unsigned long long acc = P5;
    acc += a * P2; acc = (acc << 31) | (acc >> 33); acc *= P1;
    acc += b * P2; acc = (acc << 31) | (acc >> 33); acc *= P1;
    acc += 2ull ^ (P5 ^ 3527539ull);
    return acc == ~0ull ? 1546275796ull : acc;
}

/* Objects/setobject.c restated for keys = padded cell ids (r * pc + c), hashed as the tuple (r, c); the key array
 * holds id + 1, 0 = unused (nothing is ever deleted from these sets) */
struct PySet {
    uint16_t *key;                  /* [BGW_MAZE_TABLE] */
    uint16_t *tmp;                  /* [BGW_MAZE_TABLE] scratch of a resize */
    int mask, fill, used, pc;
};

BGW_HD unsigned long long pyset_hash(const PySet &s, int id) { return py_tuple2_hash((unsigned long long)(id / s.pc), (unsigned long long)(id % s.pc)); }

BGW_HD void pyset_clear(PySet &s)
{
    s.mask = 7; s.fill = s.used = 0;
    for (int i = 0; i < 8; ++i) s.key[i] = 0;
}

BGW_HD void pyset_insert_clean(uint16_t *table, int mask, int id, unsigned long long h)   /* set_insert_clean */
{
    unsigned long long perturb = h;
    size_t i = (size_t)h & (size_t)mask;
    for (;;) {
        if (table[i] == 0) break;
        bool found = false;
        if (i + 9 <= (size_t)mask)
            for (int j = 1; j <= 9; ++j) if (table[i + j] == 0) { i += j; found = true; break; }
        if (found) break;
        perturb >>= 5;
        i = (i * 5 + 1 + (size_t)perturb) & (size_t)mask;
    }
    table[i] = (uint16_t)(id + 1);
}

BGW_HD void pyset_resize(PySet &s, int minused)             /* set_table_resize */
{
    int newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    for (int i = 0; i < newsize; ++i) s.tmp[i] = 0;
    for (int i = 0; i <= s.mask; ++i)
        if (s.key[i]) pyset_insert_clean(s.tmp, newsize - 1, s.key[i] - 1, pyset_hash(s, s.key[i] - 1));
    for (int i = 0; i < newsize; ++i) s.key[i] = s.tmp[i];
    s.mask = newsize - 1;
    s.fill = s.used;
}

BGW_HD void pyset_add(PySet &s, int id)                     /* set_add_entry */
{
    const unsigned long long h = pyset_hash(s, id);
    unsigned long long perturb = h;
    size_t i = (size_t)h & (size_t)s.mask;
    for (;;) {
        const int probes = (i + 9 <= (size_t)s.mask) ? 9 : 0;
        for (int j = 0; j <= probes; ++j) {
            if (s.key[i + j] == 0) {
                s.key[i + j] = (uint16_t)(id + 1);
                ++s.fill; ++s.used;
                if (s.fill * 5 >= s.mask * 3) pyset_resize(s, s.used > 50000 ? s.used * 2 : s.used * 4);
                return;
            }
            if (s.key[i + j] == (uint16_t)(id + 1)) return;
        }
        perturb >>= 5;
        i = (i * 5 + 1 + (size_t)perturb) & (size_t)s.mask;
    }
}

struct MazeStream {                 /* the episode's maze draws in call order: site MAZE, slot 0, k = position */
    unsigned long long seed;
    uint32_t env, episode, k;
};

BGW_HD int maze_randint(MazeStream &st, int low, int high)
{
    const uint32_t x = bgw_draw(st.seed, st.env, st.episode, 0, BGW_SITE_MAZE, 0, st.k++);
    return low + (int)bgw_index(x, (uint32_t)(high - low));
}

/* utils.py:133-153: the unvisited neighbours of `cell` become walls and are appended to out[n..] */
BGW_HD int maze_unvisited(uint8_t *grid, int pr, int pc, int cell, uint16_t *out, int n)
{
    const int r = cell / pc, c = cell % pc;
    const int nb[4][2] = {{r - 1, c}, {r + 1, c}, {r, c - 1}, {r, c + 1}};
    for (int q = 0; q < 4; ++q) {
        const int rr = nb[q][0], cc = nb[q][1];
        if (rr == 0 || rr == pr - 1 || cc == 0 || cc == pc - 1) continue;       /* not along the borders */
        if (grid[rr * pc + cc] == 2) { out[n++] = (uint16_t)(rr * pc + cc); grid[rr * pc + cc] = 1; }
    }
    return n;
}

/* generate_maze utils.py:120-212 on the padded grid (0 passage, 1 wall, 2 unvisited -> wall); `start` = padded cell */
BGW_HD void maze_generate(uint8_t *grid, int pr, int pc, int start, MazeStream &st, uint16_t *walls, PySet &set)
{
    for (int i = 0; i < pr * pc; ++i) grid[i] = 2;
    grid[start] = 0;
    int nw = maze_unvisited(grid, pr, pc, start, walls, 0);
    set.pc = pc;
    while (nw > 0) {                                            /* :196-208 */
        const int pick = maze_randint(st, 0, nw), cur = walls[pick];
        const bool up = grid[cur - pc] == 2, down = grid[cur + pc] == 2, left = grid[cur - 1] == 2, right = grid[cur + 1] == 2;
        int at = pick;
        if ((up != down) || (left != right)) {
            const int nfree = (grid[cur - pc] == 0) + (grid[cur + pc] == 0) + (grid[cur - 1] == 0) + (grid[cur + 1] == 0);
            if (nfree < 2) {
                grid[cur] = 0;
                nw = maze_unvisited(grid, pr, pc, cur, walls, nw);
                pyset_clear(set);                               /* list(set(walls + new)) :203 */
                for (int i = 0; i < nw; ++i) pyset_add(set, walls[i]);
                nw = 0;
                for (int i = 0; i <= set.mask; ++i) if (set.key[i]) walls[nw++] = (uint16_t)(set.key[i] - 1);
                at = 0;
                while (walls[at] != cur) ++at;
            }
        }
        for (int i = at; i + 1 < nw; ++i) walls[i] = walls[i + 1];   /* unvisited_walls.remove(current_cell) */
        --nw;
    }
}

/* scratch of one layout (shared memory of one CTA on the device, heap on the host).  Two size classes: the small one
 * (grids up to 10x10 = BASELINE config 4, at most 4 placed encodings) lets 32 CTAs share an SM. */
template <int PADDED, int CELLS, int TABLE, int LISTS>
struct MazeScratchT {
    static constexpr int kPadded = PADDED, kCells = CELLS, kTable = TABLE, kLists = LISTS;
    uint8_t grid[PADDED];
    uint16_t walls[PADDED];
    uint16_t tab_a[TABLE], tab_b[TABLE];                        /* TABLE = first power of two above 4 * PADDED */
    uint16_t list[LISTS][CELLS];                                /* ravelled_positions_available per encoding */
    int list_n[LISTS];
    int8_t list_of[BGW_MAX_ENCODING + 1];                       /* encoding -> list index or -1 */
};
typedef MazeScratchT<BGW_MAZE_MAX_PADDED, BGW_MAZE_MAX_CELLS, BGW_MAZE_TABLE, BGW_MAZE_MAX_LISTS> MazeScratch;
typedef MazeScratchT<144, 100, 1024, 4> MazeScratchSmall;

/* a.sort(key=distance from the start, reverse=descending): Python's sort is stable, also with reverse=True */
BGW_HD void maze_sort(uint16_t *a, int n, int cols, int sr, int sc, bool descending)
{
    for (int i = 1; i < n; ++i) {
        const uint16_t v = a[i];
        const int dr = v / cols - sr, dc = v % cols - sc, kv = dr * dr + dc * dc;
        int j = i - 1;
        while (j >= 0) {
            const int er = a[j] / cols - sr, ec = a[j] % cols - sc, kj = er * er + ec * ec;
            if (descending ? kj < kv : kj > kv) { a[j + 1] = a[j]; --j; } else break;
        }
        a[j + 1] = v;
    }
}

/* MazePlacementState.reset state.py:487-619 for (global env, episode) -> layout[A] (BGW_NONE = not placed).
 * Returns 0, or 2 when an entity finds no cell (RuntimeError state.py:598-603). */
template <typename Scratch>
BGW_HD int maze_layout(const MazeParams &p, uint32_t genv, uint32_t episode, Scratch &w, uint16_t *layout)
{
    const int rows = p.rows, cols = p.cols, HW = rows * cols, pr = rows + 2, pc = cols + 2;
    MazeStream st{p.seed, genv, episode, 0};
    int sr, sc;
    if (p.init_row[p.target] >= 0) { sr = p.init_row[p.target]; sc = p.init_col[p.target]; }          /* :531-534 */
    else { sr = maze_randint(st, 0, rows); sc = maze_randint(st, 0, cols); }
    const bool maze = p.kind == BGW_LAYOUT_MAZE;
    if (maze) {
        PySet set;
        set.key = w.tab_a; set.tmp = w.tab_b;
        maze_generate(w.grid, pr, pc, (sr + 1) * pc + (sc + 1), st, w.walls, set);
    }
    /* barrier / free cell lists in ascending cell order, then the optional sorts :546-569 */
    int nlists = 0;
    for (int e = 0; e <= BGW_MAX_ENCODING; ++e) w.list_of[e] = -1;
    for (int e = 1; e <= p.max_enc && e <= BGW_MAX_ENCODING; ++e) {
        const bool is_free = (p.free_encodings >> e) & 1ull, is_barrier = (p.barrier_encodings >> e) & 1ull;
        if (!is_free && !is_barrier) continue;
        if (nlists == Scratch::kLists) return 3;
        uint16_t *lst = w.list[nlists];
        int n = 0;
        for (int cell = 0; cell < HW; ++cell) {
            const int g = maze ? w.grid[(cell / cols + 1) * pc + (cell % cols + 1)] : -1;   /* -1: state.py:311-339, all cells */
            if (g < 0 || (is_free ? g == 0 : g != 0)) lst[n++] = (uint16_t)cell;   /* a key in both sets ends up free (dict.update) */
        }
        if (is_free ? p.scatter_free : p.cluster_barriers) maze_sort(lst, n, cols, sr, sc, !is_free);
        w.list_n[nlists] = n;
        w.list_of[e] = (int8_t)nlists++;
    }
    for (int a = 0; a < p.A; ++a) layout[a] = BGW_NONE;
    int err = 0;
    /* place(): layout + _update_available_positions state.py:126-141 */
    auto place = [&](int a, int cell) {
        layout[a] = (uint16_t)cell;
        const unsigned long long row = p.overlap[p.enc[a]];
        for (int e = 1; e <= p.max_enc; ++e) {
            const int li = w.list_of[e];
            if (li < 0 || !(p.no_overlap || !((row >> e) & 1ull))) continue;
            uint16_t *lst = w.list[li];
            int n = w.list_n[li];
            for (int i = 0; i < n; ++i)
                if (lst[i] == cell) { for (int j = i; j + 1 < n; ++j) lst[j] = lst[j + 1]; w.list_n[li] = n - 1; break; }
        }
    };
    place(p.target, sr * cols + sc);                                                            /* :503-505 */
    for (int a = 0; a < p.A; ++a)                                                               /* :507-512 */
        if (a != p.target && p.init_row[a] >= 0) place(a, p.init_row[a] * cols + p.init_col[a]);
    for (int a = 0; a < p.A; ++a) {                                                             /* :514-519, 586-619 */
        if (a == p.target || p.init_row[a] >= 0) continue;
        const int e = p.enc[a], li = w.list_of[e];
        if (li < 0 || w.list_n[li] == 0) { err = 2; continue; }
        const bool is_free = (p.free_encodings >> e) & 1ull, is_barrier = (p.barrier_encodings >> e) & 1ull;
        int cell;
        if ((is_barrier && p.cluster_barriers) || (is_free && p.scatter_free)) cell = w.list[li][w.list_n[li] - 1];   /* :588-590 */
        else cell = w.list[li][bgw_index(bgw_draw(p.seed, genv, episode, 0, BGW_SITE_PLACE, (uint32_t)a, 0), (uint32_t)w.list_n[li])];
        place(a, cell);
    }
    return err;
}

template <typename Scratch>
BGW_HD bool maze_fits(int rows, int cols, int max_enc, unsigned long long barrier, unsigned long long freee)
{
    int lists = 0;
    for (int e = 1; e <= max_enc; ++e) lists += (int)(((barrier | freee) >> e) & 1ull);
    return rows * cols <= Scratch::kCells && (rows + 2) * (cols + 2) <= Scratch::kPadded && lists <= Scratch::kLists;
}

BGW_HD bool maze_supported(int rows, int cols, int max_enc, unsigned long long barrier, unsigned long long freee)
{
    return maze_fits<MazeScratch>(rows, cols, max_enc, barrier, freee);
}
