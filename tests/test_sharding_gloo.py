"""N > 1 host logic on CPU: two gloo ranks each advance their shard of the env batch (here with the oracle, since
there is no GPU) and sum the episode statistics; the result must equal one process advancing the whole batch,
because Philox is keyed by the global env index."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E, STEPS = 6, 25


def _spec(n_envs, offset):
    from tests import scenarios
    from abmarl_b200.spec import compile_sim
    sim = scenarios.build_tb_dense(scenarios.mirror_api())
    return compile_sim(sim, n_envs=n_envs, env_offset=offset, seed=77, horizon=20, auto_reset=True)


def _rollout(spec):
    from oracle.oracle import OracleEnv
    o = OracleEnv(spec)
    o.reset()
    obs_sum = 0
    for _ in range(STEPS):
        o.step(o.sample_actions())
        obs_sum += int(o.obs.astype(np.int64).sum())
    return o, obs_sum


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from abmarl_b200.dist import shard_envs, reduce_stats
    offset, n = shard_envs(E, world, rank)
    o, obs_sum = _rollout(_spec(n, offset))
    stats = torch.from_numpy(o.state['stats'].astype(np.int64).sum(axis=0))
    stats = torch.cat([stats, torch.tensor([obs_sum])])
    reduce_stats(stats)
    if rank == 0:
        np.save(out, stats.numpy())
    np.save(out + f'.cells{rank}.npy', o.state['cell'])
    dist.destroy_process_group()


def test_two_rank_shards_equal_one_process(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'stats.npy')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    o, obs_sum = _rollout(_spec(E, 0))
    want = np.concatenate([o.state['stats'].astype(np.int64).sum(axis=0), [obs_sum]])
    np.testing.assert_array_equal(got, want)
    cells = np.concatenate([np.load(out + f'.cells{r}.npy') for r in range(2)])
    np.testing.assert_array_equal(cells, o.state['cell'])


def test_shard_envs_partitions_exactly():
    from abmarl_b200.dist import shard_envs
    for total in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_envs(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (o1, n1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + n1 == o2
