"""One-off soak: BASELINE configs[4] at full size for several episodes, oracle parity on env slices at both ends of the batch
(every output and the full state, every step), through both the two-launch and the fused sampled path; with a second
argument `rollout`, through chained rollouts (bgw_rollout_sampled) of random lengths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from abmarl_b200 import _capi as K
from abmarl_b200.engine import BatchedGridWorld
from oracle.oracle import OracleEnv
from tests.helpers import assert_state_equal

E, S, STEPS = 4096, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 650
ROLLOUT = len(sys.argv) > 2 and sys.argv[2] == 'rollout'     # chained rollouts of random lengths instead of single steps
spec = bench.build_spec(E, 0)
eng = BatchedGridWorld(spec, device='cuda:0')
head, tail = OracleEnv(spec.with_envs(S, 0)), OracleEnv(spec.with_envs(S, E - S))
eng.reset(); head.reset(); tail.reset()
total = 0
if ROLLOUT:
    rng, t = np.random.default_rng(7), 0
    while t < STEPS:
        n = int(rng.integers(1, 60))
        eng.rollout_sampled(n)
        for o in (head, tail):
            for _ in range(n):
                o.step(o.sample_actions())
        t += n
        st = eng.state_numpy()
        for o, sl in ((head, slice(0, S)), (tail, slice(E - S, E))):
            for name in ('obs', 'done', 'reward', 'all_done'):
                assert np.array_equal(getattr(eng, name)[sl].cpu().numpy(), getattr(o, name)), (t, name)
        assert_state_equal({k: (None if v is None else v[:S]) for k, v in st.items()}, head.state, f'head {t}')
        assert_state_equal({k: (None if v is None else v[E - S:]) for k, v in st.items()}, tail.state, f'tail {t}')
    print(f'soak ok (chained rollouts): {t} steps, {int(eng.stats()[K.STAT_AGENT_STEPS])} agent-steps, {int(eng.stats()[K.STAT_EPISODES])} episodes, '
          f'{int(eng.stats()[K.STAT_KILLS])} kills; {2 * S} envs bit-exact against the oracle after every rollout')
    sys.exit(0)
for t in range(STEPS):
    if t % 2:
        eng.step_sampled()
    else:
        eng.step(eng.sample_actions())
    for o, sl in ((head, slice(0, S)), (tail, slice(E - S, E))):
        o.step(o.sample_actions())
        for name in ('obs', 'done', 'reward', 'all_done'):
            assert np.array_equal(getattr(eng, name)[sl].cpu().numpy(), getattr(o, name)), (t, name)
    if t % 50 == 0 or t == STEPS - 1:
        st = eng.state_numpy()
        assert_state_equal({k: (None if v is None else v[:S]) for k, v in st.items()}, head.state, f'head {t}')
        assert_state_equal({k: (None if v is None else v[E - S:]) for k, v in st.items()}, tail.state, f'tail {t}')
    total += int(((eng.done.cpu().numpy() & K.OUT_VALID) != 0).sum())
assert int(eng.stats()[K.STAT_AGENT_STEPS]) == total
print(f'soak ok: {STEPS} steps, {total} agent-steps, {int(eng.stats()[K.STAT_EPISODES])} episodes, {int(eng.stats()[K.STAT_KILLS])} kills; '
      f'{2 * S} envs bit-exact against the oracle at every step')
