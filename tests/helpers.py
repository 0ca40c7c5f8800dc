"""Shared comparison code of the parity tests: CUDA engine (through the C-ABI) vs the CPU oracle."""
import numpy as np

from abmarl_b200 import _capi as K

STATE_KEYS = ('flags', 'cell', 'health', 'reward_acc', 'episode', 'step', 'env_flags', 'turn', 'error', 'stats', 'ammo')


def assert_state_equal(eng_state, ora_state, where):
    for k in STATE_KEYS:
        got, want = np.asarray(eng_state[k]), np.asarray(ora_state[k])
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)[:5]
            raise AssertionError(f"{where}: state[{k}] differs at {bad.tolist()}: engine "
                                 f"{got[tuple(bad[0])]} oracle {want[tuple(bad[0])]}")
    in_grid = (np.asarray(ora_state['flags']) & K.ST_IN_GRID) != 0          # `next` is meaningful only in the grid
    got, want = np.asarray(eng_state['next'])[in_grid], np.asarray(ora_state['next'])[in_grid]
    assert np.array_equal(got, want), f"{where}: cell-list order (next) differs"


def assert_outputs_equal(eng, ora, where, rewards_exact=True):
    for name in ('obs', 'done', 'all_done'):
        got, want = getattr(eng, name).cpu().numpy(), getattr(ora, name)
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)[:5]
            raise AssertionError(f"{where}: {name} differs at {bad.tolist()}: engine {got[tuple(bad[0])]} "
                                 f"oracle {want[tuple(bad[0])]}")
    got, want = eng.reward.cpu().numpy(), ora.reward
    if rewards_exact:
        assert np.array_equal(got, want), f"{where}: rewards differ (max |d| {np.abs(got - want).max()})"
    else:
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6, err_msg=where)      # north_star: rewards within 1e-6


def run_lockstep(eng, ora, steps, order_fn=None, check_every=1, label=''):
    """Reset both, then drive both with the keyed random policy for `steps` calls, comparing everything."""
    feeder = None
    if eng.spec.layout_generator is not None:                  # per-episode host-side layouts (MazePlacementState)
        from abmarl_b200.layouts import LayoutFeeder
        feeder = LayoutFeeder(eng.spec)
        rows = feeder.prime(ora.state['episode'])
        if not eng.device_layouts:                             # else the engine generates its own (bgw_generate_layouts)
            eng.set_layout(rows)
        ora.set_layout(rows)
    eng.reset()
    ora.reset()
    assert np.array_equal(eng.obs.cpu().numpy(), ora.obs), f"{label}: reset observations differ"
    assert_state_equal(eng.state_numpy(), ora.state, f"{label} reset")
    agent_steps = 0
    for t in range(steps):
        act_o = ora.sample_actions()
        act_e = eng.sample_actions()
        assert np.array_equal(act_e.cpu().numpy(), act_o), f"{label} step {t}: sampled actions differ"
        order = None if order_fn is None else order_fn(t)
        eng.step(act_e, order)
        ora.step(act_o, order)
        if feeder is not None and feeder.after_step(ora.all_done, ora.state['episode']):
            if not eng.device_layouts:
                eng.set_layout(feeder.rows)
            ora.set_layout(feeder.rows)
        if t % check_every == 0 or t == steps - 1:
            assert_outputs_equal(eng, ora, f"{label} step {t}")
            assert_state_equal(eng.state_numpy(), ora.state, f"{label} step {t}")
        agent_steps += int(((ora.done & K.OUT_VALID) != 0).sum())
    return agent_steps
