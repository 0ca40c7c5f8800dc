#!/usr/bin/env python
"""Side measurements of the other BASELINE configs (parity-test cases, not bench.py lines): device-resident
rollouts with the keyed random policy, CUDA-event timing.

    python profiles/bench_configs.py [tb_c2|pacman_c3|maze_c1|mm_allstep ...]

BGW_SPECIALIZE=1: compile the general step kernel for each config's spec first (bgw_specialize: NVRTC at run time, cached
in $BGW_JIT_CACHE); the record carries "specialized": true and the compile time.
"""
import time
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abmarl_b200 import _capi as K                      # noqa: E402
from abmarl_b200.engine import BatchedGridWorld         # noqa: E402
from abmarl_b200.spec import compile_sim                # noqa: E402
from tests import scenarios                             # noqa: E402

CONFIGS = {'tb_c2': (16384, 400), 'pacman_c3': (16384, 60), 'maze_c1': (16384, 400), 'tb_blocking': (16384, 200),
           'tb_encoding': (16384, 200), 'tb_restricted': (16384, 200), 'tb_selective_stacked': (16384, 200),
           'tb_ammo_selective': (16384, 200), 'reach_target': (16384, 200), 'traffic': (16384, 200),
           'mm_c4': (16384, 400), 'mm_allstep': (16384, 200), 'pacman_simple': (16384, 100),
           'tb_c5_small': (16384, 200), 'tb_dense': (16384, 200)}


def main(names):
    api = scenarios.mirror_api()
    for name in names:
        n_envs, steps = CONFIGS[name]
        builder, manager, _ = scenarios.SCENARIOS[name]
        spec = compile_sim(builder(api), manager=manager, n_envs=n_envs, seed=7, horizon=200, auto_reset=True)
        eng = BatchedGridWorld(spec, device='cuda:0')
        jit_s = None
        if os.environ.get('BGW_SPECIALIZE'):
            t0 = time.time()
            eng.specialize()
            jit_s = time.time() - t0
        eng.reset()
        per_step = bool(os.environ.get('BGW_PER_STEP_CALLS'))       # one bgw_step_sampled call per step instead of bgw_rollout_sampled
        eng.rollout_sampled(10)
        torch.cuda.synchronize()
        n0 = int(eng.stats()[K.STAT_AGENT_STEPS])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if per_step:
            for _ in range(steps):
                eng.step_sampled()
        else:
            eng.rollout_sampled(steps)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        n = int(eng.stats()[K.STAT_AGENT_STEPS]) - n0
        # algorithmic bytes per agent-step as SURVEY 8(d) counts them: action row + padded int8 obs row + reward 4 + done 1 +
        # agent state read and written 2 x 14
        algo = eng.action_stride + eng.dims.obs_stride + 4 + 1 + 28
        try:
            peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        except Exception:
            peak = 6650.0
        gbs = algo * n / (ms * 1e-3) / 1e9
        print(json.dumps({"algorithmic_bytes_per_agent_step": algo, "achieved_gbs": gbs, "roofline_frac": gbs / peak,"config": name, "envs": n_envs, "learners_per_env": eng.L, "entities_per_env": eng.A, "steps": steps,
                          "ms_per_step": ms / steps, "agent_steps_per_s": n / (ms * 1e-3), "calls": "bgw_step_sampled per step" if per_step else "bgw_rollout_sampled",
                          "kernel": "bgw_step_fast_kernel" if eng.dims.threads_per_env <= 128 and spec.program == K.PROG_TEAM_BATTLE and not (spec.klass & (K.AG_BLOCKING | K.AG_AMMO)).any() and spec.attack_actor <= K.ATTACK_BINARY else "bgw_step_kernel",
                          "specialized": jit_s is not None, "specialize_seconds": jit_s}), flush=True)


if __name__ == '__main__':
    main(sys.argv[1:] or list(CONFIGS))
