#!/usr/bin/env python
"""Group an `ncu --page source --csv --print-source cuda,sass` dump of bgw_step_fast_kernel by kernel phase
(line ranges of abmarl_b200/csrc/bgw_fast.cuh found from its section comments).

    python profiles/ncu_phases.py src.csv [n_envs]
"""
import csv
import re
import sys

MARKS = [('exec_attack', r'__device__ void fast_exec_attack'), ('attack rounds', r'__device__ void fast_attack_rounds'),
         ('exec_move', r'void fast_exec_move'), ('touch marks', r'void touch_mark'), ('move rounds (contested)', r'uint32_t fast_move_rounds'),
         ('move phase (uncontested movers)', r'uint32_t fast_move_phase'), ('reset (out of the step path)', r'void fast_reset_env'),
         ('init_dense', r'__device__ void fast_init_dense'), ('obs per-cell path', r'__device__ void fast_obs_chunk_slow'),
         ('obs row gather', r'__device__ void fast_obs_rows'), ('env staging (cp.async)', r'void fast_issue_env'),
         ('kernel prologue', r'__global__ void bgw_step_fast_kernel'), ('per-CTA setup', r'once per CTA'),
         ('env prologue (tickets, staging)', r'for \(; g < NT; g = gn\)'), ('relevant+acting compaction', r'relevant entities and acting'),
         ('order compaction', r'if \(order\) \{'), ('lists+summary', r'occupant lists and summary'),
         ('attack pre-pass', r'attack phase team'), ('settle+classify', r'settle attackers'),
         ('emit reward/done', r'entropy :58'), ('obs dispatch', r'---- observations ----'),
         ('store+all_done+clean', r'store the relevant')]


def main(path, n_envs=4096, src='abmarl_b200/csrc/bgw_fast.cuh'):
    lines = open(src).read().split('\n')
    starts = []
    for name, pat in MARKS:
        for i, l in enumerate(lines, 1):
            if re.search(pat, l):
                starts.append((i, name))
                break
    starts.sort()
    rows = list(csv.reader(open(path)))
    hdr, fname, agg = None, '', {}
    for r in rows:
        if r and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != '-':
            continue
        inst, samp = int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('# Samples')])
        if fname == 'bgw_fast.cuh':
            ln = int(r[0])
            name = 'helpers (cell_rc/pad_index/cenc_of_list)'
            for s, n in starts:
                if ln >= s:
                    name = n
        else:
            name = fname
        tinst = int(r[hdr.index('Predicated-On Thread Instructions Executed')]) if 'Predicated-On Thread Instructions Executed' in hdr else 0
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += inst
        a[1] += samp
        a[2] += tinst
    ti, ts, tt = (sum(v[k] for v in agg.values()) for k in range(3))
    print(f"total warp-instructions {ti} ({ti / n_envs:.0f}/env), samples {ts}, active lanes per instruction {tt / max(ti, 1):.1f}")
    for name, (i, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{name:42s} inst {100 * i / ti:5.1f}% ({i / n_envs:7.0f}/env)   samples {100 * s / ts:5.1f}%   lanes {t / max(i, 1):4.1f}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
