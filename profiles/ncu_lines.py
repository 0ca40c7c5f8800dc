#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/ncu_lines.py src.csv [top_n]
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = {}
    fname = ''
    for r in rows:
        if r and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != '-':     # per-line summary rows have '-' as address
            continue
        i_inst, i_samp = hdr.index('Instructions Executed'), hdr.index('# Samples')
        agg[(fname, int(r[0]))] = (int(r[i_inst]), int(r[i_samp]), r[1])
    tot_i = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for line, (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{line[0][:14]:14s}{line[1]:5d} inst {100 * inst / tot_i:5.1f}%  samples {100 * samp / tot_s:5.1f}%  {src.strip()[:110]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
