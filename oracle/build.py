"""Build oracle/libbgw_oracle.so from oracle/bgw_oracle.c (plain C, gcc).  Test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'bgw_oracle.c')
OUT = os.path.join(HERE, 'libbgw_oracle.so')


def build(force=False):
    deps = [SRC, os.path.join(HERE, 'bgw_oracle.h'), os.path.join(HERE, '..', 'include', 'bgw.h'),
            os.path.join(HERE, '..', 'include', 'bgw_philox.h')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    # -ffp-contract=off: line-of-sight rays are (a/b)*t in IEEE float64, never fused (utils.py:45-115)
    subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-fno-fast-math',
                           '-o', OUT, SRC, '-lm'])
    return OUT


if __name__ == '__main__':
    print(build(force=True))
