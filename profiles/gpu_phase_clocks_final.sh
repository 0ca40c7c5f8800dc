#!/bin/bash
# per-phase clocks of the final kernel inside a fused rollout (BGW_PROFILE build): early-episode steps and late ones
# build first:  BGW_PROFILE=1 BGW_OUT=$PWD/abmarl_b200/csrc/libbgw_prof.so python -m abmarl_b200.csrc.build
export BGW_PROF_FILE=$PWD/gpurun_out/prof_clocks.bin
echo "== steps 5-25"; BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_prof.so BGW_PROF_LAZY=1 python profiles/phase_clocks_chain.py 5 20
echo "== steps 150-170"; BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_prof.so BGW_PROF_LAZY=1 python profiles/phase_clocks_chain.py 150 20
