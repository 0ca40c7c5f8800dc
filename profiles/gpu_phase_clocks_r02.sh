export BGW_PROF_FILE=$PWD/gpurun_out/prof_clocks.bin
echo "== b2p chained (lazy dump)"; BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_b2p.so BGW_PROF_LAZY=1 python profiles/phase_clocks_chain.py 5 20
echo "== b2p chain off"; BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_b2p.so BGW_PROF_LAZY=1 BGW_CHAIN=0 python profiles/phase_clocks_chain.py 5 20
echo "== b1p (dump per launch: serialised)"; BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_b1p.so python profiles/phase_clocks_chain.py 5 20
