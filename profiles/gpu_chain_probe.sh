set -x
timeout -k 5 90 python -u tests/chain_probe.py 50 400 1000 3000 > gpurun_out/r01k_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r01k_probe.log
tail -12 gpurun_out/r01k_probe.log
PROBE_HORIZON=12 timeout -k 5 90 python -u tests/chain_probe.py 50 400 1000 3000 > gpurun_out/r01k_probe_h12.log 2>&1; echo "rc=$?" >> gpurun_out/r01k_probe_h12.log
tail -12 gpurun_out/r01k_probe_h12.log
if grep -q "rc=124" gpurun_out/r01k_probe.log gpurun_out/r01k_probe_h12.log; then
  PROBE_HORIZON=12 BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_chaindbg.so timeout -k 5 120 python -u tests/chain_probe.py 50 400 1000 3000 > gpurun_out/r01k_probe_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r01k_probe_dbg.log
  tail -40 gpurun_out/r01k_probe_dbg.log
  exit 0
fi
bash profiles/gpu_ab_chain.sh
