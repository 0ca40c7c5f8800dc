for rep in 1 2; do
for v in default direct vec both; do
  if [ $v = default ]; then unset BGW_LIB; else export BGW_LIB=/root/repo/abmarl_b200/csrc/libbgw_$v.so; fi
  python bench.py --no-cpu --e2e-steps 4 --e2e-shards 1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('variant', '$v', d['value'], d['ms_per_step'])"
done; done
BGW_LIB=/root/repo/abmarl_b200/csrc/libbgw_both.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
