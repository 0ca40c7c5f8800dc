python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "observe or degenerate" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 4 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.readline()); o = r['roofline']['observe_kernel']
print('ms/step %.5f' % r['ms_per_step'], 'frac %.4f' % r['roofline']['frac'], 'observe ms %.5f GB/s %.0f frac %.3f' % (o['ms_per_launch'], o['achieved'], o['frac']))"
