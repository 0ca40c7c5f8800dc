# Run under `gpurun --gpus 8`: bench.py at N = 1, 2, 4, 8 (the driver's launch line), default settings shortened to 400 steps.
for N in ${NS:-1 2 4 8}; do
  if [ $N = 1 ]; then python bench.py --gpus 1 --steps 400 --warmup 50 --no-cpu --e2e-steps 100 --kernel-steps 50 --given-steps 50 2>/dev/null | grep '^{' > gpurun_out/r02_scale_${N}.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 400 --warmup 50 --e2e-steps 100 --kernel-steps 50 --given-steps 50 2>/dev/null | grep '^{' > gpurun_out/r02_scale_${N}.json; fi
  python -c "
import json; d=json.load(open('gpurun_out/r02_scale_${N}.json')); print('N=$N value %.4e ms/step %.5f e2e %.4e strong %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('strong') or {}).get('value')))"
done
python __graft_entry__.py smoke 2>&1 | tail -2
