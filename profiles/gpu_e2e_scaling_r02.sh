# Run under `gpurun --gpus 8`: concurrent PCIe probes at 1/2/4/8 GPUs, the topology, then the bench's host-buffer loop at 8.
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; numactl -H >> gpurun_out/r02_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r02_topo.txt
for N in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N profiles/pcie_probe_multi.py 2>/dev/null | grep '^{' | tee -a gpurun_out/r02_pcie_multi.jsonl
done
for N in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 200 --warmup 20 --kernel-steps 1 --given-steps 4 --e2e-steps 100 2>/dev/null | grep '^{' > gpurun_out/r02_bench_${N}gpu.json
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${N}gpu.json')); print('N=$N value %.3e e2e %.3e strong %s' % (d['value'], d['e2e']['value'], d.get('strong')))"
done
