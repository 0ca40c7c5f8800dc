set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01j_pytest.log
python bench.py --dump-steps gpurun_out/r01j_steps.json > gpurun_out/r01j_bench.log 2> gpurun_out/r01j_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r01j_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_launches.csv python bench.py --steps 200 --warmup 20 --no-cpu --e2e-steps 4 > gpurun_out/r01j_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 30 --launch-count 1 -o gpurun_out/prof_r01j_early -f python bench.py --steps 40 --warmup 20 --no-cpu --e2e-steps 4 > gpurun_out/r01j_ncu_early.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 180 --launch-count 1 -o gpurun_out/prof_r01j_late -f python bench.py --steps 200 --warmup 20 --no-cpu --e2e-steps 4 > gpurun_out/r01j_ncu_late.log 2>&1
tail -3 gpurun_out/r01j_pytest.log; cat gpurun_out/r01j_bench.log
python profiles/bench_configs.py > gpurun_out/r01j_configs.jsonl 2> gpurun_out/r01j_configs.err
