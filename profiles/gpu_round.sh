# The script that produced this round's files (run under gpurun): tests, bench, reference arm, ncu launch list,
# ncu --set full of an early and a late step of an episode, side configs.  R = file prefix.
R=${R:-r01m}
set -x
timeout -k 5 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
timeout -k 5 600 python bench.py --dump-steps gpurun_out/${R}_steps.json > gpurun_out/${R}_bench.log 2> gpurun_out/${R}_bench.err
timeout -k 5 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${R}_ref.log 2>&1
N="--no-cpu --e2e-steps 4 --kernel-steps 1"
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 200 --warmup 20 $N > gpurun_out/${R}_ncu_launch.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 30 --launch-count 1 -o gpurun_out/prof_${R}_early -f python bench.py --steps 40 --warmup 20 $N > gpurun_out/${R}_ncu_early.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 180 --launch-count 1 -o gpurun_out/prof_${R}_late -f python bench.py --steps 200 --warmup 20 $N > gpurun_out/${R}_ncu_late.log 2>&1
tail -3 gpurun_out/${R}_pytest.log; cat gpurun_out/${R}_bench.log
timeout -k 5 900 python profiles/bench_configs.py > gpurun_out/${R}_configs.jsonl 2> gpurun_out/${R}_configs.err
