# The script that produced the round-2 files (run under gpurun, one GPU): tests, bench (default and the driver's setting),
# reference arm, ncu launch list, ncu --set full of the fused rollout (early-episode steps, and a whole episode), side configs.
R=${R:-r02}
set -x
timeout -k 5 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
timeout -k 5 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
timeout -k 5 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_driver.json 2> gpurun_out/${R}_bench_driver.err
timeout -k 5 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_reference.json 2>&1
N="--no-cpu --e2e-steps 4 --kernel-steps 1 --given-steps 1 --spinup-steps 1"
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 4 --kernel-steps 2 --given-steps 2 --spinup-steps 1 > gpurun_out/${R}_ncu_launch.log 2>&1
# launches of bgw_step_fast in that command: spin-up (1 step), warm-up (5 steps), TIMED (20 steps), ...: skip 2
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${R}_fused_early -f python bench.py --steps 20 --warmup 5 $N > gpurun_out/${R}_ncu_early.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${R}_fused_episode -f python bench.py --steps 200 --warmup 5 $N > gpurun_out/${R}_ncu_episode.log 2>&1
tail -3 gpurun_out/${R}_pytest.log; cat gpurun_out/${R}_bench_driver.json | cut -c1-400
timeout -k 5 900 python profiles/bench_configs.py > gpurun_out/${R}_configs.jsonl 2> gpurun_out/${R}_configs.err
