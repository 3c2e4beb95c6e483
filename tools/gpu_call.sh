set -x
timeout 900 python bench.py > gpurun_out/bench_r01_final5.json 2> gpurun_out/bench_r01_final5.err; head -c 300 gpurun_out/bench_r01_final5.json; tail -2 gpurun_out/bench_r01_final5.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference5.json 2>/dev/null
timeout 600 python tools/profile_step.py > gpurun_out/profile_step_r01_final5.txt 2>&1; head -4 gpurun_out/profile_step_r01_final5.txt | tail -3
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file gpurun_out/launches_r01_final5.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_final5.log 2>&1; tail -c 200 gpurun_out/bench_under_ncu_final5.log
