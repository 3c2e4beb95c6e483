timeout 900 python bench.py > gpurun_out/bench_r01_final6.json 2> gpurun_out/bench_r01_final6.err; head -c 260 gpurun_out/bench_r01_final6.json; echo; tail -2 gpurun_out/bench_r01_final6.err
timeout 600 python tools/profile_step.py > gpurun_out/profile_step_r01_final6.txt 2>&1; head -4 gpurun_out/profile_step_r01_final6.txt | tail -2
