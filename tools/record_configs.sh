#!/usr/bin/env bash
# Recorded bench runs of the other BASELINE.json configs (1 GPU): one JSON line each into gpurun_out/r02_bench_<config>.json
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_svqa.json 2> gpurun_out/r02_bench_svqa.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
for cfg in msrvtt msvd_u1 msvd_u3 msvd_u5 clip64; do
  python bench.py --config $cfg --no-eager --no-cpu --steps 10 --warmup 3 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err
done
python bench.py --dtype fp32 --no-eager --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_svqa_fp32.json 2>/dev/null
for f in gpurun_out/r02_bench_*.json; do echo "$f: $(tail -1 $f | cut -c1-330)"; done
