#!/usr/bin/env bash
# Round-2 profile set (run under gpurun): launch list of the bench command, per-kernel section captures of one eager step at
# the sizes of BASELINE configs 2, 4, 5, and a full capture (with source) of the dominant kernel. Raw CSVs land in gpurun_out/.
set -u
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-eager --no-dp-parity --no-kernel-census > gpurun_out/r02_bench_under_ncu.log 2>&1
for cfg in svqa msvd_u3 clip64; do
  CONFIG=$cfg timeout 900 ncu $SECT --clock-control none --profile-from-start off -f -o gpurun_out/r02_step_$cfg \
      python tools/ncu_step.py > gpurun_out/r02_step_$cfg.log 2>&1
  ncu -i gpurun_out/r02_step_$cfg.ncu-rep --page raw --csv > gpurun_out/r02_step_${cfg}_raw.csv 2>/dev/null
  ls -la gpurun_out/r02_step_$cfg.ncu-rep
  rm -f gpurun_out/r02_step_$cfg.ncu-rep
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lstm_seq_fwd -c 2 -f \
    -o gpurun_out/r02_lstm_seq_fwd python tools/ncu_step.py > gpurun_out/r02_lstm_seq_fwd.log 2>&1
ncu -i gpurun_out/r02_lstm_seq_fwd.ncu-rep --page raw --csv > gpurun_out/r02_lstm_seq_fwd_raw.csv 2>/dev/null
ls -la gpurun_out/
