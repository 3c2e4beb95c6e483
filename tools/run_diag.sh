#!/usr/bin/env bash
# Runs every GEMM diagnostic case in its own process under a timeout (a hung kernel must not hang the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "$@"; do
  timeout 180 python tools/diag_gemm.py "$c" 2>&1 | tail -40
  echo "--- exit $? for $c"
done
