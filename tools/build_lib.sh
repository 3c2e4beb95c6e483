#!/usr/bin/env bash
# Builds libdualvgr_b200.so (sm_100a only) in-tree. Usage: tools/build_lib.sh [-v]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="$ROOT/dualvgr-videoqa_b200/csrc"
OUT="$ROOT/dualvgr-videoqa_b200/libdualvgr_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -I"$ROOT/include")
if [[ "${1:-}" == "-v" ]]; then FLAGS+=(-Xptxas -v); fi
mkdir -p "$ROOT/build"
pids=()
objs=()
for f in "$SRC"/*.cu; do
  o="$ROOT/build/$(basename "${f%.cu}").o"
  objs+=("$o")
  if [[ ! -f "$o" || "$f" -nt "$o" || -n "$(find "$SRC" -name '*.cuh' -newer "$o" -print -quit)" || -n "$(find "$SRC" "$ROOT/include" -name '*.h' -newer "$o" -print -quit)" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
# fp32-activation variants of the streaming kernel families (same sources, act_t = float, entry points suffixed _f32)
for n in qpm fused gat; do
  f="$SRC/$n.cu"
  o="$ROOT/build/${n}_f32.o"
  objs+=("$o")
  if [[ ! -f "$o" || "$f" -nt "$o" || -n "$(find "$SRC" -name '*.cuh' -newer "$o" -print -quit)" || -n "$(find "$SRC" "$ROOT/include" -name '*.h' -newer "$o" -print -quit)" ]]; then
    "$NVCC" "${FLAGS[@]}" -DDVGR_F32 -c "$f" -o "$o" &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && { wait "$p" || rc=1; }; done
[[ $rc -eq 0 ]] || { echo "compile failed" >&2; exit 1; }
"$NVCC" -shared -o "$OUT" "${objs[@]}" -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
