#!/usr/bin/env bash
# Developer A/B builds of libevogp_b200.so: tools/build_variant.sh <name> [GEN_ENV=1 ...] [-- extra nvcc flags]
# Copies csrc/ to a scratch dir, regenerates the PTX fast paths with the given generator environment, compiles and
# links build_variants/<name>/libevogp_b200.so (git-ignored; travels with gpurun).  Run it with
#   EVOGP_B200_LIB=build_variants/<name>/libevogp_b200.so python tools/time_eval.py
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
name="$1"; shift
envs=(); flags=()
while [ $# -gt 0 ]; do
  if [ "$1" = "--" ]; then shift; flags=("$@"); break; fi
  envs+=("$1"); shift
done
work="$(mktemp -d)"; out="$ROOT/build_variants/$name"; mkdir -p "$out"
mkdir -p "$work/evogp_b200/csrc" "$work/include"
cp "$ROOT"/evogp_b200/csrc/* "$work/evogp_b200/csrc/"; cp "$ROOT"/include/*.h "$work/include/"
(cd "$work/evogp_b200/csrc" && env "${envs[@]}" python gen_fastpath.py > /dev/null)
objs=()
for f in runtime eval eval_exchange eval_acc splice generate nextgen select host_api; do
  nvcc -O3 -std=c++17 -use_fast_math -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -O3 "${flags[@]}" \
       -c "$work/evogp_b200/csrc/$f.cu" -o "$work/$f.o" 2>/dev/null &
  objs+=("$work/$f.o")
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$out/libevogp_b200.so" "${objs[@]}" -lcudart
rm -rf "$work"
echo "$out/libevogp_b200.so"
