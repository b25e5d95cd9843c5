#!/usr/bin/env bash
# Build the reference's OWN CUDA hot path, unmodified and where it lies under
# /root/reference, into oracle/_ref/ (git-ignored, travels with gpurun).
#   forward.cu  generate.cu  mutation.cu   (kernel.h:23-97 entry points)
# Flags are the reference's (setup.py:48-57) plus the sm_100a gencode; no torch,
# no reference build system.  The result is the tight GPU-side parity oracle and
# the "reference on the same B200" timing baseline.
# A second tiny shim (ref_shim.cpp, ours) includes the reference's kernel.h so
# tests can call its hash() and thrust's taus88 on the host.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF=/root/reference/src/evogp/cuda
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "no $REF here; keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
nvcc -O3 --expt-relaxed-constexpr -Xptxas=-O3 -lineinfo -use_fast_math -maxrregcount=32 \
     -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
     -o "$OUT/libevogp_ref.so" "$REF/forward.cu" "$REF/generate.cu" "$REF/mutation.cu"
g++ -O2 -std=c++17 -shared -fPIC -I/usr/local/cuda/include -I"$REF" \
    -o "$OUT/libref_shim.so" "$HERE/ref_shim.cpp"
echo "built $OUT/libevogp_ref.so $OUT/libref_shim.so"
