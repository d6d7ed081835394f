#!/bin/bash
# Build a variant of the library for A/B runs: tools/build_variant.sh NAME "-DFLAG=1 ..." -> tools/var_NAME/libseld_cuda.so
# The variant links the CUDA runtime STATICALLY: two libraries that share one libcudart and contain kernels of the same
# name were seen to run one copy for both, so every build under test carries its own runtime.
set -e
name=$1; flags=$2
here=$(cd "$(dirname "$0")" && pwd)
src=$here/../sound-event-localization-detection_b200/csrc
tmp=$(mktemp -d)
for f in "$src"/*.cu; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr $flags -c "$f" -o "$tmp/$(basename "${f%.cu}").o" &
done
wait
mkdir -p "$here/var_$name"
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o "$here/var_$name/libseld_cuda.so" "$tmp"/*.o
rm -rf "$tmp"
echo "built tools/var_$name/libseld_cuda.so [$flags]"
