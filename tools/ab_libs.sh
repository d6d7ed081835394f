#!/bin/bash
# A/B of library builds, ONE process per build (same box, back to back, two passes): tools/ab_libs.sh name=path ...
for pass in 1 2; do
  for spec in "$@"; do
    timeout 300 tools/featbench --lib "$spec" 2>&1 | grep -E "foa \(7|int16|bf16|logmel"
  done
done
